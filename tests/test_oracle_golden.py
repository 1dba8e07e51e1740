"""CPU: pin the oracle (oracle/cna_oracle.py) against reference outputs.

The golden files hold outputs of the UNMODIFIED reference (tests/golden/make_golden.py); two of
them are also printed in the reference's own demo notebook (p = 0.000999000999000999 and 9555 /
4509 neighbourhoods at FDR 5 %).  When the reference checkout is present (build container only) a
further test cross-checks oracle and reference live on a fresh random input.
"""
import numpy as np
import pandas as pd
import pytest

from oracle import cna_oracle as orc
from oracle import ref_shim
from tests import helpers
from tests.golden import cases


@pytest.fixture(scope="module")
def demo():
    return helpers.load_golden("demo_cases")


@pytest.fixture(scope="module")
def synth():
    return helpers.load_golden("synth_cases")


def test_notebook_printed_values(demo):
    """demo/demo.ipynb cell 10 stdout."""
    _, sc = demo
    assert sc["case_male_batch"]["p"] == 0.000999000999000999
    assert sc["case_male_batch"]["n_fdr05"] == 9555
    assert sc["male_case_batch"]["n_fdr05"] == 4509


@pytest.mark.parametrize("name", list(cases.DEMO_CASES))
@pytest.mark.parametrize("faithful", [False, True])
def test_oracle_demo_cases(demo, name, faithful):
    if faithful and name not in ("case_plain", "donor"):
        pytest.skip("faithful mode is exercised on the two cheapest cases")
    arrays, scalars = demo
    spec = cases.DEMO_CASES[name]
    data, kwargs = cases.build_demo_case(cases.load_demo_graph(), spec)
    res, warns = helpers.run_association(orc.association, data, kwargs, spec.get("np_seed"),
                                         faithful=faithful)
    helpers.assert_matches_golden(res, data, kwargs.get("key_added", "coef"), arrays, scalars, name,
                                  warns=warns)


@pytest.mark.parametrize("name", list(cases.SYNTH_CASES))
def test_oracle_synth_cases(synth, name):
    arrays, scalars = synth
    spec = cases.SYNTH_CASES[name]
    data, kwargs = cases.build_synth_case(spec, helpers.synth_raw(arrays, name))
    res, warns = helpers.run_association(orc.association, data, kwargs)
    helpers.assert_matches_golden(res, data, "coef", arrays, scalars, name, warns=warns)


def test_oracle_nam_svd_diffuse(demo):
    arrays, _ = demo
    data = cases.demo_anndata()
    meta = cases.demo_sample_meta()
    diag = {}
    NAM, keep = orc.nam(data, "id", batches=meta.batch, diagnostics=diag)
    # auto-stop after 4 steps with these median kurtoses
    assert diag["nsteps"] == 4
    np.testing.assert_allclose(diag["medkurt"], [16.40326205754668, 12.757992457504468,
                                                 6.115484918150449, 3.177490071480447], rtol=1e-10)
    np.testing.assert_array_equal(keep, arrays["nam/keep"])
    np.testing.assert_allclose(NAM.to_numpy()[:, :256], arrays["nam/NAM_head"], rtol=1e-10)
    np.testing.assert_allclose(NAM.to_numpy().sum(axis=1), arrays["nam/rowsum"], rtol=1e-10)
    U, svs, V = orc.svd_nam(NAM)
    np.testing.assert_allclose(svs.to_numpy()[:-1], arrays["nam/svs"][:-1], rtol=1e-9)
    np.testing.assert_allclose(helpers.sign_align(U.to_numpy()[:, :5], arrays["nam/U"][:, :5]),
                               arrays["nam/U"][:, :5], atol=1e-9)
    for s in (1, 2, 3):
        got = orc.nam(data, "id", nsteps=s)[0].to_numpy()[:, :256]
        np.testing.assert_allclose(got, arrays[f"nam/steps{s}_head"], rtol=1e-10)
    A = data.obsp["connectivities"]
    np.testing.assert_allclose(orc.diffuse(A, arrays["diffuse/s0"], 2), arrays["diffuse/s2"], rtol=1e-12)
    np.testing.assert_allclose(orc.diffuse(A, arrays["diffuse/s0"], 3, self_weight=0.5),
                               arrays["diffuse/s3_w05"], rtol=1e-12)


def test_oracle_input_errors():
    """Exception types of _association.py:131-173, :29-33."""
    data = cases.demo_anndata()
    meta = cases.demo_sample_meta()
    with pytest.raises(TypeError):
        orc.association(data, meta.case.to_numpy(), "id")
    with pytest.raises(TypeError):
        orc.association(data, meta.case, "id", covs=meta.male)
    with pytest.raises(ValueError):
        orc.association(data, meta.case.iloc[:40], "id")
    with pytest.raises(ValueError):
        orc.association(data, meta.case, "id", batches=meta.batch, donorids=meta.batch)
    y = meta.case.copy()
    y.iloc[:45] = np.nan
    with pytest.raises(ValueError, match="fewer than 10 samples"):
        orc.association(data, y, "id")
    with pytest.raises(ValueError, match="Maximum number of PCs"):
        orc.association(data, meta.case, "id", ks=[50], nsteps=1, Nnull=10)


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout only exists in the build container")
def test_oracle_vs_live_reference():
    ref = ref_shim.load()
    from cna_b200 import synth as gen
    data, meta = gen.make_dataset(1500, 16, 8, seed=21, ragged=True, knn="cpu", device="cpu")
    d2 = cases.AnnDataLike(data.obs.copy(), data.obsp["connectivities"])
    kw = dict(y=meta.case, sid_name="id", covs=meta[["age"]], batches=meta.batch, seed=13, Nnull=100)
    a, _ = helpers.run_association(ref.association, data, kw)
    b, _ = helpers.run_association(orc.association, d2, kw)
    assert a.p == b.p and a.k == b.k
    np.testing.assert_allclose(b.ncorrs.to_numpy(), a.ncorrs.to_numpy(), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(b.nullminps, a.nullminps, rtol=1e-7)
    np.testing.assert_allclose(d2.obs["coef_fdr"].to_numpy(), data.obs["coef_fdr"].to_numpy(), atol=1e-12)
