"""Generate the committed golden fixtures by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py

Needs ``/root/reference`` (present only in the build container; loaded through ``oracle/ref_shim.py``,
never copied).  Writes, next to this script:

  demo_graph.npz      the kNN graph + obs columns of the reference's bundled ``demo/data.h5ad``
                      (data fixture, read with cna_b200.utils.h5min — no h5py/anndata here)
  demo_knn.npz        the kNN distances stored in the same file
  demo_cases.npz/json reference outputs for the demo analyses listed in SURVEY.md section 8(c)
  synth_cases.npz/json reference outputs on small synthetic inputs that exercise the edge cases
                      (ragged samples, NaN phenotype, extra / shuffled sample ids, donors, ...)

The ``.json`` files hold scalars, the ``.npz`` files arrays, keyed ``<case>/<field>``.
"""
import json
import os
import sys
import warnings

import numpy as np
import pandas as pd
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from cna_b200.utils.h5min import H5File  # noqa: E402
from cna_b200 import synth  # noqa: E402
from tests.golden import cases  # noqa: E402


def export_demo_graph():
    f = H5File("/root/reference/demo/data.h5ad")
    g = "uns/neighbors/connectivities/"
    out = dict(data=f.read(g + "data"), indices=f.read(g + "indices"), indptr=f.read(g + "indptr"))
    for col in ("id", "case", "male", "batch"):
        out["obs_" + col] = f.read("obs/" + col)
    np.savez_compressed(os.path.join(HERE, "demo_graph.npz"), **out)
    # the kNN distances scanpy stored next to the graph (14 per row): what the stored connectivities were built
    # from — pins cna_b200.pp.fuzzy_simplicial_set
    d = "uns/neighbors/distances/"
    np.savez_compressed(os.path.join(HERE, "demo_knn.npz"), data=f.read(d + "data"),
                        indices=f.read(d + "indices"), indptr=f.read(d + "indptr"),
                        n_neighbors=f.read("uns/neighbors/params/n_neighbors"))
    return out


def record(res, data, key, arrays, scalars, name, warns):
    """Flatten a reference result Namespace into the arrays / scalars dictionaries."""
    sc = {}
    for fld in ("p", "k", "r", "r2", "fdr_5p_t", "fdr_10p_t", "nullr2_mean", "nullr2_std"):
        v = getattr(res, fld)
        sc[fld] = None if v is None else float(v)
    sc["ks"] = [int(k) for k in res.ks]
    sc["n_fdr05"] = int((data.obs[key + "_fdr"] <= 0.05).sum())
    sc["n_fdr10"] = int((data.obs[key + "_fdr"] <= 0.10).sum())
    sc["n_kept"] = int(res.kept.sum())
    sc["warnings"] = warns
    scalars[name] = sc
    arrays[name + "/kept"] = np.asarray(res.kept)
    arrays[name + "/ncorrs"] = res.ncorrs.to_numpy()
    arrays[name + "/coef"] = data.obs[key].to_numpy()
    arrays[name + "/coef_fdr"] = data.obs[key + "_fdr"].to_numpy()
    arrays[name + "/svs"] = res.namresid_svs.to_numpy()
    arrays[name + "/varexp"] = res.namresid_varexp.to_numpy()
    arrays[name + "/U"] = res.namresid_sampleXpc.to_numpy()
    arrays[name + "/M"] = res.M.to_numpy()
    arrays[name + "/nullminps"] = np.asarray(res.nullminps)
    arrays[name + "/beta"] = np.asarray(res.beta)
    arrays[name + "/r2_perpc"] = np.asarray(res.r2_perpc)
    arrays[name + "/yresid"] = res.yresid.to_numpy()
    arrays[name + "/yresid_hat"] = np.asarray(res.yresid_hat)
    arrays[name + "/fdrs"] = res.fdrs.to_numpy(dtype=np.float64)
    # a thin slice of the big matrices is enough to pin them
    arrays[name + "/namresid_head"] = res.namresid.to_numpy()[:, :64]
    arrays[name + "/nam_head"] = res.nam.to_numpy()[:, :64]
    arrays[name + "/V_head"] = res.namresid_nbhdXpc.to_numpy()[:64, :8]


def run_cases(ref, make, specs, arrays, scalars):
    for name, spec in specs.items():
        data, kwargs = make(spec)
        for k, v in getattr(data, "_raw", {}).items():
            arrays[name + "/raw_" + k] = v
        if spec.get("np_seed") is not None:
            np.random.seed(spec["np_seed"])
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            res = ref.association(data, return_full=True, **kwargs)
        warns = sorted({str(x.message)[:60] for x in w if issubclass(x.category, UserWarning)})
        record(res, data, kwargs.get("key_added", "coef"), arrays, scalars, name, warns)
        print(name, "p =", res.p, "k =", res.k, "fdr05 =", scalars[name]["n_fdr05"])


def main():
    ref = ref_shim.load()
    g = export_demo_graph()

    arrays, scalars = {}, {}
    run_cases(ref, lambda s: cases.build_demo_case(g, s), cases.DEMO_CASES, arrays, scalars)
    # nam() / svd_nam() / diffuse() pins
    data = cases.demo_anndata(g)
    meta = cases.demo_sample_meta()
    NAM, keep = ref.nam(data, "id", batches=meta.batch)
    U, svs, V = ref.svd_nam(NAM)
    arrays["nam/NAM_head"] = NAM.to_numpy()[:, :256]
    arrays["nam/rowsum"] = NAM.to_numpy().sum(axis=1)
    arrays["nam/keep"] = keep
    arrays["nam/svs"] = svs.to_numpy()
    arrays["nam/U"] = U.to_numpy()
    arrays["nam/V_head"] = V.to_numpy()[:64, :8]
    for s in (1, 2, 3):
        arrays[f"nam/steps{s}_head"] = ref.nam(data, "id", nsteps=s)[0].to_numpy()[:, :256]
    rng = np.random.default_rng(5)
    s0 = rng.normal(size=(data.n_obs, 3))
    arrays["diffuse/s0"] = s0
    arrays["diffuse/s2"] = np.asarray(ref.diffuse(data, s0, 2))
    arrays["diffuse/s3_w05"] = np.asarray(ref.diffuse(data, s0, 3, self_weight=0.5))
    np.savez_compressed(os.path.join(HERE, "demo_cases.npz"), **arrays)
    json.dump(scalars, open(os.path.join(HERE, "demo_cases.json"), "w"), indent=1, sort_keys=True)

    arrays, scalars = {}, {}
    run_cases(ref, cases.build_synth_case, cases.SYNTH_CASES, arrays, scalars)
    np.savez_compressed(os.path.join(HERE, "synth_cases.npz"), **arrays)
    json.dump(scalars, open(os.path.join(HERE, "synth_cases.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
