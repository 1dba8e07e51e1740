"""Definitions of the golden cases, shared by ``make_golden.py`` (reference run, build container)
and the tests (oracle / CUDA run, anywhere).  Inputs are rebuilt from the committed fixtures only —
nothing here reads ``/root/reference``."""
import os

import numpy as np
import pandas as pd
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))


class AnnDataLike:
    def __init__(self, obs, connectivities):
        self.obs = obs
        self.obsp = {"connectivities": connectivities}

    @property
    def n_obs(self):
        return len(self.obs)


# ---------------------------------------------------------------------------------------------
# demo fixture (reference demo/data.h5ad; demo/makedata.ipynb cells 3-4 for the obs layout)
# ---------------------------------------------------------------------------------------------
def load_demo_graph():
    return dict(np.load(os.path.join(HERE, "demo_graph.npz")))


def demo_anndata(g=None):
    g = load_demo_graph() if g is None else g
    n = len(g["indptr"]) - 1
    A = sp.csr_matrix((g["data"], g["indices"], g["indptr"]), shape=(n, n))
    obs = pd.DataFrame({c: g["obs_" + c] for c in ("id", "case", "male", "batch")},
                       index=pd.Index([str(i) for i in range(n)]))
    return AnnDataLike(obs, A)


def demo_sample_meta(g=None):
    g = load_demo_graph() if g is None else g
    obs = pd.DataFrame({c: g["obs_" + c] for c in ("id", "case", "male", "batch")})
    return obs.groupby("id")[["case", "male", "batch"]].mean()


DEMO_CASES = {
    # demo.ipynb cell 10: prints p = 0.000999000999000999 and 9555 neighbourhoods at FDR 5 %
    "case_male_batch": dict(y="case", covs=["male"], batches="batch", np_seed=0,
                            key_added="case_coef"),
    "male_case_batch": dict(y="male", covs=["case"], batches="batch", np_seed=0,
                            key_added="male_coef"),
    "case_plain": dict(y="case", seed=7, Nnull=500, nsteps=3),
    "donor": dict(y="donor_pheno", donors=True, seed=3, Nnull=200, nsteps=3),
    "covs_only": dict(y="case", covs=["male"], seed=11, Nnull=300, nsteps=2, ks=[1, 3, 5]),
}


def build_demo_case(g, spec):
    data = demo_anndata(g)
    meta = demo_sample_meta(g)
    kwargs = {}
    if spec.get("donors"):
        don = pd.Series(np.arange(50) // 2, index=meta.index)
        kwargs["y"] = (don >= 13).astype(float)
        kwargs["donorids"] = don
    else:
        kwargs["y"] = meta[spec["y"]]
    if "covs" in spec:
        kwargs["covs"] = meta[spec["covs"]]
    if "batches" in spec:
        kwargs["batches"] = meta[spec["batches"]]
    for k in ("seed", "Nnull", "nsteps", "key_added", "ks"):
        if k in spec:
            kwargs[k] = spec[k]
    kwargs["sid_name"] = "id"
    return data, kwargs


# ---------------------------------------------------------------------------------------------
# small synthetic edge cases (graphs are stored in synth_cases.npz as <case>/graph_*)
# ---------------------------------------------------------------------------------------------
SYNTH_CASES = {
    # unequal cells per sample, two samples with NaN phenotype / covariate, phenotype given in
    # shuffled sample order, one extra sample id that is absent from the data
    "ragged_nan_shuffled": dict(n_cells=3000, n_samples=24, k=10, seed=3, ragged=True,
                                nan_y=[2], nan_cov=[7], shuffle=True, extra_id=True, isolate=[2],
                                covs=["age"], batches=True, call=dict(seed=5, Nnull=200, nsteps=3)),
    # no batches, no covariates, auto nsteps, string sample ids held in a categorical column
    "categorical_auto": dict(n_cells=2400, n_samples=20, k=12, seed=4, categorical=True,
                             call=dict(seed=2, Nnull=150)),
    # batches strong enough to trip QC: one batch gets its own shifted cell population
    "batchy_qc": dict(n_cells=4000, n_samples=40, k=10, seed=6, batches=True, n_batches=10,
                      batch_shift=[0], covs=["age"], call=dict(seed=9, Nnull=200, nsteps=3)),
    # every batch has a nearly private cell population: QC keeps most cells (threshold = 2 x
    # median) and the ridge loop has to walk down several ridge values
    "all_batchy_ridgewalk": dict(n_cells=3000, n_samples=30, k=10, seed=7, batches=True,
                                 n_batches=10, batch_shift=list(range(10)),
                                 call=dict(seed=4, Nnull=120, nsteps=3)),
    # custom ks and ridges, many permutations (> 1000 so the local test truncates)
    "ks_ridges": dict(n_cells=2000, n_samples=40, k=8, seed=8, batches=True,
                      call=dict(seed=1, Nnull=1500, nsteps=2, ks=[2, 5], ridges=[10.0, 0.1, 0])),
}


def make_synth_inputs(spec):
    """Build the raw inputs of a synthetic case on the CPU (used by make_golden.py)."""
    from cna_b200 import synth
    data, meta = synth.make_dataset(spec["n_cells"], spec["n_samples"], spec["k"], seed=spec["seed"],
                                    ragged=spec.get("ragged", False), knn="cpu", device="cpu",
                                    n_batches=spec.get("n_batches", 4))
    A = data.obsp["connectivities"]
    sid = data.obs["id"].to_numpy()
    if spec.get("batch_shift") is not None or spec.get("isolate"):
        # rewire: drop (most) edges that cross the boundary of the listed batches, which makes
        # their neighbourhoods extremely batch specific; "isolate" cuts a sample off completely
        b = meta["batch"].to_numpy()[sid]
        coo = A.tocoo()
        drop = np.zeros(coo.nnz, dtype=bool)
        for bb in spec.get("batch_shift") or []:
            drop |= (b[coo.row] == bb) ^ (b[coo.col] == bb)
        rng = np.random.default_rng(0)
        drop &= rng.random(len(drop)) < 0.9
        for ss in spec.get("isolate") or []:
            drop |= (sid[coo.row] == ss) ^ (sid[coo.col] == ss)
        key = np.minimum(coo.row, coo.col).astype(np.int64) * A.shape[0] + np.maximum(coo.row, coo.col)
        dropkeys = np.unique(key[drop])
        keep = ~np.isin(key, dropkeys)
        A = sp.csr_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=A.shape)
        A.sort_indices()
    return dict(data=A.data, indices=A.indices.astype(np.int32), indptr=A.indptr.astype(np.int32),
                sid=sid, case=meta["case"].to_numpy(), batch=meta["batch"].to_numpy(),
                age=meta["age"].to_numpy())


def build_synth_case(spec, raw=None):
    raw = make_synth_inputs(spec) if raw is None else raw
    n = len(raw["indptr"]) - 1
    A = sp.csr_matrix((raw["data"], raw["indices"], raw["indptr"]), shape=(n, n))
    S = len(raw["case"])
    sid = raw["sid"]
    names = np.array([f"s{i:02d}" for i in range(S)])
    if spec.get("categorical"):
        col = pd.Categorical(names[sid], categories=names)
        index = pd.Index(names)
    else:
        col = sid
        index = pd.Index(np.arange(S))
    obs = pd.DataFrame({"id": col}, index=pd.Index([f"c{i}" for i in range(n)]))
    data = AnnDataLike(obs, A)
    meta = pd.DataFrame({"case": raw["case"], "batch": raw["batch"], "age": raw["age"]}, index=index)
    meta = meta.copy()
    for i in spec.get("nan_y", []):
        meta.iloc[i, meta.columns.get_loc("case")] = np.nan
    for i in spec.get("nan_cov", []):
        meta.iloc[i, meta.columns.get_loc("age")] = np.nan
    if spec.get("extra_id"):
        extra = pd.DataFrame({"case": [1.0], "batch": [0], "age": [0.3]}, index=pd.Index([S + 5]))
        meta = pd.concat([meta, extra])
    if spec.get("shuffle"):
        meta = meta.iloc[np.random.default_rng(42).permutation(len(meta))]
    kwargs = dict(y=meta["case"], sid_name="id")
    if spec.get("covs"):
        kwargs["covs"] = meta[spec["covs"]]
    if spec.get("batches"):
        kwargs["batches"] = meta["batch"]
    kwargs.update(spec["call"])
    data._raw = raw
    return data, kwargs
