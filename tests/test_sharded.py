"""Cell-axis sharding (cna_b200/sharded.py): world_size-2 tests.

CPU (gloo): the host-side plumbing — shard bounds, CSR slicing, the collective helpers, the global
median.  GPU (``-m gpu``; two ranks sharing one device over gloo, so it runs on a single-GPU box):
the sharded association() must reproduce the single-GPU result.
"""
import os
import socket
import warnings

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.golden import cases


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(fn, world, *args):
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn, args), nprocs=world, join=True)


def _entry(rank, world, port, fn, args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def _spawn_nccl(fn, world, *args):
    port = _free_port()
    mp.spawn(_entry_nccl, args=(world, port, fn, args), nprocs=world, join=True)


def _entry_nccl(rank, world, port, fn, args):
    """One rank per GPU over NCCL: the configuration bench.py --gpus N runs (all_to_all_single halo
    exchange, NCCL all-reduces / all-gathers)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def test_shard_bounds_and_slicing():
    from cna_b200.sharded import shard_bounds, slice_csr
    for n, w in [(10, 2), (11, 2), (7, 8), (1_000_000, 8), (5, 1)]:
        blocks = [shard_bounds(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        assert all(b[1] - b[0] <= b[2] for b in blocks) and len({b[2] for b in blocks}) == 1
        assert all(b[1] - b[0] == b[2] for b in blocks if b[1] < n)  # only the tail may be short
    A = sp.random(50, 50, density=0.2, format="csr", random_state=0)
    ip, ix, dv = slice_csr(A, 13, 31)
    B = sp.csr_matrix((dv, ix, ip), shape=(18, 50))
    assert (B != A[13:31]).nnz == 0


def _comm_checks(rank, world):
    from cna_b200.sharded import Comm
    comm = Comm()
    assert (comm.rank, comm.world) == (rank, world)
    t = torch.full((3,), float(rank + 1), dtype=torch.float64)
    assert comm.all_reduce(t.clone()).tolist() == [3.0] * 3
    assert comm.all_reduce(t.clone(), op="max").tolist() == [2.0] * 3
    g = comm.all_gather_rows(torch.full((2, 3), float(rank)))
    assert g.shape == (4, 3) and g[:2].eq(0).all() and g[2:].eq(1).all()
    parts = comm.all_gather_padded(torch.arange(3 + 2 * rank, dtype=torch.int64) + 10 * rank)
    assert [p.tolist() for p in parts] == [[0, 1, 2], [10, 11, 12, 13, 14]]
    assert [p.numel() for p in comm.all_gather_padded(torch.zeros(0, dtype=torch.int64))] == [0, 0]
    b = torch.arange(5, dtype=torch.int32) if rank == 0 else torch.zeros(5, dtype=torch.int32)
    assert comm.broadcast(b).tolist() == list(range(5))


def test_comm_helpers_gloo_world2():
    _spawn(_comm_checks, 2)


def _median_checks(rank, world):
    """global median == numpy median of the concatenation, with masks, NaNs and a short tail shard"""
    from cna_b200.sharded import Comm, shard_bounds
    from cna_b200.tl._nam import device_median
    torch.cuda.set_device(0)
    comm = Comm()
    rng = np.random.default_rng(0)
    for n in (11, 12, 2, 1, 5001):
        full = rng.normal(size=n)
        valid = rng.random(n) > 0.3
        valid[0] = True
        r0, r1, rows_per = shard_bounds(n, world, rank)
        loc = torch.as_tensor(full[r0:r1]).cuda()
        assert device_median(loc, comm=comm, rows_per=rows_per) == np.median(full)
        got = device_median(loc, valid=torch.as_tensor(valid[r0:r1]).cuda(), comm=comm, rows_per=rows_per)
        assert got == np.median(full[valid])
        full[0] = np.nan
        assert np.isnan(device_median(torch.as_tensor(full[r0:r1]).cuda(), comm=comm, rows_per=rows_per))


@pytest.mark.gpu
def test_sharded_median_world2():
    _spawn(_median_checks, 2)


def _sharded_vs_single(rank, world, spec, reorder=False, own_gpu=False):
    import cna_b200 as cna
    from cna_b200.sharded import shard_to_device
    torch.cuda.set_device(rank if own_gpu else 0)
    if reorder:  # shards of the Cuthill-McKee ordered graph: small halo, results in the caller's order
        os.environ["CNA_B200_REORDER"] = "1"
    g = cases.load_demo_graph()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d1, kw = cases.build_demo_case(g, spec)
        p1 = cna.tl.association(d1, **kw)
        d2, kw = cases.build_demo_case(g, spec)
        p2 = cna.tl.association(shard_to_device(d2), **kw)
    key = kw.get("key_added", "coef")
    assert p1 == p2
    # the residualisation pass is row-local, so coefficients agree bit for bit; FDRs go through the
    # all-reduced Gram (summation order) only via U, which does not enter them at all
    np.testing.assert_array_equal(d1.obs[key].to_numpy(), d2.obs[key].to_numpy())
    np.testing.assert_allclose(d1.obs[key + "_fdr"].to_numpy(), d2.obs[key + "_fdr"].to_numpy(), rtol=1e-12)
    sh = shard_to_device(d2)
    if reorder:
        assert sh.graph.order is not None
        # the halo is a subset of the other shard's rows (the 10k-cell demo graph has hubs of degree
        # 660, so it is not small here; on the 1M-cell benchmark graph it is about one shard)
        assert 0 < sh.graph.halo_ids.numel() <= len(d2.obs) - sh.graph.n


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["case_male_batch", "case_plain", "donor"])
def test_sharded_association_matches_single_gpu(name):
    spec = dict(cases.DEMO_CASES[name])
    spec.pop("np_seed", None)
    spec.setdefault("seed", 0)
    _spawn(_sharded_vs_single, 2, spec)


@pytest.mark.gpu
def test_sharded_association_on_reordered_graph():
    spec = dict(cases.DEMO_CASES["case_male_batch"])
    spec.pop("np_seed", None)
    spec.setdefault("seed", 0)
    _spawn(_sharded_vs_single, 2, spec, True)


def _sharded_auto_stop(rank, world):
    from cna_b200.sharded import shard_to_device
    from cna_b200.tl import _nam
    torch.cuda.set_device(0)
    data = cases.demo_anndata()
    meta = cases.demo_sample_meta()
    st1 = _nam._nam_device(data, "id")
    st2 = _nam._nam_device(shard_to_device(cases.demo_anndata()), "id")
    assert st1.nsteps == st2.nsteps == 4
    np.testing.assert_allclose(st2.medkurt, st1.medkurt, rtol=1e-12)
    r0 = st2.row0
    assert torch.equal(st1.s[r0:r0 + st2.N], st2.s)
    _nam._qc_device(st1, meta.batch)
    _nam._qc_device(st2, meta.batch)
    assert st1.qc_threshold == st2.qc_threshold


@pytest.mark.gpu
def test_sharded_nam_auto_stop_and_qc():
    _spawn(_sharded_auto_stop, 2)


def _nccl_halo_and_large(rank, world):
    """NCCL path at a size where the graph is stored reordered and shards have a real halo: the
    all_to_all_single exchange itself (against rows fetched directly), then whole association() against the
    single-GPU result: exact p and kept cells, identical coefficients."""
    import cna_b200 as cna
    from cna_b200 import synth
    from cna_b200.sharded import shard_to_device
    data, meta = synth.make_dataset(200_000, 60, 15, seed=1)
    kw = dict(y=meta.case, sid_name="id", batches=meta.batch, covs=meta[["age"]], nsteps=3, Nnull=500, seed=2)
    sh = shard_to_device(data)
    g = sh.graph
    assert g.comm.backend == "nccl" and g.order is not None and g.halo_ids.numel() > 0
    # the exchange: every rank's extended state must end with the owners' rows
    ld = 64
    full = torch.arange(g.n_total * ld, dtype=torch.float32, device="cuda").reshape(g.n_total, ld)
    ext = torch.zeros((g.rows_per + g.halo_ids.numel(), ld), dtype=torch.float32, device="cuda")
    ext[: g.n] = full[g.row0: g.row0 + g.n]
    g.exchange_halo(ext)
    assert torch.equal(ext[g.rows_per:], full[g.halo_ids])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        p_sh = cna.tl.association(sh, **kw)
        c_sh, f_sh = data.obs["coef"].to_numpy().copy(), data.obs["coef_fdr"].to_numpy().copy()
        p_e2e = cna.tl.association(shard_to_device(data, resident=False), **kw)  # one-shot shards (bench e2e)
        c_e2e = data.obs["coef"].to_numpy().copy()
        p_1 = cna.tl.association(cna.tl.to_device(data), **kw)
        c_1, f_1 = data.obs["coef"].to_numpy(), data.obs["coef_fdr"].to_numpy()
    assert p_sh == p_1 == p_e2e
    np.testing.assert_array_equal(np.isnan(c_sh), np.isnan(c_1))
    np.testing.assert_allclose(c_sh, c_1, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(c_e2e, c_1, rtol=1e-6, atol=1e-9)  # another cell order: column sums differ in the last bits
    np.testing.assert_allclose(f_sh, f_1, rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL, one rank per GPU)")
def test_nccl_two_gpus_halo_exchange_and_association():
    _spawn_nccl(_nccl_halo_and_large, 2)


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL, one rank per GPU)")
def test_nccl_two_gpus_golden_case():
    spec = dict(cases.DEMO_CASES["case_male_batch"])
    spec.pop("np_seed", None)
    spec.setdefault("seed", 0)
    _spawn_nccl(_sharded_vs_single, 2, spec, True, True)
