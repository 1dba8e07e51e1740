"""ctypes binding of ``libcna_b200.so`` (the C-ABI declared in ``include/cna_b200.h``).

torch is used here only to own device memory and streams: every wrapper takes torch CUDA tensors,
checks dtype / contiguity, and passes raw device pointers plus the current CUDA stream to the
library.  There is no fallback: if the shared library is missing, or a tensor is not on a CUDA
device, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libcna_b200.so")

_lib = None


class CnaError(RuntimeError):
    pass


def load():
    """Load (building first if the sources are newer and nvcc is present) the shared library."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        try:
            _build.build()
        except Exception as exc:  # no nvcc on this machine and no prebuilt library
            raise ImportError(
                f"cna_b200: {LIB_PATH} is missing and could not be built ({exc}). "
                "Run `python -m cna_b200.build` on a machine with nvcc 12.9.") from exc
    lib = ctypes.CDLL(LIB_PATH)
    _declare(lib)
    if lib.cna_abi_version() != 5:
        raise ImportError("cna_b200: ABI version mismatch between _lib.py and libcna_b200.so")
    _lib = lib
    return lib


_I32P = ctypes.c_void_p
_VP = ctypes.c_void_p
_I64 = ctypes.c_int64
_INT = ctypes.c_int
_DBL = ctypes.c_double


class ResidArgs(ctypes.Structure):
    _fields_ = [
        ("s", _VP), ("ld_s", _I64), ("n_rows", _I64), ("inv_count", _VP),
        ("colmap", _VP), ("n", _INT),
        ("row_keep", _VP),
        ("C", _VP), ("Wt", _VP), ("r", _INT),
        ("seg_order", _VP), ("seg_off", _VP), ("n_batches", _INT),
        ("y", _VP),
        ("x_out", _VP), ("ld_x", _I64), ("kurt", _VP), ("ncorr", _VP), ("row_valid", _VP),
        ("x16_hi", _VP), ("x16_lo", _VP), ("ld16", _I64),
        ("qc_kurt", _VP), ("qc_median", _VP), ("qc_out", _VP),
    ]


# name -> argtypes; every function returns int except the three bookkeeping calls
_SIGNATURES = {
    "cna_graph_colsum": [_VP, _VP, _VP, _INT, _I64, _VP, _VP],
    "cna_graph_scale": [_VP, _VP, _VP, _INT, _I64, _VP, _DBL, _VP, _VP, _INT, _I64, _VP],
    "cna_diffuse_onehot": [_VP, _VP, _VP, _VP, _VP, _I64, _INT, _VP, _I64, _I64, _VP],
    "cna_diffuse_step_f32": [_VP, _VP, _VP, _VP, _VP, _VP, _I64, _INT, _I64, _I64, _VP],
    "cna_diffuse_step_f32_qc": [_VP, _VP, _VP, _VP, _VP, _VP, _I64, _INT, _I64, _I64, _VP, _VP, _VP, _INT, _VP, _VP],
    "cna_diffuse_step_f64": [_VP, _VP, _VP, _VP, _VP, _VP, _I64, _INT, _I64, _I64, _VP],
    "cna_diffuse_tile_limits": [_VP, _VP],
    "cna_diffuse_step_f32_tiled": [_VP, _VP, _VP, _VP, _VP, _I64, _I64, _INT, _I64, _I64, _VP, _VP, _VP, _INT, _INT, _VP],
    "cna_row_kurtosis": [_VP, _I64, _I64, _INT, _VP, _VP, _VP],
    "cna_batch_kurtosis": [_VP, _I64, _I64, _VP, _VP, _VP, _INT, _INT, _VP, _VP],
    "cna_resid_pass": [ctypes.POINTER(ResidArgs), _VP],
    "cna_qc_fixup": [_VP, _VP, _I64, _VP, _I64, _VP, _VP, _I64, _VP, _VP, _VP, _VP],
    "cna_gram": [_VP, _I64, _I64, _INT, _VP, _VP],
    "cna_gram_simt": [_VP, _I64, _I64, _INT, _VP, _VP],
    "cna_right_multiply": [_VP, _I64, _I64, _INT, _VP, _I64, _INT, _VP, _I64, _VP],
    "cna_perm_stats": [_VP, _VP, _I64, _INT, _VP, _VP, _INT, _VP, _INT, _VP, _INT, _VP, _VP, _VP,
                       _I64, _INT, _VP, _VP, _I64, _VP],
    "cna_perm_minp": [_VP, _VP, _I64, _VP, _INT, _INT, _INT, _VP, _VP, _VP, _VP],
    "cna_null_hist": [_VP, _I64, _I64, _INT, _VP, _I64, _INT, _VP, _INT, _DBL, _VP, _VP],
    "cna_obs_hist": [_VP, _VP, _I64, _VP, _VP, _INT, _VP, _VP, _VP],
    "cna_obs_hist_dev": [_VP, _VP, _I64, _VP, _VP, _INT, _VP, _VP, _VP, _VP],
    "cna_median_f64": [_VP, _VP, _I64, _VP, _VP, _I64, _VP],
    "cna_fdr_thresholds": [_VP, _INT, _VP, _VP, _VP, _VP],
    "cna_absmax": [_VP, _VP, _I64, _VP, _VP],
    "cna_cell_fdr": [_VP, _VP, _I64, _VP, _VP, _INT, _VP, _VP, _VP],
    "cna_cell_fdr_dev": [_VP, _VP, _I64, _VP, _VP, _INT, _VP, _VP, _VP, _VP],
    "cna_fdr_table": [_VP, _VP, _VP, _INT, _INT, _VP, _VP, _VP],
    "cna_knn_bruteforce": [_VP, _I64, _INT, _INT, _VP, _VP, _VP],
    "cna_knn_bruteforce_range": [_VP, _I64, _INT, _INT, _I64, _I64, _VP, _VP, _VP],
    "cna_bfs_expand": [_VP, _VP, _VP, _INT, _INT, _INT, _VP, _VP, _VP, _VP, _VP],
    "cna_bfs_keys": [_VP, _INT, _VP, _VP, _VP],
    "cna_bfs_place": [_VP, _INT, _VP, _VP, _VP],
    "cna_permute_csr": [_VP, _VP, _VP, _INT, _VP, _VP, _VP, _I64, _VP, _VP, _VP],
    "cna_host_randn": [_VP, _VP, _VP, _VP, _I64, _VP, _INT],
    "cna_host_perm_blocks": [_VP, _VP, _VP, _VP, _INT, _VP, _VP, _I64, _VP, _I64, _INT],
    "cna_host_perm_done": [_VP],
    "cna_host_perm_wait": [_VP],
    "cna_split_f16": [_VP, _I64, _I64, _INT, _INT, _VP, _VP, _I64, _I64, _VP],
    "cna_gram_tc": [_VP, _VP, _I64, _I64, _INT, _VP, _VP, _I64, _VP],
    "cna_sym_eig_top": [_VP, _I64, _INT, _INT, _VP, _VP, _VP, _VP, _I64, _VP],
    "cna_right_multiply_tc": [_VP, _VP, _I64, _I64, _INT, _VP, _VP, _I64, _INT, _VP, _I64, _VP],
    "cna_null_hist_tc": [_VP, _VP, _I64, _I64, _INT, _VP, _VP, _I64, _INT, _VP, _INT, _DBL, _VP, _VP],
    "cna_null_hist_tc_dev": [_VP, _VP, _I64, _I64, _INT, _VP, _VP, _I64, _INT, _VP, _INT, _VP, _DBL, _VP, _VP],
}
EXPORTS = sorted(list(_SIGNATURES) + ["cna_abi_version", "cna_last_error", "cna_launch_count",
                                      "cna_gram_tc_workspace", "cna_host_perm_blocks_async",
                                      "cna_median_workspace", "cna_sym_eig_workspace", "cna_tc_max_ctas",
                                      "cna_perm_draw_workspace", "cna_perm_draw_device", "cna_host_upload",
                                      "cna_host_upload_async", "cna_host_upload_wait"])



def _declare(lib):
    lib.cna_abi_version.restype = ctypes.c_int
    lib.cna_abi_version.argtypes = []
    lib.cna_last_error.restype = ctypes.c_char_p
    lib.cna_last_error.argtypes = []
    lib.cna_launch_count.restype = ctypes.c_int64
    lib.cna_launch_count.argtypes = []
    lib.cna_gram_tc_workspace.restype = ctypes.c_int64
    lib.cna_gram_tc_workspace.argtypes = [ctypes.c_int]
    lib.cna_median_workspace.restype = ctypes.c_int64
    lib.cna_median_workspace.argtypes = []
    lib.cna_perm_draw_workspace.restype = ctypes.c_int64
    lib.cna_perm_draw_workspace.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
    lib.cna_perm_draw_device.restype = ctypes.c_int
    lib.cna_perm_draw_device.argtypes = [_VP, _INT, _INT, _DBL, _INT, _VP, _VP, _I64, _VP, _I64, _VP, _VP, _VP,
                                         _VP, _I64, _VP]
    lib.cna_host_upload.restype = ctypes.c_int
    lib.cna_host_upload.argtypes = [_VP, _VP, _I64, _VP, _INT]
    lib.cna_host_upload_async.restype = ctypes.c_void_p
    lib.cna_host_upload_async.argtypes = [_VP, _VP, _I64, _VP, _INT]
    lib.cna_host_upload_wait.restype = ctypes.c_int
    lib.cna_host_upload_wait.argtypes = [_VP]
    lib.cna_tc_max_ctas.restype = ctypes.c_int
    lib.cna_tc_max_ctas.argtypes = [ctypes.c_int]
    lib.cna_sym_eig_workspace.restype = ctypes.c_int64
    lib.cna_sym_eig_workspace.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.cna_host_perm_blocks_async.restype = ctypes.c_void_p
    lib.cna_host_perm_blocks_async.argtypes = _SIGNATURES["cna_host_perm_blocks"]
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = args


_PROFILE = None  # name -> list of (start event, end event) while profiling is on


def profile_start():
    """Record a CUDA-event pair around every library call (on torch's current stream, the stream
    the kernels are launched on) until ``profile_stop``."""
    global _PROFILE
    _PROFILE = {}


def profile_stop():
    """Returns {entry point: (number of calls, total milliseconds)}."""
    global _PROFILE
    prof, _PROFILE = _PROFILE, None
    torch.cuda.synchronize()
    return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (prof or {}).items()}


def _call(name, *args):
    fn = getattr(load(), name)
    if _PROFILE is None:
        rc = fn(*args)
    else:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = fn(*args)
        b.record()
        _PROFILE.setdefault(name, []).append((a, b))
    if rc != 0:
        raise CnaError(f"{name} failed ({rc}): {load().cna_last_error().decode()}")


def _ptr(t, dtype, name, allow_none=False):
    if t is None:
        if allow_none:
            return None
        raise CnaError(f"{name}: tensor required")
    if not t.is_cuda:
        raise CnaError(f"{name}: expected a CUDA tensor (cna_b200 has no CPU path)")
    if t.dtype != dtype:
        raise CnaError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise CnaError(f"{name}: tensor must be contiguous")
    return t.data_ptr()


def _stream():
    # the raw handle of torch's current stream on the current device; torch.cuda.current_stream() builds a
    # Stream object behind several availability checks (~25 us a call, x ~20 calls per association())
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


_STREAM_OBJECTS = {}


def current_stream_object():
    """torch's current stream as a Stream object, cached by raw handle (torch.cuda.current_stream() costs
    ~15 us of Python per call; events are recorded ~10 times per association())."""
    dev = torch._C._cuda_getDevice()
    key = (dev, torch._C._cuda_getCurrentRawStream(dev))
    obj = _STREAM_OBJECTS.get(key)
    if obj is None:
        obj = _STREAM_OBJECTS[key] = torch.cuda.current_stream()
    return obj


def launch_count():
    return int(load().cna_launch_count())


# ---------------------------------------------------------------------------------------------
# thin typed wrappers
# ---------------------------------------------------------------------------------------------
def graph_colsum(indptr, indices, data, colsum):
    is64 = data.dtype == torch.float64
    _call("cna_graph_colsum", _ptr(indptr, torch.int32, "indptr"), _ptr(indices, torch.int32, "indices"),
                                   _ptr(data, data.dtype, "data"), int(is64), indptr.numel() - 1,
                                   _ptr(colsum, torch.float64, "colsum"), _stream())


def graph_scale(indptr, indices, data, colsum, self_weight, vals, diag, row_offset=0):
    is64 = data.dtype == torch.float64
    out64 = vals.dtype == torch.float64
    _call("cna_graph_scale", _ptr(indptr, torch.int32, "indptr"), _ptr(indices, torch.int32, "indices"),
                                  _ptr(data, data.dtype, "data"), int(is64), indptr.numel() - 1,
                                  _ptr(colsum, torch.float64, "colsum"), float(self_weight),
                                  _ptr(vals, vals.dtype, "vals"), _ptr(diag, vals.dtype, "diag"),
                                  int(out64), int(row_offset), _stream())


def diffuse_onehot(indptr, indices, vals, diag, code, n_samples, out, n_rows=None, row_offset=0):
    _call("cna_diffuse_onehot", _ptr(indptr, torch.int32, "indptr"), _ptr(indices, torch.int32, "indices"),
                                     _ptr(vals, torch.float32, "vals"), _ptr(diag, torch.float32, "diag"),
                                     _ptr(code, torch.int32, "code"), out.shape[0] if n_rows is None else int(n_rows),
                                     int(n_samples), _ptr(out, torch.float32, "out"), out.shape[1], int(row_offset),
                                     _stream())


def diffuse_step(indptr, indices, vals, diag, src, dst, n_cols, n_rows=None, row_offset=0):
    if src.dtype == torch.float32:
        name, dt = "cna_diffuse_step_f32", torch.float32
    else:
        name, dt = "cna_diffuse_step_f64", torch.float64
    _call(name, _ptr(indptr, torch.int32, "indptr"), _ptr(indices, torch.int32, "indices"),
          _ptr(vals, dt, "vals"), _ptr(diag, dt, "diag"), _ptr(src, dt, "src"), _ptr(dst, dt, "dst"),
          dst.shape[0] if n_rows is None else int(n_rows), int(n_cols), src.shape[1], int(row_offset), _stream())


def diffuse_step_qc(indptr, indices, vals, diag, src, dst, n_cols, col_batch, inv_count_ld, batch_inv, kurt,
                    n_rows=None, row_offset=0):
    """fp32 diffusion step that also writes the batch-kurtosis QC statistic of every finished row."""
    _call("cna_diffuse_step_f32_qc", _ptr(indptr, torch.int32, "indptr"), _ptr(indices, torch.int32, "indices"),
          _ptr(vals, torch.float32, "vals"), _ptr(diag, torch.float32, "diag"), _ptr(src, torch.float32, "src"),
          _ptr(dst, torch.float32, "dst"), dst.shape[0] if n_rows is None else int(n_rows), int(n_cols),
          src.shape[1], int(row_offset), _ptr(col_batch, torch.int8, "col_batch"),
          _ptr(inv_count_ld, torch.float64, "inv_count"), _ptr(batch_inv, torch.float64, "batch_inv"),
          batch_inv.numel(), _ptr(kurt, torch.float64, "kurt"), _stream())


def diffuse_tile_limits():
    """(output rows per tile, distinct source rows per tile) of the shared-memory-staged SpMM."""
    a, b = ctypes.c_int32(0), ctypes.c_int32(0)
    load().cna_diffuse_tile_limits(ctypes.addressof(a), ctypes.addressof(b))
    return a.value, b.value


def diffuse_step_tiled(indptr, plan, diag, src, dst, n_cols, n_rows=None, row_offset=0, stage_mode=0):
    """``plan``: TilePlan (tile_row, tile_u, usrc, epair) of the graph whose rows ``indptr`` describes."""
    _call("cna_diffuse_step_f32_tiled", _ptr(indptr, torch.int32, "indptr"), _ptr(plan.epair, torch.int32, "epair"),
          _ptr(diag, torch.float32, "diag"), _ptr(src, torch.float32, "src"), _ptr(dst, torch.float32, "dst"),
          dst.shape[0] if n_rows is None else int(n_rows), src.shape[0], int(n_cols), src.shape[1], int(row_offset),
          _ptr(plan.tile_row, torch.int32, "tile_row"), _ptr(plan.tile_u, torch.int32, "tile_u"),
          _ptr(plan.usrc, torch.int32, "usrc"), plan.n_tiles, int(stage_mode), _stream())


def row_kurtosis(s, n_samples, inv_count, kurt):
    _call("cna_row_kurtosis", _ptr(s, torch.float32, "s"), s.shape[1], s.shape[0], int(n_samples),
                                   _ptr(inv_count, torch.float64, "inv_count"),
                                   _ptr(kurt, torch.float64, "kurt"), _stream())


def batch_kurtosis(s, inv_count, seg_order, seg_off, kurt):
    nb = seg_off.numel() - 1
    _call("cna_batch_kurtosis", _ptr(s, torch.float32, "s"), s.shape[1], s.shape[0],
                                     _ptr(inv_count, torch.float64, "inv_count"),
                                     _ptr(seg_order, torch.int32, "seg_order"),
                                     _ptr(seg_off, torch.int32, "seg_off"), nb, seg_order.numel(),
                                     _ptr(kurt, torch.float64, "kurt"), _stream())


def resid_pass(s, inv_count, colmap, row_keep, C, Wt, seg_order, seg_off, y, x_out, kurt, ncorr, row_valid,
               planes=None, qc_kurt=None, qc_median=None, qc_out=None):
    """``row_keep`` (uint8 mask) or ``qc_kurt`` + ``qc_median`` (device: keep iff kurt < max(6, 2 median))
    decide which rows survive the QC; both None keeps every row.  ``qc_out``: receives the QC statistic of
    every row instead (the caller decides afterwards, ``qc_fixup``)."""
    a = ResidArgs()
    a.qc_out = _ptr(qc_out, torch.float64, "qc_out", allow_none=True)
    a.qc_kurt = _ptr(qc_kurt, torch.float64, "qc_kurt", allow_none=True) if row_keep is None else None
    a.qc_median = _ptr(qc_median, torch.float64, "qc_median", allow_none=True) if a.qc_kurt else None
    a.s = _ptr(s, torch.float32, "s"); a.ld_s = s.shape[1]; a.n_rows = s.shape[0]
    a.inv_count = _ptr(inv_count, torch.float64, "inv_count")
    a.colmap = _ptr(colmap, torch.int32, "colmap"); a.n = colmap.numel()
    a.row_keep = _ptr(row_keep, torch.uint8, "row_keep", allow_none=True)
    r = 0 if C is None else C.shape[1]
    a.C = _ptr(C, torch.float64, "C") if r else None
    a.Wt = _ptr(Wt, torch.float64, "Wt") if r else None
    a.r = r
    if seg_off is not None and seg_off.numel() > 2:
        a.seg_order = _ptr(seg_order, torch.int32, "seg_order")
        a.seg_off = _ptr(seg_off, torch.int32, "seg_off")
        a.n_batches = seg_off.numel() - 1
    else:
        a.seg_order = None; a.seg_off = None; a.n_batches = 1
    a.y = _ptr(y, torch.float64, "y")
    a.x_out = _ptr(x_out, torch.float32, "x_out", allow_none=True)
    a.ld_x = x_out.shape[1] if x_out is not None else 0
    if planes is not None:
        a.x16_hi = _ptr(planes.hi, torch.float16, "x16_hi"); a.x16_lo = _ptr(planes.lo, torch.float16, "x16_lo")
        a.ld16 = planes.ld
    else:
        a.x16_hi = None; a.x16_lo = None; a.ld16 = 0
    a.kurt = _ptr(kurt, torch.float64, "kurt", allow_none=True)
    a.ncorr = _ptr(ncorr, torch.float64, "ncorr")
    a.row_valid = _ptr(row_valid, torch.uint8, "row_valid")
    _call("cna_resid_pass", ctypes.byref(a), _stream())


def qc_fixup(qc, median, x, planes, kurt, ncorr, row_valid):
    """Blank the rows whose QC statistic fails ``qc < max(6, 2 median)`` (after ``resid_pass(qc_out=...)``)."""
    _call("cna_qc_fixup", _ptr(qc, torch.float64, "qc"), _ptr(median, torch.float64, "median"), qc.numel(),
          _ptr(x, torch.float32, "x", allow_none=True), 0 if x is None else x.shape[1],
          None if planes is None else _ptr(planes.hi, torch.float16, "x16_hi"),
          None if planes is None else _ptr(planes.lo, torch.float16, "x16_lo"),
          0 if planes is None else planes.ld, _ptr(kurt, torch.float64, "kurt", allow_none=True),
          _ptr(ncorr, torch.float64, "ncorr"), _ptr(row_valid, torch.uint8, "row_valid"), _stream())


def gram(x, n, out, simt=False):
    _call("cna_gram_simt" if simt else "cna_gram", _ptr(x, torch.float32, "x"), x.shape[1], x.shape[0],
          int(n), _ptr(out, torch.float64, "gram"), _stream())


def right_multiply(x, n, b, n_out, out):
    _call("cna_right_multiply", _ptr(x, torch.float32, "x"), x.shape[1], x.shape[0], int(n),
                                     _ptr(b, torch.float32, "b"), b.shape[1], int(n_out),
                                     _ptr(out, torch.float32, "out"), out.shape[1], _stream())


def perm_stats(y, perm, C, W, Ut, ks, ssered, ssefull, ycond, n_local, planes=None):
    """Ut / ks / ssefull may be None (conditioning only); ``planes`` = Planes [n_local x n] that
    receive the conditioned phenotypes transposed (zero-initialised by the caller)."""
    K, n = perm.shape
    r = 0 if C is None else C.shape[1]
    kmax = 0 if Ut is None else Ut.shape[0]
    _call("cna_perm_stats", _ptr(y, torch.float64, "y"), _ptr(perm, torch.int32, "perm"), K, n,
                                 _ptr(C, torch.float64, "C") if r else None,
                                 _ptr(W, torch.float64, "W") if r else None, r,
                                 _ptr(Ut, torch.float64, "Ut") if kmax else None, kmax,
                                 _ptr(ks, torch.int32, "ks") if kmax else None, ks.numel() if kmax else 0,
                                 _ptr(ssered, torch.float64, "ssered", allow_none=True),
                                 _ptr(ssefull, torch.float64, "ssefull", allow_none=True),
                                 _ptr(ycond, torch.float32, "ycond", allow_none=True),
                                 0 if ycond is None else ycond.shape[1], int(n_local),
                                 None if planes is None else _ptr(planes.hi, torch.float16, "yt_hi"),
                                 None if planes is None else _ptr(planes.lo, torch.float16, "yt_lo"),
                                 0 if planes is None else planes.ld, _stream())


def perm_minp(ssered, ssefull, ks, n, r, minp, argk, r2):
    _call("cna_perm_minp", _ptr(ssered, torch.float64, "ssered"), _ptr(ssefull, torch.float64, "ssefull"),
          ssered.numel(), _ptr(ks, torch.int32, "ks"), ks.numel(), int(n), int(r),
          _ptr(minp, torch.float64, "minp"), _ptr(argk, torch.int32, "argk"), _ptr(r2, torch.float64, "r2"),
          _stream())


def null_hist(x, n, ycond, n_null, edges, edge0, hist):
    _call("cna_null_hist", _ptr(x, torch.float32, "x"), x.shape[1], x.shape[0], int(n),
                                _ptr(ycond, torch.float32, "ycond"), ycond.shape[1], int(n_null),
                                _ptr(edges, torch.float64, "edges"), edges.numel(), float(edge0),
                                _ptr(hist, torch.int32, "hist"), _stream())


def obs_hist(ncorr, row_valid, edges, thresholds, rank_hist, det_hist):
    _call("cna_obs_hist", _ptr(ncorr, torch.float64, "ncorr"),
                               _ptr(row_valid, torch.uint8, "row_valid", allow_none=True), ncorr.numel(),
                               _ptr(edges, torch.float64, "edges"), _ptr(thresholds, torch.float64, "thresholds"),
                               edges.numel(), _ptr(rank_hist, torch.int32, "rank_hist"),
                               _ptr(det_hist, torch.int32, "det_hist"), _stream())


_MEDIAN_WS = {}
MEDIAN_SKIP = None  # float64 scalar whose bit pattern cna_median_f64 ignores (set on first use)


def median_skip_value():
    """The float64 (a signalling-NaN bit pattern) that marks entries ``median`` must ignore."""
    global MEDIAN_SKIP
    if MEDIAN_SKIP is None:
        import numpy as np
        MEDIAN_SKIP = np.array([0x7FF4DEADBEEF0001], dtype=np.uint64).view(np.float64)
    return MEDIAN_SKIP


def median(v, valid, out):
    """out[0] = np.median of the float64 device vector ``v`` restricted to ``valid`` (uint8 mask or None)
    and to entries that are not the skip pattern; NaN if any such entry is NaN or none exists.
    out[1] = number of entries considered.  Stays on the device: nothing is copied back."""
    need = int(load().cna_median_workspace())
    ws = _MEDIAN_WS.get(v.device)
    if ws is None:
        ws = _MEDIAN_WS[v.device] = torch.empty(need, dtype=torch.uint8, device=v.device)
    _call("cna_median_f64", _ptr(v, torch.float64, "v"), _ptr(valid, torch.uint8, "valid", allow_none=True),
          v.numel(), _ptr(out, torch.float64, "out"), ws.data_ptr(), need, _stream())


def fdr_thresholds(maxabs, thresholds, edges, count):
    """Thresholds / histogram edges of _association.py:101-102 + _stats.py:51 from the device-resident
    max |ncorr| (``maxabs`` [1] float64); ``count`` [1] int32 receives their number."""
    _call("cna_fdr_thresholds", _ptr(maxabs, torch.float64, "maxabs"), thresholds.numel(),
          _ptr(thresholds, torch.float64, "thresholds"), _ptr(edges, torch.float64, "edges"),
          _ptr(count, torch.int32, "count"), _stream())


def obs_hist_dev(ncorr, row_valid, edges, thresholds, count, rank_hist, det_hist):
    _call("cna_obs_hist_dev", _ptr(ncorr, torch.float64, "ncorr"),
          _ptr(row_valid, torch.uint8, "row_valid", allow_none=True), ncorr.numel(),
          _ptr(edges, torch.float64, "edges"), _ptr(thresholds, torch.float64, "thresholds"), edges.numel(),
          _ptr(count, torch.int32, "count"), _ptr(rank_hist, torch.int32, "rank_hist"),
          _ptr(det_hist, torch.int32, "det_hist"), _stream())


def null_hist_tc_dev(xp, n, ytp, n_null, edges, count, hist):
    """``null_hist_tc`` with the number of edges read from the device (``count`` [1] int32)."""
    _call("cna_null_hist_tc_dev", _ptr(xp.hi, torch.float16, "xh"), _ptr(xp.lo, torch.float16, "xl"), xp.ld,
          xp.rows, int(n), _ptr(ytp.hi, torch.float16, "yth"), _ptr(ytp.lo, torch.float16, "ytl"), ytp.ld,
          int(n_null), _ptr(edges, torch.float64, "edges"), edges.numel(), _ptr(count, torch.int32, "count"),
          0.0, _ptr(hist, torch.int64, "hist"), _stream())


def absmax(v, row_valid, out):
    _call("cna_absmax", _ptr(v, torch.float64, "v"), _ptr(row_valid, torch.uint8, "row_valid", allow_none=True),
                             v.numel(), _ptr(out, torch.float64, "out"), _stream())


def cell_fdr(ncorr, row_valid, thresholds, prefix_min_fdr, coef, fdr):
    _call("cna_cell_fdr", _ptr(ncorr, torch.float64, "ncorr"),
                               _ptr(row_valid, torch.uint8, "row_valid", allow_none=True), ncorr.numel(),
                               _ptr(thresholds, torch.float64, "thresholds"),
                               _ptr(prefix_min_fdr, torch.float64, "prefix_min_fdr"), thresholds.numel(),
                               _ptr(coef, torch.float64, "coef"), _ptr(fdr, torch.float64, "fdr"), _stream())


def cell_fdr_dev(ncorr, row_valid, thresholds, prefix_min_fdr, count, coef, fdr):
    _call("cna_cell_fdr_dev", _ptr(ncorr, torch.float64, "ncorr"),
          _ptr(row_valid, torch.uint8, "row_valid", allow_none=True), ncorr.numel(),
          _ptr(thresholds, torch.float64, "thresholds"), _ptr(prefix_min_fdr, torch.float64, "prefix_min_fdr"),
          thresholds.numel(), _ptr(count, torch.int32, "count"), _ptr(coef, torch.float64, "coef"),
          _ptr(fdr, torch.float64, "fdr"), _stream())


def fdr_table(null_hist, rank_hist, count, n_null, fdr, prefix_min_fdr):
    """fdr / running-min-fdr per threshold from the device-resident histograms (int64 null histogram
    summed over the nulls, int32 observed rank histogram)."""
    _call("cna_fdr_table", _ptr(null_hist, torch.int64, "null_hist"), _ptr(rank_hist, torch.int32, "rank_hist"),
          _ptr(count, torch.int32, "count"), fdr.numel(), int(n_null), _ptr(fdr, torch.float64, "fdr"),
          _ptr(prefix_min_fdr, torch.float64, "prefix_min_fdr"), _stream())


def bfs_expand(indptr, indices, frontier, pos_base, next_level, level, first_parent, nxt, next_count):
    _call("cna_bfs_expand", _ptr(indptr, torch.int32, "indptr"), _ptr(indices, torch.int32, "indices"),
          _ptr(frontier, torch.int32, "frontier"), frontier.numel(), int(pos_base), int(next_level),
          _ptr(level, torch.int32, "level"), _ptr(first_parent, torch.int32, "first_parent"),
          _ptr(nxt, torch.int32, "next"), _ptr(next_count, torch.int32, "next_count"), _stream())


def bfs_keys(nxt, n, first_parent, keys):
    _call("cna_bfs_keys", _ptr(nxt, torch.int32, "next"), int(n), _ptr(first_parent, torch.int32, "first_parent"),
          _ptr(keys, torch.int64, "keys"), _stream())


def bfs_place(sorted_keys, n, order, offset, frontier):
    _call("cna_bfs_place", _ptr(sorted_keys, torch.int64, "sorted_keys"), int(n),
          _ptr(order, torch.int64, "order") + 8 * int(offset), _ptr(frontier, torch.int32, "frontier"), _stream())


def permute_csr(indptr, indices, data, order, inv, new_indptr, new_indices, new_data):
    _call("cna_permute_csr", _ptr(indptr, torch.int32, "indptr"), _ptr(indices, torch.int32, "indices"),
          _ptr(data, data.dtype, "data"), int(data.dtype == torch.float64), _ptr(order, torch.int64, "order"),
          _ptr(inv, torch.int32, "inv"), _ptr(new_indptr, torch.int32, "new_indptr"), order.numel(),
          _ptr(new_indices, torch.int32, "new_indices"), _ptr(new_data, data.dtype, "new_data"), _stream())


def round_up(x, m):
    return (x + m - 1) // m * m


class Planes:
    """fp16 hi/lo planes of an fp32 matrix (x = hi + lo to 2^-22): the operand format of the
    tcgen05 kernels.  ``t`` is one [2, rows, ld16] tensor (plane 0 = hi, 1 = lo)."""

    def __init__(self, rows, cols, device, zero=False):
        self.rows, self.cols = int(rows), int(cols)
        self.ld = round_up(max(self.cols, 1), 16)
        alloc = torch.zeros if zero else torch.empty
        self.t = alloc((2, self.rows, self.ld), dtype=torch.float16, device=device)

    @property
    def hi(self):
        return self.t[0]

    @property
    def lo(self):
        return self.t[1]


def split_f16(src, n_cols, transpose=False, planes=None):
    """src [rows x ld] fp32 -> Planes of src[:, :n_cols] (or of its transpose)."""
    rows = src.shape[0]
    if planes is None:
        planes = Planes(n_cols, rows, src.device) if transpose else Planes(rows, n_cols, src.device)
    _call("cna_split_f16", _ptr(src, torch.float32, "src"), src.shape[1], rows, int(n_cols), int(transpose),
          _ptr(planes.hi, torch.float16, "hi"), _ptr(planes.lo, torch.float16, "lo"), planes.ld, planes.rows,
          _stream())
    return planes


_GRAM_WS = {}


def gram_tc(xp, n, out):
    """out[n x n] (fp64) += X^T X on the tensor cores; xp = Planes of X."""
    need = int(load().cna_gram_tc_workspace(int(n)))
    if need < 0:
        raise CnaError(f"cna_gram_tc: n={n} not supported (n <= 512)")
    key = (xp.t.device, need)
    ws = _GRAM_WS.get(key)
    if ws is None:
        _GRAM_WS.clear()
        ws = _GRAM_WS[key] = torch.empty(need, dtype=torch.uint8, device=xp.t.device)
    _call("cna_gram_tc", _ptr(xp.hi, torch.float16, "xh"), _ptr(xp.lo, torch.float16, "xl"), xp.ld, xp.rows,
          int(n), _ptr(out, torch.float64, "gram"), ws.data_ptr(), need, _stream())


def tc_max_ctas(cap):
    """Cap the grid of the persistent tensor-core kernels (0 = one CTA per SM); returns the previous cap."""
    return int(load().cna_tc_max_ctas(int(cap)))


_EIG_WS = {}


def sym_eig_top(G, k, w_out, ut_out, de_out=None):
    """Leading k eigenpairs of the symmetric [n x n] fp64 device matrix G: w_out[k] descending,
    ut_out[k x n] (row c = eigenvector of the c-th largest eigenvalue).  One launch, no host round trip."""
    n = int(G.shape[0])
    need = int(load().cna_sym_eig_workspace(n, int(k)))
    key = (G.device, need)
    ws = _EIG_WS.get(key)
    if ws is None:
        _EIG_WS.clear()
        ws = _EIG_WS[key] = torch.empty(need, dtype=torch.uint8, device=G.device)
    _call("cna_sym_eig_top", _ptr(G, torch.float64, "G"), int(G.stride(0)), n, int(k),
          _ptr(w_out, torch.float64, "w_out"), _ptr(ut_out, torch.float64, "ut_out"),
          _ptr(de_out, torch.float64, "de_out", allow_none=True), ws.data_ptr(), need, _stream())


def right_multiply_tc(xp, n, btp, n_out, out):
    """out[rows x ld_out] (fp32) = X . B; xp = Planes of X, btp = Planes of B^T [n_out x n]."""
    _call("cna_right_multiply_tc", _ptr(xp.hi, torch.float16, "xh"), _ptr(xp.lo, torch.float16, "xl"), xp.ld,
          xp.rows, int(n), _ptr(btp.hi, torch.float16, "bth"), _ptr(btp.lo, torch.float16, "btl"), btp.ld,
          int(n_out), _ptr(out, torch.float32, "out"), out.shape[1], _stream())


def null_hist_tc(xp, n, ytp, n_null, edges, edge0, hist):
    """hist: int64 [n_edges], the threshold histogram summed over the n_null columns."""
    _call("cna_null_hist_tc", _ptr(xp.hi, torch.float16, "xh"), _ptr(xp.lo, torch.float16, "xl"), xp.ld,
          xp.rows, int(n), _ptr(ytp.hi, torch.float16, "yth"), _ptr(ytp.lo, torch.float16, "ytl"), ytp.ld,
          int(n_null), _ptr(edges, torch.float64, "edges"), edges.numel(), float(edge0),
          _ptr(hist, torch.int64, "hist"), _stream())


# ---------------------------------------------------------------------------------------------
# host-side permutation drawing on numpy's legacy global generator
# ---------------------------------------------------------------------------------------------
class _LegacyState:
    """np.random's global RandomState unpacked for the C-ABI, written back on exit."""

    def __enter__(self):
        import numpy as np
        kind, key, pos, has_gauss, gauss = np.random.get_state()
        if kind != "MT19937":
            raise CnaError(f"numpy's global generator is {kind}, expected MT19937")
        self.key = np.array(key, dtype=np.uint32)
        self.pos = ctypes.c_int(int(pos))
        self.has_gauss = ctypes.c_int(int(has_gauss))
        self.gauss = ctypes.c_double(float(gauss))
        return self

    def args(self):
        return (self.key.ctypes.data, ctypes.addressof(self.pos), ctypes.addressof(self.has_gauss),
                ctypes.addressof(self.gauss))

    def __exit__(self, *exc):
        import numpy as np
        np.random.set_state(("MT19937", self.key, self.pos.value, self.has_gauss.value, self.gauss.value))


def _host_call(name, *args):
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise CnaError(f"{name} failed ({rc}): {load().cna_last_error().decode()}")


def host_randn(count, n_threads=0):
    """``np.random.randn(count)`` from (and advancing) numpy's legacy global generator."""
    import numpy as np
    out = np.empty(int(count), dtype=np.float64)
    with _LegacyState() as st:
        _host_call("cna_host_randn", *st.args(), int(count), out.ctypes.data, int(n_threads))
    return out


def host_perm_blocks(block_off, src_pos, num, n_threads=0):
    """int32 [num x total_rows] permutation matrix, see cna_host_perm_blocks in the header."""
    import numpy as np
    block_off = np.ascontiguousarray(block_off, dtype=np.int32)
    total = int(block_off[-1])
    out = np.empty((int(num), total), dtype=np.int32)
    sp = None if src_pos is None else np.ascontiguousarray(src_pos, dtype=np.int32)
    with _LegacyState() as st:
        _host_call("cna_host_perm_blocks", *st.args(), len(block_off) - 1, block_off.ctypes.data,
                   None if sp is None else sp.ctypes.data, int(num), out.ctypes.data, total, int(n_threads))
    return out


class HostPermJob:
    """``host_perm_blocks`` running on a library-owned thread.  numpy's global generator must not
    be touched between construction and ``result()``, which writes the advanced state back."""

    def __init__(self, block_off, src_pos, num, n_threads=0):
        import numpy as np
        self.block_off = np.ascontiguousarray(block_off, dtype=np.int32)
        self.src_pos = None if src_pos is None else np.ascontiguousarray(src_pos, dtype=np.int32)
        total = int(self.block_off[-1])
        # page-locked when a GPU is present: the index matrix goes straight to the device afterwards
        self.out_t = torch.empty((int(num), total), dtype=torch.int32, pin_memory=torch.cuda.is_available())
        self.out = self.out_t.numpy()
        self.state = _LegacyState().__enter__()
        self.handle = load().cna_host_perm_blocks_async(
            *self.state.args(), len(self.block_off) - 1, self.block_off.ctypes.data,
            None if self.src_pos is None else self.src_pos.ctypes.data, int(num), self.out.ctypes.data, total,
            int(n_threads))

    def done(self):
        return self.handle is None or bool(load().cna_host_perm_done(self.handle))

    def result(self):
        if self.handle is not None:
            rc = load().cna_host_perm_wait(self.handle)
            self.handle = None
            self.state.__exit__(None, None, None)
            if rc != 0:
                raise CnaError(f"cna_host_perm_blocks failed ({rc}): {load().cna_last_error().decode()}")
        return self.out

    def result_tensor(self):
        """The (pinned) host tensor behind ``result()``."""
        self.result()
        return self.out_t

    def __del__(self):
        # the library thread writes into buffers owned by this object: never let them be freed
        # (an exception unwinding the caller, interpreter shutdown) before it has finished
        try:
            if getattr(self, "handle", None) is not None:
                self.result()
        except Exception:
            pass


class HostUpload:
    """A numpy array on its way to the device through ``cna_host_upload`` (staged by a few host threads when
    the buffer is pageable).  ``background=True`` runs the upload on a library thread: the array must stay
    untouched until ``wait()``; the destination tensor may be queued on right away (the copies are ordered
    on the stream the upload was started on — callers on other streams wait for ``wait()`` first)."""

    def __init__(self, arr, device, background=False, n_threads=0):
        import numpy as np
        self.arr = np.ascontiguousarray(arr)
        self.tensor = torch.empty(self.arr.shape, dtype=getattr(torch, self.arr.dtype.name), device=device)
        self.handle = None
        args = (self.tensor.data_ptr(), self.arr.ctypes.data, self.arr.nbytes, _stream(), int(n_threads))
        if background:
            self.handle = load().cna_host_upload_async(*args)
            if not self.handle:
                raise CnaError(f"cna_host_upload_async failed: {load().cna_last_error().decode()}")
        else:
            rc = load().cna_host_upload(*args)
            if rc != 0:
                raise CnaError(f"cna_host_upload failed ({rc}): {load().cna_last_error().decode()}")

    def wait(self):
        if self.handle is not None:
            rc = load().cna_host_upload_wait(self.handle)
            self.handle = None
            if rc != 0:
                raise CnaError(f"cna_host_upload failed ({rc}): {load().cna_last_error().decode()}")
        return self.tensor

    def __del__(self):
        try:
            self.wait()
        except Exception:  # noqa: BLE001
            pass


_DRAW_WS = {}
_DRAW_STREAM = {}
_DRAW_RING = {}


class DevicePermJob:
    """``host_perm_blocks`` on the GPU (``cna_perm_draw_device``), queued on a side stream of its own so that
    it runs beside whatever the caller launches next.  Same contract as ``HostPermJob``: numpy's global
    generator must not be touched between construction and ``result()`` / ``finish()``, which writes the
    advanced state back.  ``finish()`` returns False when the draw has to be repeated with the host engine
    (an argsort within rounding of a tie, or a stream that was too short: probability ~1e-6 per call); the
    generator is then left in its state from before the draw."""

    _turn = 0

    def __init__(self, block_off, src_pos, num, device):
        import numpy as np
        kind, key, pos, has_gauss, gauss = np.random.get_state()
        if kind != "MT19937":
            raise CnaError(f"numpy's global generator is {kind}, expected MT19937")
        self.before = (kind, key, pos, has_gauss, gauss)
        self.block_off = np.ascontiguousarray(block_off, dtype=np.int32)
        key = np.ascontiguousarray(key, dtype=np.uint32)
        n, nb = int(self.block_off[-1]), len(self.block_off) - 1
        self.count, self.first = n * int(num), 1 if has_gauss else 0
        dev = torch.device(device)
        need = int(load().cna_perm_draw_workspace(n, int(num), nb))
        ws = _DRAW_WS.get((dev, need))
        if ws is None:
            _DRAW_WS.clear()
            ws = _DRAW_WS[(dev, need)] = torch.empty(need, dtype=torch.uint8, device=dev)
        # Output buffers come from a ring of two per shape, allocated once: the draw runs on a stream of its
        # own, so nothing here may be ordered by (or wait for) the caller's stream.  A buffer is reused two
        # draws later, when the call that consumed it has long returned.
        ring = _DRAW_RING.setdefault((dev, int(num), n), [])
        if len(ring) < 2:
            ring.append((torch.empty((int(num), n), dtype=torch.int32, device=dev),
                         torch.empty(650, dtype=torch.int32, device=dev),  # state (626) | pad | tail (4 f64) | flag
                         torch.empty(650, dtype=torch.int32, pin_memory=True)))
            torch.cuda.current_stream(dev).synchronize()  # first use only: the allocations are settled
        self.out, self.small, self.host = ring[DevicePermJob._turn % 2] if len(ring) == 2 else ring[0]
        DevicePermJob._turn += 1
        src = None if src_pos is None else np.ascontiguousarray(src_pos, dtype=np.int32)
        self._keep = (ws, key, src)
        side = _DRAW_STREAM.get(dev)
        if side is None:
            # high priority: its single-CTA recurrence should start as soon as it is queued, beside the
            # diffusion, instead of waiting for a free slot behind the grid of the running kernel
            side = _DRAW_STREAM[dev] = torch.cuda.Stream(device=dev, priority=-1)
        with torch.cuda.stream(side):
            base = self.small.data_ptr()
            rc = load().cna_perm_draw_device(key.ctypes.data, int(pos), int(has_gauss), float(gauss), nb,
                                             self.block_off.ctypes.data, None if src is None else src.ctypes.data,
                                             int(num), self.out.data_ptr(), n, base, base + 640 * 4, base + 648 * 4,
                                             ws.data_ptr(), need, side.cuda_stream)
            if rc != 0:
                raise CnaError(f"cna_perm_draw_device failed ({rc}): {load().cna_last_error().decode()}")
            self.host.copy_(self.small, non_blocking=True)
            self.event = torch.cuda.Event()
            self.event.record(side)
        self.finished = None

    def done(self):
        return True  # queued: whatever consumes the permutations waits on the device

    def result_tensor_device(self):
        """The [num x n] int32 index matrix on the device; the current stream waits for the draw."""
        current_stream_object().wait_event(self.event)
        return self.out

    def finish(self):
        """Waits for the draw, hands numpy's generator its advanced state; False = repeat on the host."""
        import math

        import numpy as np
        if self.finished is not None:
            return self.finished
        self.event.synchronize()
        h = self.host.numpy()
        state = h[:626].view(np.uint32)
        tail = h[640:648].view(np.float64)
        ok = bool(state[625]) and int(h[648]) == 0
        if ok:
            has_gauss, gauss = 0, 0.0
            if (self.count - self.first) % 2 == 1:  # legacy_gauss caches the second deviate of the last pair
                r2, x1 = float(tail[1]), float(tail[2])
                has_gauss, gauss = 1, math.sqrt(-2.0 * math.log(r2) / r2) * x1
            np.random.set_state(("MT19937", state[:624].copy(), int(state[624]), has_gauss, gauss))
        else:
            np.random.set_state(self.before)
        self.finished = ok
        return ok

    def result(self):
        ok = self.finish()
        if not ok:
            raise CnaError("device permutation draw must be repeated on the host")
        return self.out.cpu().numpy()


def knn_bruteforce(points, k, queries=None):
    """points: [n, dim] float32 CUDA tensor.  Returns (idx int64 [nq, k], dist2 float32 [nq, k]) for the
    queries ``queries`` = (q0, q1) (default: all points)."""
    n, dim = points.shape
    q0, q1 = (0, n) if queries is None else queries
    pad = next(d for d in (4, 8, 16, 32, 64) if d >= dim) if dim <= 64 else None
    if pad is None:
        raise CnaError("knn_bruteforce: at most 64 dimensions")
    if pad != dim:
        points = torch.nn.functional.pad(points, (0, pad - dim))
    points = points.contiguous()
    idx = torch.empty((q1 - q0, k), dtype=torch.int32, device=points.device)
    d2 = torch.empty((q1 - q0, k), dtype=torch.float32, device=points.device)
    _call("cna_knn_bruteforce_range", _ptr(points, torch.float32, "points"), n, pad, int(k), int(q0), int(q1 - q0),
          _ptr(idx, torch.int32, "idx"), _ptr(d2, torch.float32, "dist2"), _stream())
    return idx.long(), d2
