"""ctypes binding of libmma_b200.so (include/mma_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails,
the product path raises.  (The CPU oracle under oracle/ is test infrastructure
and is never imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMA_B200_LIB") or os.path.join(_HERE, "libmma_b200.so")     # override: A/B of two builds (dev tools)

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_WORKSPACE = 0, -1, -2, -3, -4
MAX_AGGR, MAX_SCALER = 8, 8
K1_ARGS_LOCAL = 1        # flags of mmconv_aggregate_fwd / _bwd_dst: arg indices are CSR slots
AGGR_KINDS = {"sum": 0, "mean": 1, "min": 2, "max": 3, "var": 4, "std": 5}
SCALER_KINDS = {"identity": 0, "amplification": 1, "attenuation": 2, "linear": 3, "inverse_linear": 4}
NC_COMBINE = {"sum": 0, "mean": 1, "max": 2, "min": 3, "none": 4}
ACT_SIGMOID, ACT_RAW = 0, 1

_vp, _i64, _i32, _f32, _u64, _u32 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_uint64, C.c_uint32



class K1Args(C.Structure):
    """mma_k1_args_t (include/mma_b200.h): the versioned argument block of mmconv_aggregate_fwd_args / _bwd_dst_args."""
    _fields_ = [("struct_size", _u32), ("flags", _i32),
                ("rowptr", _vp), ("col", _vp), ("perm", _vp), ("edge_gid", _vp), ("E_total", _i64),
                ("row_map", _vp), ("rng_row", _vp), ("rng_row0", _i64), ("row_chunks", _vp), ("n_chunks", _i64),
                ("vrowptr", _vp), ("n_vrows", _i64), ("seg_tab", _vp), ("split_tab", _vp), ("n_split", _i64),
                ("seg_ws", _vp), ("n_rows", _i64), ("E", _i64),
                ("P", _vp), ("ldp", _i64), ("Q", _vp), ("ldq", _i64), ("R", _vp), ("ldr", _i64), ("keep", _vp), ("ldk", _i64),
                ("p_drop", _f32), ("reserved0", _i32), ("seed", _u64), ("seed_dev", _vp),
                ("T", _i32), ("F_in", _i32), ("A", _i32), ("S", _i32), ("aggr_kinds", _vp), ("scaler_kinds", _vp),
                ("scale_tab", _vp), ("tab_stride", _i64),
                ("Y", _vp), ("ldy", _i64), ("arg_min", _vp), ("arg_max", _vp), ("stat_mean", _vp), ("stat_var", _vp),
                ("gslot", _vp), ("G", _vp), ("ldg", _i64), ("dP", _vp), ("lddp", _i64),
                ("col0", _i32), ("ncols", _i32)]


_SIGS = {
    "mma_b200_version": ([], C.c_int),
    "mma_last_cuda_error": ([], C.c_char_p),
    "mma_csr_build_workspace_bytes": ([_i64, _i64, C.POINTER(C.c_size_t)], C.c_int),
    "mma_csr_build": ([_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, C.c_size_t, _vp], C.c_int),
    "mma_invert_perm": ([_vp, _i64, _vp, _vp], C.c_int),
    "mmconv_aggregate_fwd": ([_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp,
                              _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64,
                              _vp, _i64, _f32, _u64, _vp, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _i64,
                              _vp, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp], C.c_int),
    "mmconv_aggregate_bwd_dst": ([_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp,
                                  _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64,
                                  _vp, _i64, _f32, _u64, _vp, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _i64,
                                  _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp],
                                 C.c_int),
    "mmconv_aggregate_fwd_args": ([C.POINTER(K1Args), _vp], C.c_int),
    "mmconv_aggregate_bwd_dst_args": ([C.POINTER(K1Args), _vp], C.c_int),
    "mma_gather_rows": ([_vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp, _i64, _vp], C.c_int),
    "mma_segment_sum_rows": ([_vp, _vp, _vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp], C.c_int),
    "mma_nc_aggregate_fwd": ([_vp, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32,
                              _vp, _vp, _vp, _f32, _u64, _vp, _vp, _vp, _vp], C.c_int),
    "mma_nc_aggregate_bwd_dst": ([_vp, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32,
                                  _vp, _vp, _vp, _f32, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp], C.c_int),
    "mma_nc_aggregate_bwd_src": ([_vp, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32,
                                  _vp, _vp, _f32, _u64, _vp, _vp, _vp, _i64, _vp, _i64, _vp], C.c_int),
    "mma_dropout_keep_scale": ([_f32, _u64, _u32, _i64, _i32, _vp, _i64, _vp], C.c_int),
    "mma_dropout_keep_scale_rows": ([_vp, _vp, _vp, _i64, _i64, _i64, _f32, _u64, _i32, _vp, _i64, _vp], C.c_int),
    "mma_tf32_split": ([_vp, _vp, _vp, _i64, _vp], C.c_int),
    "mma_linear_tf32x3": ([_vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp, _i64, _i64, _i64, _i32, _vp, _i64,
                           _vp, _i64, _vp, _vp, _vp, _i64, _i32, _i32, _vp], C.c_int),
    "mma_wgrad_tf32x3": ([_vp, _i64, _i32, _vp, _i64, _i32, _vp, _i64, _i32, _i64, _vp, _i64, _vp, _i32, _i32, _vp],
                         C.c_int),
    "mma_reduce_slabs": ([_vp, _vp, _i64, _i64, _vp, _vp], C.c_int),
    "mma_reduce_slabs_segmented": ([_vp, _vp, _i64, _i64, _vp, _vp], C.c_int),
    "mma_small_gemm": ([_vp, _i64, _i32, _vp, _i64, _i32, _vp, _i64, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "mma_small_gemm_splits": ([_i32, _i32], C.c_int),
    "mma_compose_post_weight": ([_vp, _i32, _i32, _i32, _i32, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp], C.c_int),
    "mma_compose_post_wgrad": ([_vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _i64, _i32, _vp, _i64, _vp], C.c_int),
    "mma_peer_epoch_advance": ([_vp, _vp, _vp], C.c_int),
    "mma_bn_relu_fwd": ([_vp, _i64, _i64, _i32, _vp, _vp, _f32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _f32, _vp], C.c_int),
    "mma_bn_relu_bwd": ([_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _i32, _vp, _i64, _vp, _vp, _vp], C.c_int),
    "mma_peer_alloc": ([C.c_size_t, C.POINTER(_vp), _vp], C.c_int),
    "mma_peer_free": ([_vp], C.c_int),
    "mma_peer_open": ([_vp, C.POINTER(_vp)], C.c_int),
    "mma_peer_close": ([_vp], C.c_int),
    "mma_peer_enable_access": ([_i32], C.c_int),
    "mma_peer_copy": ([_vp, _vp, C.c_size_t, _vp], C.c_int),
    "mma_peer_copy_2d": ([_vp, C.c_size_t, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp], C.c_int),
    "mma_peer_wait": ([_vp, _vp, _i32, _i32, _u64, _vp, _vp], C.c_int),
    "mma_sum_slices": ([C.POINTER(_vp), _i32, _i64, _i32, _vp, _i64, _vp], C.c_int),
}
EXPORTS = tuple(_SIGS)

_lib = None


class MMAError(RuntimeError):
    pass


def lib():
    """Loads the library once.  Raises (never falls back) if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise MMAError(
                f"{LIB_PATH} not found: build it with `python -m mma_b200.build` "
                "(nvcc, sm_100a).  mma_b200 has no CPU or PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGS.items():
            fn = getattr(l, name)          # AttributeError if a declared symbol is not exported
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = l
    return _lib


_ERR = {ERR_INVALID: "invalid argument", ERR_CUDA: "CUDA error", ERR_UNSUPPORTED: "unsupported size/limit",
        ERR_WORKSPACE: "workspace too small"}


def check(rc: int, what: str) -> None:
    if rc != OK:
        extra = ""
        if rc == ERR_CUDA:
            extra = ": " + lib().mma_last_cuda_error().decode()
        raise MMAError(f"{what} failed: {_ERR.get(rc, rc)}{extra}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("mma_b200 kernels need CUDA tensors (there is no CPU fallback); got a "
                               f"{t.device} tensor")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    if dev is None:
        raise RuntimeError("no tensor given")
    return dev


def i32_array(values):
    return (C.c_int32 * len(values))(*values)


# ---------------------------------------------------------------------------
# launch accounting / per-kernel CUDA-event timing (used by bench.py)
# ---------------------------------------------------------------------------
import contextlib

LAUNCH_COUNTS = {}          # kernel-launching C-ABI call name -> number of launches
_TIMERS = None              # None = off; else dict name -> list of (start, end) cuda events


def reset_counters():
    LAUNCH_COUNTS.clear()


def enable_timing(on: bool = True):
    global _TIMERS
    _TIMERS = {} if on else None


def timing_summary():
    """name -> (launches, mean ms) from the recorded CUDA events (call after a synchronize)."""
    out = {}
    for name, evs in (_TIMERS or {}).items():
        ms = [a.elapsed_time(b) for a, b in evs]
        out[name] = (len(ms), sum(ms) / max(len(ms), 1))
    return out


@contextlib.contextmanager
def kernel_scope(name: str, device):
    """Wraps ONE kernel launch through the C ABI: counts it and, when timing is on, brackets it
    with CUDA events on the launching (current) stream."""
    LAUNCH_COUNTS[name] = LAUNCH_COUNTS.get(name, 0) + 1
    if _TIMERS is None:
        with torch.cuda.device(device):
            yield
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.device(device):
        st = torch.cuda.current_stream(device)
        a.record(st)
        yield
        b.record(st)
    _TIMERS.setdefault(name, []).append((a, b))
