"""ctypes binding of libdcvic_b200.so (the C ABI declared in include/dcvic_b200.h).

The product path has NO fallback: if the library is missing or a call returns a negative
status, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from .build import LIB_PATH

OK = 0
VQ_REUSE_PREP, VQ_FORCE_EXACT, VQ_FORCE_TENSOR = 1, 2, 4
VQ_STAGE_SEARCH_ONLY, VQ_STAGE_FINISH_ONLY, VQ_STAGE_PREP_ONLY, VQ_TWO_KERNELS = 8, 16, 64, 128
GC_PRECISE = 2

_lock = threading.Lock()
_lib = None

_p, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list EVERY symbol include/dcvic_b200.h declares
SIGNATURES = {
    "dcvic_version": (C.c_char_p, []),
    "dcvic_error_string": (C.c_char_p, [_i]),
    "dcvic_vq_path": (_i, [_i, _i, _i]),
    "dcvic_vq_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "dcvic_vq_forward": (_i, [_p, _p, _i, _i, _i, _i, _i, _f, _i, _p, _p, _p, _p, _p, _i, _p, _sz, _p]),
    "dcvic_vq_backward": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _i, _p, _p, _p]),
    "dcvic_codebook_gather": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "dcvic_onehot_nchw": (_i, [_p, _i, _i, _i, _p, _p]),
    "dcvic_token_decode": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "dcvic_token_decode_ex": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _i, _f, _p, _p, _p, _p, _p, _p]),
    "dcvic_token_ce_backward": (_i, [_p, _p, _p, _p, _i, _i, _i, _f, _f, _p, _p]),
    "dcvic_gc_workspace_bytes": (_sz, [_i64, _i64]),
    "dcvic_gc_forward": (_i, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _f, _f, _i, _p, _p, _p, _p, _sz, _p]),
    "dcvic_gc_forward_dual": (_i, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _f, _f, _p, _p, _p, _p, _p, _p,
                                   _sz, _p]),
    "dcvic_gc_backward": (_i, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _f, _f, _p, _p, _p, _p]),
    "dcvic_gc_build_indexes": (_i, [_p, _i64, _p, _i, _f, _p, _p]),
    "dcvic_gc_codec_step": (_i, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _p, _i, _f, _f, _p, _p, _p, _p, _p]),
    "dcvic_eb_workspace_bytes": (_sz, [_i, _i, _i]),
    "dcvic_eb_forward": (_i, [_p, _p, C.POINTER(_p), _i, _i, _i, _f, _i, _p, _p, _p, _p, _sz, _p]),
    "dcvic_eb_backward": (_i, [_p, _p, _p, C.POINTER(_p), _i, _i, _i, _f, _p, C.POINTER(_p), _p, _sz, _p]),
    "dcvic_rate_workspace_bytes": (_sz, [_i64, _i64]),
    "dcvic_rate_bits": (_i, [_p, _i64, _i64, _p, _p, _sz, _p]),
    "dcvic_rate_bits_backward": (_i, [_p, _p, _i64, _i64, _p, _p]),
    "dcvic_ste_round": (_i, [_p, _i64, _p, _p]),
    "dcvic_pmf_to_quantized_cdf": (_i, [C.POINTER(C.c_float), _i, _i, C.POINTER(C.c_int32)]),
    "dcvic_pmf_to_quantized_cdf_rows": (_i, [_p, _i, _i, _p, _p, _i, _p, _p, _p]),
    "dcvic_tile_gather": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _i, _i, _p, _p]),
    "dcvic_tile_stitch": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _p, _i, _i, _p]),
    "dcvic_rans_encode": (_i, [_p, _p, _p, _i, _p, _i, _i, _p, _p, _p, _p, _p, _p]),
    "dcvic_rans_decode": (_i, [_p, _i64, _p, _p, _i64, _p, _i, _i, _p, _p, _p, _p]),
}


def load(path: str = LIB_PATH):
    """Load the library (once).  Raises RuntimeError when it has not been built.
    DCVIC_B200_LIB overrides the path (used by tools/trace_run.py to load the -DDCVIC_TRACE build)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("DCVIC_B200_LIB", path)
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} not found: the CUDA extension is not built (run `python -m dc_vic_b200.build`). "
                "dc_vic_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
        return lib


def check(rc: int, what: str) -> None:
    if rc != OK:
        msg = load().dcvic_error_string(int(rc)).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def cur_stream():
    """Raw handle of torch's current stream on the current device (no Stream object: this is on the per-call path of
    the small in-model invocations, where host overhead is all there is)."""
    import torch
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


class on_device:
    """``torch.cuda.device(dev)`` that does nothing when `dev` is already current (two cudaSetDevice calls saved)."""

    def __init__(self, device):
        import torch
        self._idx = device.index if device.index is not None else torch.cuda.current_device()
        self._guard = None

    def __enter__(self):
        import torch
        if torch.cuda.current_device() != self._idx:
            self._guard = torch.cuda.device(self._idx)
            self._guard.__enter__()
        return self

    def __exit__(self, *exc):
        if self._guard is not None:
            self._guard.__exit__(*exc)
        return False


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dc_vic_b200 runs on CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor")
