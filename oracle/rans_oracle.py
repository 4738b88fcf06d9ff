"""TEST INFRASTRUCTURE (oracle): ctypes wrapper of oracle/rans_oracle.c, the CPU restatement of CompressAI 1.2.4's
rANS coder (`compressai.ans`: rans_interface.cpp over ryg_rans' rans64.h; un-vendored dependency of the reference,
PARITY UNPINNED - see the header of the C file), plus the reference's wire format helpers restated from
src/utils/codec_utils.py:16-66.  Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rans_oracle.c")
LIB = os.path.join(HERE, "_build", "librans_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, SRC], check=True)
    return LIB


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.rans_oracle_encode.restype = C.c_long
        lib.rans_oracle_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_long]
        lib.rans_oracle_decode.restype = None
        lib.rans_oracle_decode.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _table(cdfs) -> np.ndarray:
    if isinstance(cdfs, np.ndarray) and cdfs.ndim == 2:
        return _i32(cdfs)
    width = max(len(r) for r in cdfs)
    t = np.zeros((len(cdfs), width), dtype=np.int32)
    for i, r in enumerate(cdfs):
        t[i, : len(r)] = np.asarray(r, dtype=np.int32)
    return t


def encode_with_indexes(symbols, indexes, cdfs, cdf_lengths, offsets) -> bytes:
    """compressai.ans.RansEncoder().encode_with_indexes(...)."""
    sym, idx = _i32(symbols).reshape(-1), _i32(indexes).reshape(-1)
    tab, ln, off = _table(cdfs), _i32(cdf_lengths).reshape(-1), _i32(offsets).reshape(-1)
    cap = 4 * sym.size + 16
    out = np.empty(cap, dtype=np.uint32)
    n = _load().rans_oracle_encode(sym.ctypes.data, idx.ctypes.data, sym.size, tab.ctypes.data, tab.shape[1],
                                   ln.ctypes.data, off.ctypes.data, out.ctypes.data, cap)
    assert n >= 0
    return out[:n].tobytes()


class RansDecoder:
    """compressai.ans.RansDecoder: set_stream + decode_stream (state persists between calls)."""

    def set_stream(self, stream: bytes) -> None:
        self.words = np.frombuffer(stream, dtype=np.uint32).copy()
        self.state = np.zeros(4, dtype=np.uint64)

    def decode_stream(self, indexes, cdfs, cdf_lengths, offsets) -> List[int]:
        idx = _i32(indexes).reshape(-1)
        tab, ln, off = _table(cdfs), _i32(cdf_lengths).reshape(-1), _i32(offsets).reshape(-1)
        out = np.empty(idx.size, dtype=np.int32)
        _load().rans_oracle_decode(self.words.ctypes.data, self.words.size, self.state.ctypes.data, idx.ctypes.data,
                                   idx.size, tab.ctypes.data, tab.shape[1], ln.ctypes.data, off.ctypes.data,
                                   out.ctypes.data)
        return out.tolist()

    def decode_with_indexes(self, stream: bytes, indexes, cdfs, cdf_lengths, offsets) -> List[int]:
        self.set_stream(stream)
        return self.decode_stream(indexes, cdfs, cdf_lengths, offsets)


# ---- wire format (src/utils/codec_utils.py:16-66)
def header_encode(img_size: Sequence[int], max_abs_y_hat: int, quality_ind: int) -> bytes:
    """HeaderHandler.encode: uint16 H, uint16 W, uint8 max|y_hat|, uint8 quality index = 6 bytes."""
    return (np.array(list(img_size), dtype=np.uint16).tobytes() + np.array(max_abs_y_hat, dtype=np.uint8).tobytes()
            + np.array(quality_ind, dtype=np.uint8).tobytes())


def header_decode(b: bytes) -> dict:
    hw = np.frombuffer(b[:4], dtype=np.uint16)
    return {"img_size": (int(hw[0]), int(hw[1])), "max_sample": int(np.frombuffer(b[4:5], dtype=np.uint8)[0]),
            "quality_ind": int(np.frombuffer(b[5:6], dtype=np.uint8)[0])}


def pack_strings(strings: Sequence[bytes]) -> bytes:
    """save_byte_strings: per string a uint32 length, then the payload."""
    return b"".join(np.array(len(s), dtype=np.uint32).tobytes() + s for s in strings)


def unpack_strings(blob: bytes) -> List[bytes]:
    out, p = [], 0
    while p < len(blob):
        n = int(np.frombuffer(blob[p:p + 4], dtype=np.uint32)[0])
        out.append(blob[p + 4:p + 4 + n])
        p += 4 + n
    return out
