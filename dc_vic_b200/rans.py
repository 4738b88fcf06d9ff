"""``compressai.ans`` surface (RansEncoder / RansDecoder / BufferedRansEncoder) for the compressai shim.

Call sites in the reference: ``EntropyModel.compress/decompress`` (CompressAI 1.2.4) and the CHARM decode loop
``minnen20_charm_context_model.py:175-202`` (``RansDecoder.set_stream`` / ``decode_stream`` per slice).
SURVEY 8(f) row 2.  The coder itself is CompressAI's ``rans64`` (ryg_rans, 64-bit state, 32-bit renormalisation,
one stream, symbols pushed in reverse): a sequential integer state machine.  Here it runs on the GPU through the C
ABI (``dcvic_rans_*``), one thread per independent stream, fed straight from the symbol / index tensors the entropy
kernels produced on the device - no ``.tolist()`` marshaling and no device->host hop of the latents.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import _lib

_ROW = "SURVEY 8(f) row 2 (CDF tables + rANS coder on GPU)"


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(f"dc_vic_b200.rans needs a CUDA device: the rANS coder of {_ROW} runs on the GPU "
                           "(no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _as_i32(x, device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.int32).contiguous()
    return torch.tensor(x, dtype=torch.int32, device=device)


class _Tables:
    """Quantized CDF tables on the device (flattened [rows, max_len] int32 + lengths + offsets)."""

    def __init__(self, cdfs, cdf_lengths, offsets, device):
        if isinstance(cdfs, torch.Tensor):
            self.cdf = cdfs.to(device=device, dtype=torch.int32).contiguous()
        else:
            width = max(len(r) for r in cdfs)
            t = torch.zeros(len(cdfs), width, dtype=torch.int32)
            for i, r in enumerate(cdfs):
                t[i, : len(r)] = torch.as_tensor(r, dtype=torch.int32)
            self.cdf = t.to(device)
        self.lengths = _as_i32(cdf_lengths, device)
        self.offsets = _as_i32(offsets, device)


_TABLE_CACHE = {}


def _tables_cached(cdfs, cdf_lengths, offsets, device) -> _Tables:
    """The CHARM decode loop passes the SAME Python list of CDF rows for every slice
    (minnen20_charm_context_model.py:175-177): upload it once."""
    if isinstance(cdfs, torch.Tensor):
        return _Tables(cdfs, cdf_lengths, offsets, device)
    key = (id(cdfs), id(cdf_lengths), id(offsets), str(device))
    hit = _TABLE_CACHE.get(key)
    if hit is None or hit[0] is not cdfs:
        _TABLE_CACHE.clear()
        hit = (cdfs, _Tables(cdfs, cdf_lengths, offsets, device))
        _TABLE_CACHE[key] = hit
    return hit[1]


def encode_with_indexes(symbols, indexes, cdfs, cdf_lengths, offsets) -> bytes:
    """One rANS stream over ``symbols`` (any shape, flattened in C order) -> the bytes CompressAI's
    ``RansEncoder.encode_with_indexes`` produces."""
    dev = _device()
    sym, idx = _as_i32(symbols, dev).reshape(-1), _as_i32(indexes, dev).reshape(-1)
    tab = _tables_cached(cdfs, cdf_lengths, offsets, dev)
    return encode_batch([sym], [idx], tab)[0]


def encode_batch(symbols: Sequence[torch.Tensor], indexes: Sequence[torch.Tensor], tab: _Tables) -> List[bytes]:
    """Independent streams (one per entry), encoded concurrently: one GPU thread per stream."""
    lib = _lib.load()
    dev = tab.cdf.device
    n = [int(s.numel()) for s in symbols]
    starts = torch.tensor([0] + list(torch.tensor(n).cumsum(0).tolist()), dtype=torch.int64, device=dev)
    sym = torch.cat([s.reshape(-1) for s in symbols]) if len(symbols) > 1 else symbols[0].reshape(-1)
    idx = torch.cat([s.reshape(-1) for s in indexes]) if len(indexes) > 1 else indexes[0].reshape(-1)
    S = len(n)
    # worst case 32 bits of payload per symbol is impossible at 16-bit precision; CompressAI's bypass coding can emit
    # several raw nibbles per out-of-range symbol, so budget 4 words per symbol + the final state
    cap_words = [4 * k + 8 for k in n]
    out_starts = torch.tensor([0] + list(torch.tensor(cap_words).cumsum(0).tolist()), dtype=torch.int64, device=dev)
    out = torch.empty(int(out_starts[-1]), dtype=torch.int32, device=dev)
    nwords = torch.zeros(S, dtype=torch.int32, device=dev)
    with _lib.on_device(dev):
        rc = lib.dcvic_rans_encode(_lib.ptr(sym), _lib.ptr(idx), _lib.ptr(starts), S, _lib.ptr(tab.cdf),
                                   tab.cdf.shape[0], tab.cdf.shape[1], _lib.ptr(tab.lengths), _lib.ptr(tab.offsets),
                                   _lib.ptr(out), _lib.ptr(out_starts), _lib.ptr(nwords), _lib.cur_stream())
        _lib.check(rc, "dcvic_rans_encode")
    nw = nwords.cpu().tolist()
    ends = out_starts.cpu().tolist()
    res = []
    for s in range(S):
        # the kernel writes the words of stream s backwards from the end of its slot (the encoder runs in reverse);
        # only the words that were written travel to the host
        end = int(ends[s + 1])
        res.append(out[end - nw[s]: end].cpu().numpy().tobytes())
    return res


def decode_with_indexes(stream: bytes, indexes, cdfs, cdf_lengths, offsets) -> torch.Tensor:
    dec = RansDecoder()
    dec.set_stream(stream)
    return dec.decode_stream_tensor(indexes, cdfs, cdf_lengths, offsets)


class RansEncoder:
    def encode_with_indexes(self, symbols, indexes, cdfs, cdf_lengths, offsets) -> bytes:
        return encode_with_indexes(symbols, indexes, cdfs, cdf_lengths, offsets)


class BufferedRansEncoder:
    """Collects (symbols, indexes) pairs; ``flush`` encodes them as ONE stream in push order."""

    def __init__(self):
        self._sym, self._idx, self._tab_args = [], [], None

    def encode_with_indexes(self, symbols, indexes, cdfs, cdf_lengths, offsets) -> None:
        dev = _device()
        self._sym.append(_as_i32(symbols, dev).reshape(-1))
        self._idx.append(_as_i32(indexes, dev).reshape(-1))
        self._tab_args = (cdfs, cdf_lengths, offsets)

    def flush(self) -> bytes:
        if not self._sym:
            return encode_with_indexes([], [], [[0, 65536]], [2], [0])
        sym, idx = torch.cat(self._sym), torch.cat(self._idx)
        out = encode_with_indexes(sym, idx, *self._tab_args)
        self._sym, self._idx = [], []
        return out


class RansDecoder:
    """Sequential decoder with a persistent (state, read position) on the device between ``decode_stream`` calls,
    as the CHARM loop needs (one call per slice, each depending on the previous slice's values)."""

    def __init__(self):
        self._words = None
        self._state = None

    def set_stream(self, stream: bytes) -> None:
        dev = _device()
        if len(stream) % 4 != 0 or len(stream) < 8:
            raise ValueError("a rans64 stream is a whole number (>= 2) of 32-bit words")
        self._words = torch.from_numpy(np.frombuffer(stream, dtype=np.int32).copy()).to(dev)
        self._state = torch.zeros(4, dtype=torch.int64, device=dev)      # [state, word position, initialised, -]

    def decode_stream_tensor(self, indexes, cdfs, cdf_lengths, offsets) -> torch.Tensor:
        if self._words is None:
            raise RuntimeError("RansDecoder.set_stream was not called")
        dev = self._words.device
        idx = _as_i32(indexes, dev).reshape(-1)
        tab = _tables_cached(cdfs, cdf_lengths, offsets, dev)
        out = torch.empty(idx.numel(), dtype=torch.int32, device=dev)
        lib = _lib.load()
        with _lib.on_device(dev):
            rc = lib.dcvic_rans_decode(_lib.ptr(self._words), self._words.numel(), _lib.ptr(self._state),
                                       _lib.ptr(idx), idx.numel(), _lib.ptr(tab.cdf), tab.cdf.shape[0],
                                       tab.cdf.shape[1], _lib.ptr(tab.lengths), _lib.ptr(tab.offsets), _lib.ptr(out),
                                       _lib.cur_stream())
            _lib.check(rc, "dcvic_rans_decode")
        return out

    def decode_stream(self, indexes, cdfs, cdf_lengths, offsets) -> List[int]:
        return self.decode_stream_tensor(indexes, cdfs, cdf_lengths, offsets).cpu().tolist()

    def decode_with_indexes(self, stream: bytes, indexes, cdfs, cdf_lengths, offsets) -> List[int]:
        self.set_stream(stream)
        return self.decode_stream(indexes, cdfs, cdf_lengths, offsets)
