"""Writes tests/golden/rans_kat.npz: a fixed symbol sequence, its CDF tables and the stream oracle/rans_oracle.c
produces for it (known-answer vector: guards the oracle - and through it the GPU coder - against silent changes).
    python tests/golden/make_rans_kat.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import entropy_oracle as EO  # noqa: E402
from oracle import rans_oracle as R  # noqa: E402

rng = np.random.default_rng(2024)
rows, width = 4, 12
cdf = np.zeros((rows, width), dtype=np.int32)
lengths, offsets = [], []
for r in range(rows):
    k = 3 + 2 * r
    pmf = rng.random(k).astype(np.float32) + 0.01
    pmf /= pmf.sum() * 1.02
    row = EO.pmf_to_quantized_cdf(np.concatenate([pmf, [max(1.0 - pmf.sum(), 1e-6)]]).astype(np.float32), 16)
    cdf[r, : len(row)] = np.asarray(row)
    lengths.append(len(row))
    offsets.append(-(k // 2))
indexes = rng.integers(0, rows, 600).astype(np.int32)
symbols = np.array([rng.integers(offsets[i] - 20, offsets[i] + lengths[i] + 20) if j % 17 == 0
                    else rng.integers(offsets[i], offsets[i] + lengths[i] - 2) for j, i in enumerate(indexes)], dtype=np.int32)
stream = np.frombuffer(R.encode_with_indexes(symbols, indexes, cdf, lengths, offsets), dtype=np.uint8)
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rans_kat.npz"), symbols=symbols, indexes=indexes,
         cdf=cdf, lengths=np.array(lengths, dtype=np.int32), offsets=np.array(offsets, dtype=np.int32), stream=stream)
print("wrote rans_kat.npz:", stream.size, "bytes of stream for", symbols.size, "symbols")
