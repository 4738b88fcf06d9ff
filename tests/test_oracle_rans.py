"""CPU: the rANS oracle (oracle/rans_oracle.c) and the wire format.  CompressAI's coder is not available offline
(parity unpinned at that boundary): the oracle is pinned by construction (round trips, coder invariants, a committed
known-answer stream) and the wire format against the reference's own src/utils/codec_utils.py when /root/reference
exists."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import entropy_oracle as EO
from oracle import rans_oracle as R
from ref_shims import have_reference, REFERENCE_ROOT

G = os.path.join(os.path.dirname(__file__), "golden")


def _tables(rng, rows=5):
    cdfs, lens, offs = [], [], []
    for r in range(rows):
        k = int(rng.integers(2, 12))
        pmf = rng.random(k) + 1e-3
        pmf /= pmf.sum() * 1.05
        tail = max(1.0 - pmf.sum(), 1e-6)
        row = EO.pmf_to_quantized_cdf(np.concatenate([pmf, [tail]]).astype(np.float32), 16)
        cdfs.append([int(v) for v in row])
        lens.append(len(row))
        offs.append(-int(rng.integers(0, k)))
    return cdfs, lens, offs


def test_round_trip_with_bypass_and_sliced_decoding():
    rng = np.random.default_rng(1)
    cdfs, lens, offs = _tables(rng)
    idx = rng.integers(0, len(cdfs), 20000)
    sym = np.array([rng.integers(offs[i] - 40, offs[i] + lens[i] + 40) if rng.random() < 0.05
                    else rng.integers(offs[i], offs[i] + lens[i] - 2) for i in idx])
    s = R.encode_with_indexes(sym, idx, cdfs, lens, offs)
    assert len(s) % 4 == 0
    d = R.RansDecoder()
    d.set_stream(s)
    out = []
    for a in range(0, 20000, 3333):          # slice by slice, as the CHARM loop decodes
        out += d.decode_stream(idx[a:a + 3333], cdfs, lens, offs)
    assert np.array_equal(np.array(out), sym)
    # the stream ends exactly where the decoder stops reading, and the final state is the encoder's initial one
    assert int(d.state[1]) == len(s) // 4 and int(d.state[0]) == 1 << 31


def test_known_answer_stream():
    fx = np.load(os.path.join(G, "rans_kat.npz"))
    s = R.encode_with_indexes(fx["symbols"], fx["indexes"], fx["cdf"], fx["lengths"], fx["offsets"])
    assert s == fx["stream"].tobytes()
    assert R.RansDecoder().decode_with_indexes(s, fx["indexes"], fx["cdf"], fx["lengths"], fx["offsets"]) == \
        fx["symbols"].tolist()


def test_compression_is_close_to_the_entropy():
    rng = np.random.default_rng(2)
    pmf = np.array([0.5, 0.25, 0.125, 0.0625, 0.0625 - 1e-4], dtype=np.float32)
    row = EO.pmf_to_quantized_cdf(np.concatenate([pmf, [1e-4]]).astype(np.float32), 16)
    cdfs, lens, offs = [[int(v) for v in row]], [len(row)], [0]
    sym = rng.choice(5, size=100000, p=pmf / pmf.sum())
    s = R.encode_with_indexes(sym, np.zeros_like(sym), cdfs, lens, offs)
    h = -(pmf / pmf.sum() * np.log2(pmf / pmf.sum())).sum()
    assert abs(len(s) * 8 / 100000 - h) < 0.01 * h


def test_wire_format_restatement():
    hdr = R.header_encode((512, 768), 37, 3)
    assert hdr == bytes([0, 2, 0, 3, 37, 3]) and len(hdr) == 6
    assert R.header_decode(hdr) == {"img_size": (512, 768), "max_sample": 37, "quality_ind": 3}
    blob = R.pack_strings([hdr, b"zz", b"yyyyy"])
    assert blob[:4] == (6).to_bytes(4, "little") and R.unpack_strings(blob) == [hdr, b"zz", b"yyyyy"]
    from dc_vic_b200 import bitstream as BS
    assert BS.HeaderHandler().encode((512, 768), torch.tensor([[-37.0, 12.0]]), 3) == hdr
    assert BS.HeaderHandler().decode(hdr) == R.header_decode(hdr)
    assert BS.pack_byte_strings([hdr, b"zz", b"yyyyy"]) == blob and BS.unpack_byte_strings(blob) == [hdr, b"zz", b"yyyyy"]


@pytest.mark.skipif(not have_reference(), reason="/root/reference is not present on this box")
def test_wire_format_against_the_reference_file(tmp_path):
    """Pin: the reference's own codec_utils.py (plain numpy/torch, importable as it is)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_codec_utils", os.path.join(REFERENCE_ROOT, "src/utils/codec_utils.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from dc_vic_b200 import bitstream as BS
    y_hat = torch.tensor([[3.0, -41.0, 7.5]])
    for size, q in (((512, 768), 0), ((1365, 2048), 4), ((64, 64), 2)):
        h_ref = ref.HeaderHandler().encode(size, y_hat, q)
        assert BS.HeaderHandler().encode(size, y_hat, q) == h_ref == R.header_encode(size, 41, q)
        assert BS.HeaderHandler().decode(h_ref) == ref.HeaderHandler().decode(h_ref)
    strings = [h_ref, os.urandom(123), os.urandom(4567)]
    p1, p2 = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    ref.save_byte_strings(p1, strings)
    BS.save_byte_strings(p2, strings)
    assert open(p1, "rb").read() == open(p2, "rb").read() == R.pack_strings(strings)
    assert BS.load_byte_strings(p1) == ref.load_byte_strings(p2) == strings
