"""GPU rANS coder + CDF tables + compress / decompress surface against oracle/rans_oracle.c:
byte-for-byte streams, exact round trips, the reference's call patterns (EntropyModel.compress per image,
RansDecoder.set_stream / decode_stream slice by slice, the 3-string .bin layout)."""
import os

import numpy as np
import pytest
import torch

import dc_vic_b200 as D
from dc_vic_b200 import bitstream as BS
from dc_vic_b200 import rans
from oracle import entropy_oracle as EO
from oracle import rans_oracle as R
from synth import entropy_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = os.path.join(os.path.dirname(__file__), "golden")


def test_known_answer_stream_on_the_gpu():
    fx = np.load(os.path.join(G, "rans_kat.npz"))
    s = rans.encode_with_indexes(torch.from_numpy(fx["symbols"]), torch.from_numpy(fx["indexes"]),
                                 torch.from_numpy(fx["cdf"]), fx["lengths"].tolist(), fx["offsets"].tolist())
    assert s == fx["stream"].tobytes()
    out = rans.RansDecoder().decode_with_indexes(s, fx["indexes"].tolist(), fx["cdf"].tolist(), fx["lengths"].tolist(),
                                                 fx["offsets"].tolist())
    assert out == fx["symbols"].tolist()


def test_streams_match_the_oracle_byte_for_byte_including_bypass_and_batches():
    rng = np.random.default_rng(5)
    rows, width = 9, 40
    cdf = np.zeros((rows, width), dtype=np.int32)
    lens, offs = [], []
    for r in range(rows):
        k = int(rng.integers(2, 36))
        pmf = (rng.random(k) ** 3 + 1e-4).astype(np.float32)
        pmf /= pmf.sum() * 1.01
        row = EO.pmf_to_quantized_cdf(np.concatenate([pmf, [max(1 - pmf.sum(), 1e-7)]]).astype(np.float32), 16)
        cdf[r, : len(row)] = row
        lens.append(len(row))
        offs.append(-int(rng.integers(0, k)))
    tab = rans._Tables(torch.from_numpy(cdf), lens, offs, torch.device(DEV))
    syms, idxs = [], []
    for n in (1, 31, 32, 33, 1000, 50001):               # chunk boundaries of the 32-symbol look-up groups
        idx = rng.integers(0, rows, n)
        sym = np.array([rng.integers(offs[i] - 300, offs[i] + lens[i] + 300) if rng.random() < 0.03
                        else rng.integers(offs[i], offs[i] + lens[i] - 2) for i in idx])
        syms.append(sym)
        idxs.append(idx)
    got = rans.encode_batch([torch.from_numpy(s).int().to(DEV) for s in syms],
                            [torch.from_numpy(i).int().to(DEV) for i in idxs], tab)
    for s, i, g in zip(syms, idxs, got):
        assert g == R.encode_with_indexes(s, i, cdf, lens, offs)
    # decode the longest one slice by slice, as minnen20_charm_context_model.py:175-202 does
    dec = rans.RansDecoder()
    dec.set_stream(got[-1])
    cdf_l, out = cdf.tolist(), []
    for a in range(0, 50001, 8192):
        out += dec.decode_stream(idxs[-1][a:a + 8192].tolist(), cdf_l, lens, offs)
    assert np.array_equal(np.array(out), syms[-1])


def test_entropy_bottleneck_compress_decompress():
    torch.manual_seed(3)
    eb = D.SteEntropyBottleneck(channels=12).to(DEV)
    with torch.no_grad():
        eb.quantiles[:, 0, 0] -= 3 * torch.rand(12, device=DEV)
        eb.quantiles[:, 0, 2] += 3 * torch.rand(12, device=DEV)
        eb.quantiles[:, 0, 1] += torch.randn(12, device=DEV)
    eb.update(force=True)
    z = 6 * torch.randn(3, 12, 8, 12, device=DEV)        # beyond the tabulated range too: bypass codes
    strings = eb.compress(z)
    assert len(strings) == 3 and all(isinstance(s, bytes) for s in strings)
    with torch.no_grad():
        z_hat, _ = eb(z, is_train=False)
    back = eb.decompress(strings, (8, 12))
    assert torch.equal(back, z_hat)
    # the strings are what the oracle coder writes for the same symbols and tables
    med = eb._get_medians().detach().reshape(1, 12, 1, 1)
    sym = torch.round(z - med).int().cpu().numpy()
    idx = np.broadcast_to(np.arange(12, dtype=np.int32).reshape(1, 12, 1, 1), sym.shape)
    cdf, ln, off = eb._quantized_cdf.cpu().numpy(), eb._cdf_length.cpu().numpy(), eb._offset.cpu().numpy()
    for b in range(3):
        assert strings[b] == R.encode_with_indexes(sym[b], idx[b], cdf, ln, off)
    # the CPU-tensor call pattern of codec_setup / _compress_estimate_entropy (the module stays on the GPU)
    eb.to("cpu")
    assert eb.compress(z.cpu()) == strings
    assert torch.equal(eb.decompress(strings, (8, 12)).cpu(), z_hat.cpu())


def test_gaussian_conditional_bin_file_round_trip(tmp_path):
    """hyperprior_dc_vic_model.py:308-328 / :378-387 on a kodim03-sized latent, then the 3-string .bin file."""
    gc = D.SteGaussianMeanScaleConditional(scale_bound=0.11).to(DEV)
    gc.update_scale_table(D.get_scale_table(), force=True)
    y, params = entropy_inputs(1, B=1, C=192, H=32, W=48)
    y, params = y.to(DEV), params.to(DEV)
    means, scales = params.chunk(2, 1)
    indexes = gc.build_indexes(scales)
    y_str = gc.compress(y, indexes, means=means)
    with torch.no_grad():
        y_hat, lik = gc(y, params, is_train=False)
    assert torch.equal(gc.decompress(y_str, indexes, means=means), y_hat)
    bits = float(-torch.log2(lik).sum())
    assert abs(len(y_str[0]) * 8 - bits) < 0.02 * bits + 64          # the real size tracks the estimated rate
    sym = torch.round(y - means).int().cpu().numpy().reshape(-1)
    assert y_str[0] == R.encode_with_indexes(sym, indexes.cpu().numpy().reshape(-1), gc._quantized_cdf.cpu().numpy(),
                                             gc._cdf_length.cpu().numpy(), gc._offset.cpu().numpy())
    header = BS.HeaderHandler().encode((512, 768), y_hat, 0)
    path = str(tmp_path / "kodim03.bin")
    BS.save_byte_strings(path, [header, b"\x00" * 8, y_str[0]])
    h2, _, y2 = BS.load_byte_strings(path)
    assert BS.HeaderHandler().decode(h2)["img_size"] == (512, 768) and y2 == y_str[0]
    assert os.path.getsize(path) == 6 + 8 + len(y_str[0]) + 12


def test_cdf_tables_are_built_on_the_device_bit_exactly():
    """_pmf_to_cdf on the GPU (one thread per row) == the host construction == the oracle's, on the real tables."""
    gc = D.SteGaussianMeanScaleConditional(scale_bound=0.11).to(DEV)
    gc.update_scale_table(D.get_scale_table(), force=True)
    eb = D.SteEntropyBottleneck(channels=16).to(DEV)
    with torch.no_grad():
        eb.quantiles[:, 0, 0] -= 5 * torch.rand(16, device=DEV)
        eb.quantiles[:, 0, 2] += 5 * torch.rand(16, device=DEV)
    eb.update(force=True)
    for m in (gc, eb):
        cdf, ln = m._quantized_cdf, m._cdf_length
        assert cdf.is_cuda and cdf.dtype == torch.int32
        for r in range(cdf.shape[0]):
            row = cdf[r, : int(ln[r])].cpu()
            assert int(row[0]) == 0 and int(row[-1]) == 65536 and bool((row[1:] > row[:-1]).all())
            assert bool((cdf[r, int(ln[r]):] == 0).all())
    # same PMF rows through the host entry point and through the oracle
    pmf = torch.rand(7, 33, device=DEV) ** 4 + 1e-7
    pmf = pmf / pmf.sum(1, keepdim=True) * 0.97
    lens = torch.tensor([33, 5, 17, 1, 33, 20, 9], dtype=torch.int32, device=DEV)
    tail = torch.rand(7, 1, device=DEV) * 0.03 + 1e-6
    dev_cdf = gc._pmf_to_cdf(pmf, tail, lens, 33)
    host_cdf = gc._pmf_to_cdf(pmf.cpu(), tail.cpu(), lens.cpu(), 33)
    assert torch.equal(dev_cdf.cpu(), host_cdf)
    for r in range(7):
        n = int(lens[r])
        want = EO.pmf_to_quantized_cdf(np.concatenate([pmf[r, :n].cpu().numpy(), tail[r].cpu().numpy()]), 16)
        assert dev_cdf[r, : n + 2].cpu().tolist() == [int(v) for v in want]
