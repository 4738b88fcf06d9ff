"""Cross-checks of the CompressAI-1.2.4 restatement (parity UNPINNED: no wheel, no reference
goldens -- see oracle/entropy_oracle.py).  What can be checked without the wheel: FP64
evaluation of the same formulas, closed-form Gaussian identities, table invariants and the
DC-VIC wrapper semantics (reference src/models/subnet/entropy_model/*.py).  CPU only.
"""
import math

import numpy as np
import pytest
import torch
from scipy.stats import norm

from oracle import entropy_oracle as E
from synth import entropy_inputs, entropy_inputs_init, noise_like


def test_lower_bound_gradient_rule():
    x = torch.tensor([0.05, 0.2, 0.05, 0.2], requires_grad=True)
    lb = E.LowerBound(0.11)
    y = lb(x)
    assert torch.allclose(y, torch.tensor([0.11, 0.2, 0.11, 0.2]))
    y.backward(torch.tensor([1.0, 1.0, -1.0, -1.0]))
    assert torch.equal(x.grad, torch.tensor([0.0, 1.0, -1.0, -1.0]))


def test_quantize_modes():
    m = E.EntropyModel()
    x = torch.tensor([0.5, 1.5, 2.5, -0.5, 1.2])
    mu = torch.tensor([0.25, 0.25, 0.25, 0.25, 0.25])
    assert torch.equal(m.quantize(x, "dequantize"), torch.tensor([0.0, 2.0, 2.0, -0.0, 1.0]))  # half-to-even
    assert torch.equal(m.quantize(x, "symbols", mu), torch.round(x - mu).int())
    assert torch.equal(m.quantize(x, "dequantize", mu), torch.round(x - mu) + mu)
    nz = torch.full_like(x, 0.25)
    assert torch.equal(m.quantize(x, "noise", mu, noise=nz), x + nz)   # means ignored in noise mode
    with pytest.raises(ValueError):
        m.quantize(x, "bogus")


@pytest.mark.parametrize("q", [0, 2, 4])
def test_gaussian_likelihood_fp32_vs_fp64_and_closed_form(q):
    y, params = entropy_inputs(q, B=2, C=8, H=16, W=16)
    gc = E.GaussianMeanScaleConditional(scale_bound=0.11)
    y_hat, lik = gc(y, params, is_train=False)
    mu, sig = params.chunk(2, 1)
    assert torch.equal(y_hat, torch.round(y - mu) + mu)
    # closed form in FP64
    s = sig.double().clamp_min(0.11)
    v = (y_hat.double() - mu.double()).abs()
    ref = norm.cdf(((0.5 - v) / s).numpy()) - norm.cdf(((-0.5 - v) / s).numpy())
    ref = np.maximum(ref, 1e-9)
    rel = np.abs(lik.double().numpy() - ref) / ref
    assert rel.max() < 1e-4, rel.max()
    gc64 = E.GaussianMeanScaleConditional(scale_bound=0.11).double()
    lik64 = gc64._likelihood(y_hat.double(), sig.double(), mu.double()).clamp_min(1e-9)
    assert np.allclose(lik64.numpy(), ref, rtol=1e-9, atol=1e-15)
    bits32, _ = E.likelihood_to_bit(lik, 1)
    bits64 = -np.log2(ref).sum()
    assert abs(float(bits32) - bits64) / bits64 < 1e-5


def test_gaussian_scale_bound_and_likelihood_bound():
    y, params = entropy_inputs_init()
    gc = E.SteGaussianMeanScaleConditional(scale_bound=0.11)
    y_hat, lik = gc(y, params, is_train=False)
    assert float(lik.min()) >= float(torch.tensor(1e-9)) and float(lik.max()) <= 1.0
    mu, _ = params.chunk(2, 1)
    assert torch.equal(y_hat, torch.round(y - mu) + mu)
    nz = noise_like(y, 200)
    y_hat_t, lik_t = gc(y, params, is_train=True, noise=nz)
    assert torch.equal(y_hat_t, (torch.round(y - mu) - (y - mu)) + (y - mu) + mu)
    _, lik_plain = E.GaussianMeanScaleConditional(scale_bound=0.11)(y, params, is_train=True, noise=nz)
    assert torch.equal(lik_t, lik_plain)
    with pytest.raises(TypeError):
        E.GaussianMeanScaleConditional(scale_bound=None)     # 1.2.4 behaviour the YAML relies on never hitting


def test_gaussian_training_gradients_match_fp64():
    y, params = entropy_inputs(1, B=1, C=4, H=8, W=8)
    nz = noise_like(y, 201)
    outs = []
    for dt in (torch.float32, torch.float64):
        yy = y.detach().clone().to(dt).requires_grad_(True)
        pp = params.detach().clone().to(dt).requires_grad_(True)
        gc = E.GaussianMeanScaleConditional(scale_bound=0.11).to(dt)
        _, lik = gc(yy, pp, is_train=True, noise=nz.to(dt))
        bits, _ = E.likelihood_to_bit(lik, 1)
        bits.backward()
        outs.append((yy.grad.double(), pp.grad.double()))
    assert torch.allclose(outs[0][0], outs[1][0], rtol=2e-3, atol=1e-5)
    assert torch.allclose(outs[0][1], outs[1][1], rtol=2e-3, atol=1e-5)


def test_scale_table_and_build_indexes():
    t = E.get_scale_table()
    assert t.numel() == 64 and abs(float(t[0]) - 0.11) < 1e-6 and abs(float(t[-1]) - 256) < 1e-3
    gc = E.GaussianMeanScaleConditional(scale_bound=0.11)
    gc.update_scale_table(t)
    s = torch.tensor([0.0, 0.11, 0.1100001, 1.0, 255.9, 256.0, 1e4, -3.0])
    idx = gc.build_indexes(s)
    brute = torch.tensor([min(int((t < max(float(v), 0.11)).sum()), 63) for v in s], dtype=torch.int32)
    assert torch.equal(idx, brute)
    assert gc._quantized_cdf.shape[0] == 64
    L = gc._cdf_length
    for i in (0, 31, 63):
        row = gc._quantized_cdf[i, : int(L[i])]
        assert int(row[0]) == 0 and int(row[-1]) == 65536 and bool((row[1:] > row[:-1]).all())
    assert torch.equal(gc._offset, -torch.ceil(t * (-norm.ppf(1e-9 / 2))).int())
    assert gc.update_scale_table(t) is False and gc.update_scale_table(t, force=True) is True


def test_pmf_to_quantized_cdf_invariants():
    rng = np.random.default_rng(0)
    for n in (2, 5, 33, 400):
        p = rng.random(n).astype(np.float32) ** 8
        p /= p.sum()
        cdf = E.pmf_to_quantized_cdf(p, 16)
        assert cdf.dtype == np.int32 and cdf.size == n + 1
        assert cdf[0] == 0 and cdf[-1] == 65536 and np.all(np.diff(cdf) >= 1)
    # known answer: uniform 4-symbol pmf
    assert E.pmf_to_quantized_cdf([0.25] * 4).tolist() == [0, 16384, 32768, 49152, 65536]
    # zero-probability symbol steals one count from the cheapest donor with freq > 1
    assert E.pmf_to_quantized_cdf([0.5, 0.0, 0.5]).tolist() == [0, 32767, 32768, 65536]


def test_entropy_bottleneck_init_forward_loss_update():
    torch.manual_seed(7)
    eb = E.SteEntropyBottleneck(channels=6)
    names = sorted(n for n, _ in eb.named_parameters())
    assert names == sorted([f"_matrix{i}" for i in range(5)] + [f"_bias{i}" for i in range(5)]
                           + [f"_factor{i}" for i in range(4)] + ["quantiles"])
    assert eb._matrix0.shape == (6, 3, 1) and eb._matrix4.shape == (6, 1, 3) and eb.quantiles.shape == (6, 1, 3)
    assert abs(float(eb.target[2]) - math.log(2 / 1e-9 - 1)) < 1e-5
    x = 3 * torch.randn(2, 6, 4, 5)
    x_hat, lik = eb(x, is_train=False)
    med = eb._get_medians().view(1, 6, 1, 1)
    assert torch.equal(x_hat, torch.round(x - med) + med)
    assert lik.shape == x.shape and float(lik.min()) >= float(torch.tensor(1e-9)) and float(lik.max()) <= 1
    # FP64 agreement
    eb64 = E.SteEntropyBottleneck(channels=6).double()
    eb64.load_state_dict({k: v.double() for k, v in eb.state_dict().items()})
    _, lik64 = eb64(x.double(), is_train=False)
    assert float(((lik.double() - lik64).abs() / lik64).max()) < 1e-4
    # the density integrates to ~1 over the integer grid (factorized CDF is a proper CDF)
    grid = torch.arange(-400, 401).float().view(1, 1, -1, 1).expand(1, 6, -1, 1).contiguous()
    _, pg = eb(grid, is_train=False)
    assert torch.allclose(pg.sum(dim=2).flatten(), torch.ones(6), atol=1e-3)
    # STE train output and aux loss
    nz = noise_like(x, 9)
    xh, _ = eb(x, is_train=True, noise=nz)
    assert torch.allclose(xh, torch.round(x - med) + med, atol=1e-6)
    loss = eb.loss()
    loss.backward()
    assert eb.quantiles.grad is not None and eb._matrix0.grad is None
    assert eb.update() is True and eb.update() is False
    assert eb._quantized_cdf.shape[0] == 6 and int(eb._quantized_cdf[:, 0].abs().sum()) == 0
    for c in range(6):
        row = eb._quantized_cdf[c, : int(eb._cdf_length[c])]
        assert int(row[-1]) == 65536 and bool((row[1:] > row[:-1]).all())


def test_rate_summary():
    lik = torch.tensor([[0.5, 0.25], [0.125, 1.0]])
    bits, bpp = E.likelihood_to_bit(lik, 4)
    assert abs(float(bits) - 6.0) < 1e-6 and abs(float(bpp) - 1.5) < 1e-6
    assert torch.allclose(E.batch_bits(lik), torch.tensor([3.0, 3.0]))
