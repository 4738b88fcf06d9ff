"""CPU oracle for the rate/entropy model -- TEST INFRASTRUCTURE, NOT PRODUCT.

PARITY UNPINNED at the CompressAI boundary.  The arithmetic of this half of the hot path
lives in the un-vendored dependency ``compressai==1.2.4`` (reference pyproject.toml:13,
poetry.lock:312-313).  It is not installed here, there is no wheel for it in the offline
wheelhouse and the reference ships neither tests nor golden vectors for it, so this file
restates the published 1.2.4 algorithm (``compressai/entropy_models/entropy_models.py``,
``compressai/ops/bound_ops.py``, ``compressai/cpp_exts/ops/ops.cpp``) from its
documentation/knowledge of the source, anchored on the reference's own call sites:

* ``src/models/subnet/entropy_model/entropy_bottleneck.py:13-28``   (EntropyBottleneck, Ste..)
* ``src/models/subnet/entropy_model/gaussian_conditional.py:9-24``  (Gaussian{Scale,MeanScale}Conditional)
* ``src/models/subnet/entropy_model/ste_gaussian_conditional.py:9-23``
* ``src/models/subnet/entropy_model/ste_round.py:4-5``
* ``src/models/comp_model/hyperprior_vic_model.py:66-82``           (likelihood_to_bit, rate summary)
* ``src/models/subnet/context_model/minnen20_charm_context_model.py:84-102,164`` (per-slice calls)

Cross-checks available without the wheel (tests/test_oracle_entropy.py): FP64 evaluation
of the same formulas, closed-form Gaussian identities (scipy.stats.norm), PMF tables that
sum to one, CDF-table invariants of ``pmf_to_quantized_cdf``.  All functions are plain torch
CPU ops so ``torch.autograd`` on the oracle is the backward oracle as well.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor


# --------------------------------------------------------------------------- bound_ops.py
class _LowerBoundFn(torch.autograd.Function):
    """max(x, bound); gradient passes where x >= bound or the gradient pushes x upward."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        keep = (x >= bound) | (grad_output < 0)
        return keep * grad_output, None


class LowerBound(nn.Module):
    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x: Tensor) -> Tensor:
        return _LowerBoundFn.apply(x, self.bound.to(x.dtype))


# ------------------------------------------------------------------- ste_round.py:4-5
def ste_round(x: Tensor) -> Tensor:
    return (torch.round(x) - x).detach() + x


# --------------------------------------------------- cpp_exts/ops/ops.cpp (ryg_rans port)
def pmf_to_quantized_cdf(pmf, precision: int = 16) -> np.ndarray:
    """Integer CDF with every symbol given frequency >= 1 ("steal" from the cheapest donor)."""
    p = np.asarray(pmf, dtype=np.float32)
    if not np.all(np.isfinite(p)) or np.any(p < 0):
        raise ValueError("invalid pmf")
    scale = np.float32(1 << precision)
    freq = np.empty(p.size + 1, dtype=np.uint64)
    freq[0] = 0
    # std::round (half away from zero) of a float product
    prod = (p * scale).astype(np.float32)
    freq[1:] = np.floor(prod.astype(np.float64) + 0.5).astype(np.uint64)
    total = int(freq.sum())
    if total == 0:
        raise ValueError("pmf sums to zero")
    freq = (np.uint64(1 << precision) * freq) // np.uint64(total)
    cdf = np.cumsum(freq).astype(np.int64)
    cdf[-1] = 1 << precision
    n = cdf.size
    for i in range(n - 1):
        if cdf[i] == cdf[i + 1]:
            widths = cdf[1:] - cdf[:-1]
            ok = widths > 1
            if not ok.any():
                raise ValueError("cannot build CDF: no donor symbol")
            masked = np.where(ok, widths, np.iinfo(np.int64).max)
            donor = int(np.argmin(masked))  # first minimum == strict '<' scan
            if donor < i:
                cdf[donor + 1:i + 1] -= 1
            else:
                cdf[i + 1:donor + 1] += 1
    assert cdf[0] == 0 and cdf[-1] == (1 << precision)
    assert np.all(cdf[1:] > cdf[:-1])
    return cdf.astype(np.int32)


def get_scale_table(lo: float = 0.11, hi: float = 256.0, levels: int = 64) -> Tensor:
    """compressai.models.google.get_scale_table (used at hyperprior_dc_vic_model.py:66-68)."""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


# ------------------------------------------------------------------------- EntropyModel
class EntropyModel(nn.Module):
    def __init__(self, likelihood_bound: float = 1e-9, entropy_coder_precision: int = 16):
        super().__init__()
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    def quantize(self, inputs: Tensor, mode: str, means: Optional[Tensor] = None,
                 noise: Optional[Tensor] = None) -> Tensor:
        """``noise``: x + U(-.5,.5) (means ignored); else round(x - means) [+ means | .int()].

        ``noise=`` lets a test inject the uniform sample the CUDA path is given, so that
        training-mode outputs can be compared exactly (SURVEY hard part 5).
        """
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            if noise is None:
                noise = torch.empty_like(inputs).uniform_(-0.5, 0.5)
            return inputs + noise
        outputs = inputs.clone()
        if means is not None:
            outputs -= means
        outputs = torch.round(outputs)
        if mode == "dequantize":
            if means is not None:
                outputs += means
            return outputs
        return outputs.int()

    @staticmethod
    def dequantize(inputs: Tensor, means: Optional[Tensor] = None, dtype=torch.float) -> Tensor:
        if means is not None:
            outputs = inputs.type_as(means)
            outputs += means
        else:
            outputs = inputs.type(dtype)
        return outputs

    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
        for i, p in enumerate(pmf):
            prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
            row = torch.from_numpy(pmf_to_quantized_cdf(prob.tolist(), self.entropy_coder_precision))
            cdf[i, : row.numel()] = row
        return cdf


# ------------------------------------------------------------------ GaussianConditional
class GaussianConditional(EntropyModel):
    def __init__(self, scale_table=None, scale_bound: float = 0.11, tail_mass: float = 1e-9, **kw):
        super().__init__(**kw)
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = scale_table[0]
        if scale_bound <= 0:  # NB: scale_bound=None raises TypeError, as in 1.2.4
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table",
                             torch.Tensor(tuple(float(s) for s in scale_table)) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))

    @staticmethod
    def _standardized_cumulative(x: Tensor) -> Tensor:
        return 0.5 * torch.erfc(-(2 ** -0.5) * x)

    def _likelihood(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None) -> Tensor:
        values = inputs - means if means is not None else inputs
        scales = self.lower_bound_scale(scales)
        values = torch.abs(values)
        upper = self._standardized_cumulative((0.5 - values) / scales)
        lower = self._standardized_cumulative((-0.5 - values) / scales)
        return upper - lower

    def forward(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                training: Optional[bool] = None, noise: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        if training is None:
            training = self.training
        outputs = self.quantize(inputs, "noise" if training else "dequantize", means, noise=noise)
        likelihood = self._likelihood(outputs, scales, means)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        return outputs, likelihood

    def build_indexes(self, scales: Tensor) -> Tensor:
        scales = self.lower_bound_scale(scales)
        indexes = scales.new_full(scales.size(), len(self.scale_table) - 1).int()
        for s in self.scale_table[:-1]:
            indexes -= (scales <= s).int()
        return indexes

    def update_scale_table(self, scale_table, force: bool = False) -> bool:
        if self._offset.numel() > 0 and not force:
            return False
        self.scale_table = torch.Tensor(tuple(float(s) for s in scale_table))
        self.update()
        return True

    def update(self):
        from scipy.stats import norm
        multiplier = -norm.ppf(self.tail_mass / 2)
        pmf_center = torch.ceil(self.scale_table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = int(torch.max(pmf_length).item())
        samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
        samples_scale = self.scale_table.unsqueeze(1).float()
        upper = self._standardized_cumulative((0.5 - samples) / samples_scale)
        lower = self._standardized_cumulative((-0.5 - samples) / samples_scale)
        pmf = upper - lower
        tail_mass = 2 * lower[:, :1]
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self._offset = -pmf_center
        self._cdf_length = pmf_length + 2


# -------------------------------------------------------------------- EntropyBottleneck
class EntropyBottleneck(EntropyModel):
    def __init__(self, channels: int, tail_mass: float = 1e-9, init_scale: float = 10,
                 filters=(3, 3, 3, 3), **kw):
        super().__init__(**kw)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        f = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / f[i + 1]))
            matrix = torch.Tensor(channels, f[i + 1], f[i])
            matrix.data.fill_(init)
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(matrix))
            bias = torch.Tensor(channels, f[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))
            if i < len(self.filters):
                factor = torch.Tensor(channels, f[i + 1], 1)
                nn.init.zeros_(factor)
                self.register_parameter(f"_factor{i:d}", nn.Parameter(factor))
        self.quantiles = nn.Parameter(torch.Tensor(channels, 1, 3))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles.data = init.repeat(self.quantiles.size(0), 1, 1)
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def _get_medians(self) -> Tensor:
        return self.quantiles[:, :, 1:2]

    def _logits_cumulative(self, inputs: Tensor, stop_gradient: bool) -> Tensor:
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = getattr(self, f"_matrix{i:d}")
            bias = getattr(self, f"_bias{i:d}")
            if stop_gradient:
                matrix, bias = matrix.detach(), bias.detach()
            logits = torch.matmul(F.softplus(matrix), logits)
            logits = logits + bias
            if i < len(self.filters):
                factor = getattr(self, f"_factor{i:d}")
                if stop_gradient:
                    factor = factor.detach()
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def _likelihood(self, inputs: Tensor) -> Tensor:
        lower = self._logits_cumulative(inputs - 0.5, stop_gradient=False)
        upper = self._logits_cumulative(inputs + 0.5, stop_gradient=False)
        sign = (-torch.sign(lower + upper)).detach()
        return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))

    def forward(self, x: Tensor, training: Optional[bool] = None,
                noise: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        if training is None:
            training = self.training
        perm = list(range(x.dim()))
        perm[0], perm[1] = 1, 0
        xc = x.permute(*perm).contiguous()
        shape = xc.size()
        values = xc.reshape(xc.size(0), 1, -1)
        nz = None
        if noise is not None:
            nz = noise.permute(*perm).contiguous().reshape(xc.size(0), 1, -1)
        outputs = self.quantize(values, "noise" if training else "dequantize", self._get_medians(), noise=nz)
        likelihood = self._likelihood(outputs)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        outputs = outputs.reshape(shape).permute(*perm).contiguous()
        likelihood = likelihood.reshape(shape).permute(*perm).contiguous()
        return outputs, likelihood

    def loss(self) -> Tensor:
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def update(self, force: bool = False) -> bool:
        if self._offset.numel() > 0 and not force:
            return False
        medians = self.quantiles[:, 0, 1]
        minima = torch.clamp(torch.ceil(medians - self.quantiles[:, 0, 0]).int(), min=0)
        maxima = torch.clamp(torch.ceil(self.quantiles[:, 0, 2] - medians).int(), min=0)
        self._offset = -minima
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = int(pmf_length.max().item())
        samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
        with torch.no_grad():
            lower = self._logits_cumulative(samples - 0.5, stop_gradient=True)
            upper = self._logits_cumulative(samples + 0.5, stop_gradient=True)
            sign = -torch.sign(lower + upper)
            pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
            tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self._cdf_length = pmf_length + 2
        return True


# ----------------------------------------------- DC-VIC wrappers (reference, in-tree files)
class DcvicEntropyBottleneck(EntropyBottleneck):
    """entropy_bottleneck.py:13-16."""

    def forward(self, x: Tensor, is_train: bool, noise: Optional[Tensor] = None):
        return super().forward(x, training=is_train, noise=noise)


class SteEntropyBottleneck(DcvicEntropyBottleneck):
    """entropy_bottleneck.py:19-28."""

    def forward(self, x: Tensor, is_train: bool = True, noise: Optional[Tensor] = None):
        if not is_train:
            return super().forward(x, is_train)
        _, lik = super().forward(x, is_train, noise=noise)
        mu = self._get_medians()
        return ste_round(x - mu) + mu, lik


class GaussianScaleConditional(GaussianConditional):
    """gaussian_conditional.py:9-15."""

    def __init__(self, scale_bound=None):
        super().__init__(scale_table=None, scale_bound=scale_bound)

    def forward(self, y: Tensor, params: Tensor, is_train: bool = True, noise: Optional[Tensor] = None):
        return super().forward(y, scales=params, means=None, training=is_train, noise=noise)


class GaussianMeanScaleConditional(GaussianConditional):
    """gaussian_conditional.py:17-24."""

    def __init__(self, scale_bound=None):
        super().__init__(scale_table=None, scale_bound=scale_bound)

    def forward(self, y: Tensor, params: Tensor, is_train: bool = True, noise: Optional[Tensor] = None):
        mean, std = params.chunk(2, 1)
        return super().forward(y, scales=std, means=mean, training=is_train, noise=noise)


class SteGaussianMeanScaleConditional(GaussianMeanScaleConditional):
    """ste_gaussian_conditional.py:9-23."""

    def __init__(self, scale_bound=None, entropy_quant_type="noise", **kwargs):
        super().__init__(scale_bound=scale_bound)
        assert entropy_quant_type == "noise"
        self.entropy_quant_type = entropy_quant_type

    def forward(self, y: Tensor, params: Tensor, is_train: bool = True, noise: Optional[Tensor] = None):
        mean, _ = params.chunk(2, 1)
        _, lik = super().forward(y, params, is_train=is_train, noise=noise)
        if is_train:
            y_hat = ste_round(y - mean) + mean
        else:
            y_hat = self.quantize(y, mode="dequantize", means=mean)
        return y_hat, lik


# ------------------------------------------------- hyperprior_vic_model.py:80-82, 66-78
def likelihood_to_bit(likelihood: Tensor, num_pixel: int) -> Tuple[Tensor, Tensor]:
    bitcost = -(torch.log(likelihood).sum()) / np.log(2)
    return bitcost, bitcost / num_pixel


def batch_bits(likelihood: Tensor) -> Tensor:
    """Per-sample bit cost (dual_cond_rate_distortion_vq_code_trainer.py:100-108)."""
    return -(torch.log(likelihood).flatten(1).sum(dim=1)) / np.log(2)
