"""B200-native rate / entropy models with the reference's module interfaces.

Two layers, as in the reference:

* the CompressAI 1.2.4 surface DC-VIC depends on (un-vendored there: pyproject.toml:13) --
  ``EntropyModel``, ``GaussianConditional``, ``EntropyBottleneck``, ``LowerBound`` -- same ctor
  arguments, parameter / buffer names (state-dict compatible: ``_matrix{i}``, ``_bias{i}``,
  ``_factor{i}``, ``quantiles``, ``target``, ``_offset``, ``_quantized_cdf``, ``_cdf_length``,
  ``scale_table``, ``scale_bound``, ``*.bound``) and forward signatures;
* DC-VIC's own wrappers from ``src/models/subnet/entropy_model/`` -- ``DcvicEntropyBottleneck``
  (registered there as ``EntropyBottleneck``), ``SteEntropyBottleneck``,
  ``GaussianScaleConditional``, ``GaussianMeanScaleConditional``,
  ``SteGaussianMeanScaleConditional`` (entropy_bottleneck.py:13-28, gaussian_conditional.py:9-24,
  ste_gaussian_conditional.py:9-23) plus ``ste_round`` (ste_round.py:4-5).

quantize / likelihood / rate arithmetic runs in hand-written CUDA behind the C ABI
(``dcvic_gc_*``, ``dcvic_eb_*``, ``dcvic_rate_*``); CPU tensors raise -- no fallback.
The rANS bitstream coder (``compress`` / ``decompress``) is a later scope row and raises
NotImplementedError; the CDF tables it needs (``update`` / ``update_scale_table`` /
``build_indexes``) are built here.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from . import _lib

__all__ = ["LowerBound", "EntropyModel", "GaussianConditional", "EntropyBottleneck", "DcvicEntropyBottleneck",
           "SteEntropyBottleneck", "GaussianScaleConditional", "GaussianMeanScaleConditional",
           "SteGaussianMeanScaleConditional", "ste_round", "get_scale_table", "pmf_to_quantized_cdf",
           "likelihood_to_bit", "batch_bits", "gaussian_rate_dual", "gaussian_codec_step", "DevicePinned",
           "FUSED_FORWARDS"]

_WS = {}


def _workspace(tag: str, nbytes: int, device) -> Tensor:
    stream = torch.cuda.current_stream(device).cuda_stream
    key = (tag, device.index, stream)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


def _batch_view(t: Tensor, B: int, n: int) -> Tuple[Tensor, int]:
    """Return (tensor, batch stride in elements) such that sample b is n contiguous floats at
    data_ptr + b*stride.  ``params.chunk(2, 1)`` halves qualify without a copy."""
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() >= 1 and t.shape[0] == B and t[0].is_contiguous() and (B == 1 or t.stride(0) >= n):
        return t, (t.stride(0) if B > 1 else n)
    t = t.contiguous()
    return t, n


# ------------------------------------------------------------------------------ small ops
class _SteRound(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _lib.require_cuda(x)
        xc = x.contiguous().float()
        out = torch.empty_like(xc)
        with _lib.on_device(x.device):
            _lib.check(_lib.load().dcvic_ste_round(_lib.ptr(xc), xc.numel(), _lib.ptr(out), _lib.cur_stream()),
                       "dcvic_ste_round")
        return out

    @staticmethod
    def backward(ctx, g):
        return g


def ste_round(x: Tensor) -> Tensor:
    """ste_round.py:4-5: ``(round(x) - x).detach() + x``."""
    return _SteRound.apply(x)


class _LowerBoundFn(torch.autograd.Function):
    """compressai.ops.LowerBound: only used stand-alone (the fused kernels apply it inline)."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, g):
        x, bound = ctx.saved_tensors
        return ((x >= bound) | (g < 0)) * g, None


class LowerBound(nn.Module):
    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x: Tensor) -> Tensor:
        return _LowerBoundFn.apply(x, self.bound)


def get_scale_table(min: float = 0.11, max: float = 256.0, levels: int = 64) -> Tensor:
    """compressai.models.google.get_scale_table (hyperprior_dc_vic_model.py:66-68)."""
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


def pmf_to_quantized_cdf(pmf, precision: int = 16) -> Tensor:
    """compressai ``_CXX.pmf_to_quantized_cdf`` through the library's host entry point."""
    arr = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    out = np.empty(arr.size + 1, dtype=np.int32)
    rc = _lib.load().dcvic_pmf_to_quantized_cdf(arr.ctypes.data_as(C.POINTER(C.c_float)), int(arr.size),
                                                int(precision), out.ctypes.data_as(C.POINTER(C.c_int32)))
    _lib.check(rc, "dcvic_pmf_to_quantized_cdf")
    return torch.from_numpy(out)


# ------------------------------------------------------------------------------ rate
class _RateBits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lik, B):
        _lib.require_cuda(lik)
        lc = lik.contiguous().float()
        n = lc.numel() // B
        lib = _lib.load()
        with _lib.on_device(lik.device):
            bits = torch.empty(B, dtype=torch.float32, device=lik.device)
            ws = _workspace("rate", lib.dcvic_rate_workspace_bytes(B, n), lik.device)
            _lib.check(lib.dcvic_rate_bits(_lib.ptr(lc), B, n, _lib.ptr(bits), _lib.ptr(ws), ws.numel(),
                                           _lib.cur_stream()), "dcvic_rate_bits")
        ctx.save_for_backward(lc)
        ctx.B = B
        ctx.shape = lik.shape
        return bits

    @staticmethod
    def backward(ctx, g_bits):
        (lc,) = ctx.saved_tensors
        B = ctx.B
        n = lc.numel() // B
        with _lib.on_device(lc.device):
            d = torch.empty_like(lc)
            gb = g_bits.contiguous().float()
            _lib.check(_lib.load().dcvic_rate_bits_backward(_lib.ptr(lc), _lib.ptr(gb), B, n, _lib.ptr(d),
                                                            _lib.cur_stream()), "dcvic_rate_bits_backward")
        return d.view(ctx.shape), None


def likelihood_to_bit(likelihood: Tensor, num_pixel: int) -> Tuple[Tensor, Tensor]:
    """hyperprior_vic_model.py:80-82: ``bits = -(log(L).sum()) / ln 2``; returns (bits, bits/num_pixel)."""
    bits = _RateBits.apply(likelihood, 1).squeeze(0)
    return bits, bits / num_pixel


def batch_bits(likelihood: Tensor) -> Tensor:
    """Per-sample bit cost [B] (dual_cond_rate_distortion_vq_code_trainer.py:100-108)."""
    return _RateBits.apply(likelihood, likelihood.shape[0])


# ------------------------------------------------------------------------------ base
def _module_device(module) -> torch.device:
    for t in module.buffers():
        return t.device
    for t in module.parameters():
        return t.device
    return torch.device("cpu")


class DevicePinned:
    """Keeps a CUDA-resident entropy model on its GPU when the host code moves it to the CPU, and lets it take CPU
    tensors.  DC-VIC's ``codec_setup`` does ``entropy_model_z.to("cpu")`` / ``entropy_model_y.to("cpu")`` and then
    feeds ``y.cpu()``, ``z.cpu()`` (hyperprior_dc_vic_model.py:65-73,308-328): with this mixin the call sites stay
    as they are, the arithmetic still runs in the CUDA kernels (there is no CPU implementation in this package),
    inputs are uploaded and results returned on the caller's device.  Set ``allow_cpu_move = True`` on an instance to
    get torch's normal behaviour back (its forward will then refuse to run)."""

    allow_cpu_move = False

    def _apply(self, fn, *args, **kwargs):
        if not self.allow_cpu_move:
            dev = _module_device(self)
            if dev.type == "cuda":
                try:
                    target = fn(torch.empty(0, device=dev)).device
                except Exception:      # not a device/dtype conversion: let torch handle it
                    target = dev
                if target.type == "cpu":
                    return self
        return super()._apply(fn, *args, **kwargs)

    def _upload(self, *tensors):
        """-> (device the caller works on, the tensors on this module's GPU)."""
        dev = _module_device(self)
        home = next((t.device for t in tensors if t is not None), dev)
        if dev.type != "cuda":
            return home, tensors
        return home, tuple(None if t is None else (t if t.device == dev else t.to(dev)) for t in tensors)

    @staticmethod
    def _download(home, *tensors):
        return tuple(t if (t is None or t.device == home) else t.to(home) for t in tensors)


class EntropyModel(DevicePinned, nn.Module):
    """compressai.entropy_models.EntropyModel surface used by DC-VIC (base_model.py:88-104)."""

    def __init__(self, likelihood_bound: float = 1e-9, entropy_coder=None, entropy_coder_precision: int = 16):
        super().__init__()
        self.entropy_coder = entropy_coder
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self._likelihood_bound = float(likelihood_bound) if self.use_likelihood_bound else 0.0
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    def quantize(self, inputs: Tensor, mode: str, means: Optional[Tensor] = None) -> Tensor:
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        home, (inputs, means) = self._upload(inputs, means)
        _lib.require_cuda(inputs)
        if mode == "noise":
            noise = torch.empty_like(inputs).uniform_(-0.5, 0.5)
            return self._download(home, inputs + noise)[0]
        x = inputs if means is None else inputs - means
        r = _SteRound.apply(x.detach())          # value == round(x), one kernel
        if mode == "dequantize":
            return self._download(home, r if means is None else r + means)[0]
        return self._download(home, r.int())[0]

    @staticmethod
    def dequantize(inputs: Tensor, means: Optional[Tensor] = None, dtype: torch.dtype = torch.float) -> Tensor:
        if means is not None:
            outputs = inputs.type_as(means)
            outputs += means
        else:
            outputs = inputs.type(dtype)
        return outputs

    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        if pmf.is_cuda:
            # the whole table on the device, one thread per row (dcvic_pmf_to_quantized_cdf_rows)
            dev = pmf.device
            pm = pmf.detach().float().contiguous()
            rows, width = pm.shape[0], pm.shape[1]
            tm = tail_mass.detach().float().reshape(-1).contiguous().to(dev)
            ln = pmf_length.detach().int().reshape(-1).contiguous().to(dev)
            with _lib.on_device(dev):
                cdf = torch.empty(rows, width + 2, dtype=torch.int32, device=dev)
                status = torch.zeros(1, dtype=torch.int32, device=dev)
                rc = _lib.load().dcvic_pmf_to_quantized_cdf_rows(_lib.ptr(pm), rows, width, _lib.ptr(tm), _lib.ptr(ln),
                                                                 self.entropy_coder_precision, _lib.ptr(cdf),
                                                                 _lib.ptr(status), _lib.cur_stream())
                _lib.check(rc, "dcvic_pmf_to_quantized_cdf_rows")
            _lib.check(int(status), "dcvic_pmf_to_quantized_cdf_rows (row)")
            return cdf[:, : max_length + 2].contiguous()
        pmf, tail_mass, pmf_length = pmf.cpu(), tail_mass.cpu(), pmf_length.cpu()
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
        for i, p in enumerate(pmf):
            prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
            row = pmf_to_quantized_cdf(prob.numpy(), self.entropy_coder_precision)
            cdf[i, : row.numel()] = row
        return cdf

    # ---- bitstreams (CompressAI 1.2.4 EntropyModel.compress / decompress; call sites
    # hyperprior_dc_vic_model.py:308-328,378-387).  Symbols are produced and range-coded on the GPU (dc_vic_b200.rans:
    # one warp per image, no .tolist() marshaling); only the finished byte strings travel to the host.
    def _check_tables(self):
        if self._quantized_cdf.numel() == 0 or self._offset.numel() == 0 or self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if self._quantized_cdf.dim() != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")

    def compress(self, inputs: Tensor, indexes: Tensor, means: Optional[Tensor] = None):
        from . import rans
        if inputs.dim() < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if inputs.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        self._check_tables()
        _, (inputs, indexes, means) = self._upload(inputs, indexes, means)
        symbols = self.quantize(inputs, "symbols", means)
        dev = symbols.device
        tab = rans._Tables(self._quantized_cdf.to(dev), self._cdf_length.reshape(-1).to(dev),
                           self._offset.reshape(-1).to(dev), dev)
        B = symbols.shape[0]
        return rans.encode_batch([symbols[i].reshape(-1) for i in range(B)],
                                 [indexes[i].reshape(-1).int() for i in range(B)], tab)

    def decompress(self, strings, indexes: Tensor, dtype: torch.dtype = torch.float, means: Optional[Tensor] = None):
        from . import rans
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if indexes.dim() < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        self._check_tables()
        home, (indexes, means) = self._upload(indexes, means)
        dev = indexes.device
        outputs = torch.empty(indexes.size(), dtype=torch.int32, device=dev)
        for i, s in enumerate(strings):
            dec = rans.RansDecoder()
            dec.set_stream(s)
            outputs[i] = dec.decode_stream_tensor(indexes[i].reshape(-1).int(), self._quantized_cdf.to(dev),
                                                  self._cdf_length.reshape(-1).to(dev),
                                                  self._offset.reshape(-1).to(dev)).view(outputs[i].shape)
        return self._download(home, self.dequantize(outputs, means, dtype))[0]


# ------------------------------------------------------------------------------ Gaussian
class _GaussianFn(torch.autograd.Function):
    """Fused quantize + likelihood (+ both LowerBounds).  Outputs (y_hat, likelihood)."""

    @staticmethod
    def forward(ctx, y, scales, means, noise, scale_bound, lik_bound, y_hat_mode):
        _lib.require_cuda(y, scales, means, noise)
        lib = _lib.load()
        B = y.shape[0]
        n = y.numel() // B
        yv, ys = _batch_view(y.detach(), B, n)
        sv, ss = _batch_view(scales.detach(), B, n)
        mv, ms = (None, 0) if means is None else _batch_view(means.detach(), B, n)
        nz = None if noise is None else noise.detach().contiguous().float()
        with _lib.on_device(y.device):
            y_hat = torch.empty(y.shape, dtype=torch.float32, device=y.device)
            lik = torch.empty(y.shape, dtype=torch.float32, device=y.device)
            rc = lib.dcvic_gc_forward(_lib.ptr(yv), _lib.ptr(mv), _lib.ptr(sv), _lib.ptr(nz), B, n, ys, ms, ss,
                                      float(scale_bound), float(lik_bound), int(y_hat_mode), _lib.ptr(y_hat),
                                      _lib.ptr(lik), None, None, 0, _lib.cur_stream())
            _lib.check(rc, "dcvic_gc_forward")
        ctx.save_for_backward(yv, sv, mv, nz)
        ctx.meta = (B, n, ys, ms, ss, float(scale_bound), float(lik_bound), int(y_hat_mode), y.shape,
                    scales.shape, None if means is None else means.shape)
        return y_hat, lik

    @staticmethod
    def backward(ctx, g_yhat, g_lik):
        yv, sv, mv, nz = ctx.saved_tensors
        B, n, ys, ms, ss, scale_bound, lik_bound, mode, y_shape, s_shape, m_shape = ctx.meta
        lib = _lib.load()
        dev = yv.device
        with _lib.on_device(dev):
            d_y = torch.empty(y_shape, dtype=torch.float32, device=dev)
            d_s = torch.empty(y_shape, dtype=torch.float32, device=dev)
            d_m = torch.empty(y_shape, dtype=torch.float32, device=dev) if mv is not None else None
            gl = (g_lik if g_lik is not None else torch.zeros(y_shape, device=dev)).contiguous().float()
            rc = lib.dcvic_gc_backward(_lib.ptr(gl), _lib.ptr(yv), _lib.ptr(mv), _lib.ptr(sv), _lib.ptr(nz), B, n,
                                       ys, ms, ss, scale_bound, lik_bound, _lib.ptr(d_y), _lib.ptr(d_m),
                                       _lib.ptr(d_s), _lib.cur_stream())
            _lib.check(rc, "dcvic_gc_backward")
        if g_yhat is not None:
            training = nz is not None
            if mode == 0 and training:          # y_hat = y + noise
                d_y = d_y + g_yhat
            elif mode == 1 and training:        # y_hat = ste_round(y - mu) + mu
                d_y = d_y + g_yhat
            elif d_m is not None:               # eval: y_hat = round(y - mu) + mu
                d_m = d_m + g_yhat
        return (d_y, d_s.view(s_shape) if d_s.numel() == math.prod(s_shape) else d_s,
                d_m, None, None, None, None)


class GaussianConditional(EntropyModel):
    """compressai.entropy_models.GaussianConditional (1.2.4)."""

    def __init__(self, scale_table, *args, scale_bound: float = 0.11, tail_mass: float = 1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = self.scale_table[0]
        if scale_bound <= 0:       # scale_bound=None -> TypeError, exactly like 1.2.4
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self._scale_bound = float(scale_bound)
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    def _standardized_quantile(self, quantile):
        from scipy.stats import norm
        return norm.ppf(quantile)

    def _likelihood(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None) -> Tensor:
        """Un-bounded likelihood of already-quantized ``inputs`` (used by update())."""
        zero = torch.zeros_like(inputs)
        _, lik = _GaussianFn.apply(inputs, scales, means, zero, self._scale_bound, 0.0, _lib.GC_PRECISE)
        return lik

    def forward(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                training: Optional[bool] = None, noise: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        if training is None:
            training = self.training
        if training and noise is None:
            noise = torch.empty_like(inputs).uniform_(-0.5, 0.5)   # same RNG call as EntropyModel.quantize
        if not training:
            noise = None
        return _GaussianFn.apply(inputs, scales, means, noise, self._scale_bound, self._likelihood_bound, 0)

    def build_indexes(self, scales: Tensor) -> Tensor:
        home, (scales,) = self._upload(scales)
        _lib.require_cuda(scales)
        sc = scales.contiguous().float()
        table = self.scale_table.to(sc.device).contiguous().float()
        with _lib.on_device(sc.device):
            out = torch.empty(sc.shape, dtype=torch.int32, device=sc.device)
            rc = _lib.load().dcvic_gc_build_indexes(_lib.ptr(sc), sc.numel(), _lib.ptr(table), table.numel(),
                                                    self._scale_bound, _lib.ptr(out), _lib.cur_stream())
            _lib.check(rc, "dcvic_gc_build_indexes")
        return self._download(home, out)[0]

    def update_scale_table(self, scale_table, force: bool = False) -> bool:
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        self.update()
        return True

    def update(self):
        device = self.scale_table.device
        multiplier = -self._standardized_quantile(self.tail_mass / 2)
        pmf_center = torch.ceil(self.scale_table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = int(torch.max(pmf_length).item())
        samples = torch.abs(torch.arange(max_length, device=device).int() - pmf_center[:, None]).float()
        samples_scale = self.scale_table.unsqueeze(1).float().expand_as(samples).contiguous()
        # pmf = Phi((.5 - s)/sigma) - Phi((-.5 - s)/sigma): the likelihood kernel at integer samples
        pmf = self._likelihood(samples.contiguous(), samples_scale)
        # tail_mass = 2 * lower[:, :1] with lower = Phi((-.5 - s)/sigma) at the first sample
        # (setup-time, 64 values: plain torch on the table's device, FP32 like the reference)
        lower0 = 0.5 * torch.erfc(-(2 ** -0.5) * ((-0.5 - samples[:, :1]) / self.scale_table.unsqueeze(1).float()))
        tail_mass = 2 * lower0
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length).to(device)
        self._offset = -pmf_center
        self._cdf_length = pmf_length + 2


def gaussian_rate_dual(y: Tensor, params: Tensor, noise: Tensor, scale_bound: float = 0.11,
                       lik_bound: float = 1e-9, want_bits: bool = True):
    """One pass over a CHARM slice producing what the reference gets from two entropy-model
    calls (minnen20_charm_context_model.py:96-101): STE ``y_hat``, noisy and quantized likelihood
    and, optionally, both per-sample bit sums.  Forward-only (the reference runs the quantized
    branch under no_grad and the fix_entropy_models stages run the whole path without grad)."""
    _lib.require_cuda(y, params, noise)
    lib = _lib.load()
    B = y.shape[0]
    n = y.numel() // B
    mean, std = params.chunk(2, 1)
    yv, ys = _batch_view(y.detach(), B, n)
    mv, ms = _batch_view(mean.detach(), B, n)
    sv, ss = _batch_view(std.detach(), B, n)
    nz = noise.detach().contiguous().float()
    dev = y.device
    with _lib.on_device(dev):
        y_hat = torch.empty(y.shape, dtype=torch.float32, device=dev)
        lik = torch.empty_like(y_hat)
        lik_q = torch.empty_like(y_hat)
        bits = torch.empty(B, dtype=torch.float32, device=dev) if want_bits else None
        bits_q = torch.empty(B, dtype=torch.float32, device=dev) if want_bits else None
        ws = _workspace("gc", lib.dcvic_gc_workspace_bytes(B, n), dev)
        rc = lib.dcvic_gc_forward_dual(_lib.ptr(yv), _lib.ptr(mv), _lib.ptr(sv), _lib.ptr(nz), B, n, ys, ms, ss,
                                       float(scale_bound), float(lik_bound), _lib.ptr(y_hat), _lib.ptr(lik),
                                       _lib.ptr(lik_q), _lib.ptr(bits), _lib.ptr(bits_q), _lib.ptr(ws), ws.numel(),
                                       _lib.cur_stream())
        _lib.check(rc, "dcvic_gc_forward_dual")
    return y_hat, lik, lik_q, bits, bits_q


def gaussian_codec_step(y: Tensor, params: Tensor, scale_table: Tensor, scale_bound: float = 0.11,
                        lik_bound: float = 1e-9):
    """Compress-side step of one CHARM slice in one pass: what the reference computes with
    ``entropy_model_y(y_slice, cat([mu, sigma]), is_train=False)`` (minnen20_charm_context_model.py:146) plus this
    slice's share of ``build_indexes(y_scale)`` (:164) and of ``quantize(y, "symbols", means)`` inside ``compress``
    (:165).  Returns ``(y_hat, likelihood, symbols int32, indexes int32)``, all shaped like ``y``."""
    _lib.require_cuda(y, params, scale_table)
    lib = _lib.load()
    B = y.shape[0]
    n = y.numel() // B
    mean, std = params.chunk(2, 1)
    yv, ys = _batch_view(y.detach(), B, n)
    mv, ms = _batch_view(mean.detach(), B, n)
    sv, ss = _batch_view(std.detach(), B, n)
    table = scale_table.detach().to(y.device).contiguous().float()
    dev = y.device
    with _lib.on_device(dev):
        y_hat = torch.empty(y.shape, dtype=torch.float32, device=dev)
        lik = torch.empty_like(y_hat)
        symbols = torch.empty(y.shape, dtype=torch.int32, device=dev)
        indexes = torch.empty(y.shape, dtype=torch.int32, device=dev)
        rc = lib.dcvic_gc_codec_step(_lib.ptr(yv), _lib.ptr(mv), _lib.ptr(sv), B, n, ys, ms, ss, _lib.ptr(table),
                                     table.numel(), float(scale_bound), float(lik_bound), _lib.ptr(y_hat),
                                     _lib.ptr(lik), _lib.ptr(symbols), _lib.ptr(indexes), _lib.cur_stream())
        _lib.check(rc, "dcvic_gc_codec_step")
    return y_hat, lik, symbols, indexes


# ------------------------------------------------------------------------------ bottleneck
class _BottleneckFn(torch.autograd.Function):
    N_PARAMS = 14

    @staticmethod
    def forward(ctx, x, noise, lik_bound, x_hat_mode, quantiles, *params):
        _lib.require_cuda(x, noise, quantiles, *params)
        lib = _lib.load()
        xc = x.detach().contiguous().float()
        B, Cc = xc.shape[0], xc.shape[1]
        HW = xc.numel() // (B * Cc)
        nz = None if noise is None else noise.detach().contiguous().float()
        ps = [p.detach().contiguous().float() for p in params] + [quantiles.detach().contiguous().float()]
        arr = (C.c_void_p * 15)(*[p.data_ptr() for p in ps])
        with _lib.on_device(x.device):
            x_hat = torch.empty_like(xc)
            lik = torch.empty_like(xc)
            rc = lib.dcvic_eb_forward(_lib.ptr(xc), _lib.ptr(nz), arr, B, Cc, HW, float(lik_bound), int(x_hat_mode),
                                      _lib.ptr(x_hat), _lib.ptr(lik), None, None, 0, _lib.cur_stream())
            _lib.check(rc, "dcvic_eb_forward")
        ctx.save_for_backward(xc, nz, *ps)
        ctx.meta = (B, Cc, HW, float(lik_bound), int(x_hat_mode))
        return x_hat, lik

    @staticmethod
    def backward(ctx, g_xhat, g_lik):
        xc, nz, *ps = ctx.saved_tensors
        B, Cc, HW, lik_bound, mode = ctx.meta
        lib = _lib.load()
        dev = xc.device
        n_par = _BottleneckFn.N_PARAMS
        grads = [None] * n_par
        d_x = None
        if nz is not None and g_lik is not None:
            with _lib.on_device(dev):
                d_x = torch.empty_like(xc)
                grads = [torch.empty_like(p) for p in ps[:n_par]]
                parr = (C.c_void_p * 15)(*[p.data_ptr() for p in ps])
                garr = (C.c_void_p * 14)(*[g.data_ptr() for g in grads])
                ws = _workspace("eb", lib.dcvic_eb_workspace_bytes(B, Cc, HW), dev)
                gl = g_lik.contiguous().float()
                rc = lib.dcvic_eb_backward(_lib.ptr(gl), _lib.ptr(xc), _lib.ptr(nz), parr, B, Cc, HW, lik_bound,
                                           _lib.ptr(d_x), garr, _lib.ptr(ws), ws.numel(), _lib.cur_stream())
                _lib.check(rc, "dcvic_eb_backward")
        elif g_lik is not None:
            raise NotImplementedError("EntropyBottleneck eval-mode likelihood gradients are not implemented "
                                      "(the reference evaluates that branch under torch.no_grad())")
        if g_xhat is not None and nz is not None:     # x_hat = x + noise  |  ste_round(x - med) + med
            d_x = g_xhat if d_x is None else d_x + g_xhat
        # quantiles: the medians only shift the eval-mode rounding -> no gradient (as in CompressAI training mode)
        return (d_x, None, None, None, None, *grads)


class EntropyBottleneck(EntropyModel):
    """compressai.entropy_models.EntropyBottleneck (1.2.4), filters=(3,3,3,3)."""

    def __init__(self, channels: int, *args, tail_mass: float = 1e-9, init_scale: float = 10,
                 filters: Tuple[int, ...] = (3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        if self.filters != (3, 3, 3, 3):
            raise NotImplementedError("the sm_100a EntropyBottleneck kernels cover filters=(3,3,3,3) "
                                      "(the only configuration DC-VIC uses)")
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        channels = self.channels
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            matrix = torch.Tensor(channels, filters[i + 1], filters[i])
            matrix.data.fill_(init)
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(matrix))
            bias = torch.Tensor(channels, filters[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))
            if i < len(self.filters):
                factor = torch.Tensor(channels, filters[i + 1], 1)
                nn.init.zeros_(factor)
                self.register_parameter(f"_factor{i:d}", nn.Parameter(factor))
        self.quantiles = nn.Parameter(torch.Tensor(channels, 1, 3))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles.data = init.repeat(self.quantiles.size(0), 1, 1)
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def _get_medians(self) -> Tensor:
        return self.quantiles[:, :, 1:2]

    def _param_list(self, detach: bool = False):
        ps = [getattr(self, f"_matrix{i}") for i in range(5)] + [getattr(self, f"_bias{i}") for i in range(5)] + \
             [getattr(self, f"_factor{i}") for i in range(4)]
        return [p.detach() for p in ps] if detach else ps

    def _logits_cumulative(self, inputs: Tensor, stop_gradient: bool) -> Tensor:
        """Tiny per-channel MLP on [C,1,M] inputs; used by loss()/update() on 3 / ~O(30) points
        per channel, so it stays in torch (stop_gradient=True everywhere DC-VIC calls it)."""
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = getattr(self, f"_matrix{i:d}")
            bias = getattr(self, f"_bias{i:d}")
            if stop_gradient:
                matrix, bias = matrix.detach(), bias.detach()
            logits = torch.matmul(torch.nn.functional.softplus(matrix), logits) + bias
            if i < len(self.filters):
                factor = getattr(self, f"_factor{i:d}")
                if stop_gradient:
                    factor = factor.detach()
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def forward(self, x: Tensor, training: Optional[bool] = None, noise: Optional[Tensor] = None,
                _x_hat_mode: int = 0) -> Tuple[Tensor, Tensor]:
        if training is None:
            training = self.training
        if training and noise is None:
            noise = _bottleneck_noise(x)
        if not training:
            noise = None
        return _BottleneckFn.apply(x, noise, self._likelihood_bound, _x_hat_mode, self.quantiles,
                                   *self._param_list())

    def loss(self) -> Tensor:
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    @staticmethod
    def _build_indexes(size):
        dims = len(size)
        N, Cc = size[0], size[1]
        view_dims = np.ones((dims,), dtype=np.int64)
        view_dims[1] = -1
        indexes = torch.arange(Cc).view(*view_dims)
        return indexes.int().repeat(N, 1, *size[2:])

    @staticmethod
    def _extend_ndims(tensor, n):
        return tensor.reshape(-1, *([1] * n)) if n > 0 else tensor.reshape(-1)

    def compress(self, x: Tensor):
        indexes = self._build_indexes(x.size()).to(self._quantized_cdf.device)
        medians = self._get_medians().detach()
        spatial_dims = len(x.size()) - 2
        medians = self._extend_ndims(medians, spatial_dims)
        medians = medians.expand(x.size(0), *([-1] * (spatial_dims + 1)))
        return super().compress(x, indexes, medians)

    def decompress(self, strings, size):
        output_size = (len(strings), self._quantized_cdf.size(0), *size)
        indexes = self._build_indexes(output_size).to(self._quantized_cdf.device)
        medians = self._extend_ndims(self._get_medians().detach(), len(size))
        medians = medians.expand(len(strings), *([-1] * (len(size) + 1)))
        return super().decompress(strings, indexes, medians.dtype, medians)

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
        device = pmf_start.device
        samples = torch.arange(max_length, device=device)[None, :] + pmf_start[:, None, None]
        with torch.no_grad():
            lower = self._logits_cumulative(samples - 0.5, stop_gradient=True)
            upper = self._logits_cumulative(samples + 0.5, stop_gradient=True)
            sign = -torch.sign(lower + upper)
            pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
            tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length).to(device)
        self._cdf_length = pmf_length + 2
        return True


# ------------------------------------------------------------- DC-VIC wrappers (registry names)
# The fused forwards are MIXINS without state of their own: they only need the attributes every CompressAI-shaped
# module has (parameters ``_matrix{i}`` / ``_bias{i}`` / ``_factor{i}`` / ``quantiles``, ``likelihood_lower_bound``,
# ``lower_bound_scale``).  ``register.register_entropy_models`` puts them IN FRONT of the reference's own wrapper
# classes (``type(name, (mixin, reference_class), {})``), so that the class built from ``ENTROPYMODEL_REGISTRY`` is a
# subclass of ``src.models.subnet.entropy_model.entropy_bottleneck.EntropyBottleneck`` /
# ``compressai.entropy_models.GaussianConditional`` and ``base_model.py:76-104,128-130``'s isinstance checks hold;
# below they are combined with this file's own base classes for stand-alone use.
def _bound_of(module, cache_name: str, private_name: str, bound_module_name: str, default: float) -> float:
    v = getattr(module, private_name, None)
    if v is not None:
        return v
    v = module.__dict__.get(cache_name)
    if v is None:
        lb = getattr(module, bound_module_name, None)
        v = float(lb.bound) if lb is not None else default       # (one host read, then cached)
        module.__dict__[cache_name] = v
    return v


def _lik_bound_of(module) -> float:
    if not getattr(module, "use_likelihood_bound", True):
        return 0.0
    return _bound_of(module, "_dcvic_lik_bound", "_likelihood_bound", "likelihood_lower_bound", 1e-9)


def _scale_bound_of(module) -> float:
    return _bound_of(module, "_dcvic_scale_bound", "_scale_bound", "lower_bound_scale", 0.11)


def _bottleneck_noise(x: Tensor) -> Tensor:
    # the reference draws the noise on the permuted [C, 1, B*HW] view; draw it there so the RNG
    # stream is consumed identically, then bring it back to NCHW
    Cc, B = x.shape[1], x.shape[0]
    nz = torch.empty(Cc, 1, x.numel() // Cc, device=x.device, dtype=torch.float32).uniform_(-0.5, 0.5)
    return nz.view(Cc, B, *x.shape[2:]).transpose(0, 1).contiguous()


def _bottleneck_params(module):
    return [getattr(module, f"_matrix{i}") for i in range(5)] + [getattr(module, f"_bias{i}") for i in range(5)] + \
           [getattr(module, f"_factor{i}") for i in range(4)]


class FusedBottleneckForward(DevicePinned):
    """``EntropyBottleneck.forward(x, is_train)`` of entropy_bottleneck.py:13-16 as one kernel."""

    _x_hat_mode = 0          # 0: x + noise (training) / round about the median (eval)

    def forward(self, x: Tensor, is_train: bool = True, noise: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        home, (x, noise) = self._upload(x, noise)
        if is_train and noise is None:
            noise = _bottleneck_noise(x)
        if not is_train:
            noise = None
        out = _BottleneckFn.apply(x, noise, _lik_bound_of(self), self._x_hat_mode if is_train else 0,
                                  self.quantiles, *_bottleneck_params(self))
        return self._download(home, *out)


class FusedSteBottleneckForward(FusedBottleneckForward):
    """entropy_bottleneck.py:19-28: the training output is ``ste_round(x - med) + med``."""

    _x_hat_mode = 1


class _FusedGaussianForward(DevicePinned):
    _y_hat_mode = 0          # 0: y + noise (training); 1: ste_round(y - mu) + mu (training).  Eval: round about mu.
    _has_means = True

    def forward(self, y: Tensor, params: Tensor, is_train: bool = True,
                noise: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        home, (y, params, noise) = self._upload(y, params, noise)
        if self._has_means:
            mean, std = params.chunk(2, 1)
        else:
            mean, std = None, params
        if is_train and noise is None:
            noise = torch.empty_like(y).uniform_(-0.5, 0.5)        # same RNG call as EntropyModel.quantize("noise")
        if not is_train:
            noise = None
        out = _GaussianFn.apply(y, std, mean, noise, _scale_bound_of(self), _lik_bound_of(self), self._y_hat_mode)
        return self._download(home, *out)


class FusedGaussianScaleForward(_FusedGaussianForward):
    """gaussian_conditional.py:9-15 (``forward(y, scales, is_train)``, no means)."""

    _has_means = False


class FusedGaussianMeanScaleForward(_FusedGaussianForward):
    """gaussian_conditional.py:17-24."""


class FusedSteGaussianMeanScaleForward(_FusedGaussianForward):
    """ste_gaussian_conditional.py:9-23: y_hat is the STE-rounded value in training, the plain de-quantized value in
    eval; the likelihood is the parent's (noisy in training)."""

    _y_hat_mode = 1


FUSED_FORWARDS = {
    "EntropyBottleneck": FusedBottleneckForward,
    "SteEntropyBottleneck": FusedSteBottleneckForward,
    "GaussianScaleConditional": FusedGaussianScaleForward,
    "GaussianMeanScaleConditional": FusedGaussianMeanScaleForward,
    "SteGaussianMeanScaleConditional": FusedSteGaussianMeanScaleForward,
}


class DcvicEntropyBottleneck(FusedBottleneckForward, EntropyBottleneck):
    """``EntropyBottleneck`` of entropy_bottleneck.py:13-16 (forward(x, is_train))."""


class SteEntropyBottleneck(FusedSteBottleneckForward, DcvicEntropyBottleneck):
    """entropy_bottleneck.py:19-28."""


class GaussianScaleConditional(FusedGaussianScaleForward, GaussianConditional):
    """gaussian_conditional.py:9-15."""

    def __init__(self, scale_bound=None):
        super().__init__(scale_table=None, scale_bound=scale_bound)


class GaussianMeanScaleConditional(FusedGaussianMeanScaleForward, GaussianConditional):
    """gaussian_conditional.py:17-24."""

    def __init__(self, scale_bound=None):
        super().__init__(scale_table=None, scale_bound=scale_bound)


class SteGaussianMeanScaleConditional(FusedSteGaussianMeanScaleForward, GaussianMeanScaleConditional):
    """ste_gaussian_conditional.py:9-23."""

    def __init__(self, scale_bound=None, entropy_quant_type: str = "noise", **kwargs) -> None:
        super().__init__(scale_bound=scale_bound)
        assert entropy_quant_type == "noise"
        self.entropy_quant_type = entropy_quant_type
