"""B200-native VQGAN codebook quantizers with the reference's module interface.

Mirrors (same ctor arguments, attribute / state-dict names, forward signatures, return tuples):

* ``VectorQuantizer``   -- taming/modules/vqvae/quantize.py:9-107   (V1: one-hot + perplexity)
* ``VectorQuantizer2``  -- taming/modules/vqvae/quantize.py:213-329 (the class DC-VIC uses via
  ldm/models/autoencoder.py:6,39-41; ``sane_index_shape`` is set at
  src/models/comp_model/hyperprior_vic_model.py:61)

All arithmetic runs in hand-written CUDA behind the C ABI of ``include/dcvic_b200.h``
(``dcvic_vq_forward`` / ``dcvic_vq_backward`` / ``dcvic_codebook_gather`` / ``dcvic_onehot_nchw``).
There is no CPU or PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch
import torch.nn as nn

from . import _lib

__all__ = ["VectorQuantizer", "VectorQuantizer2", "codebook_lookup", "onehot_feature", "swap_quantizer", "decode_tokens",
           "CrossEntropyLoss", "FocalCrossEntropyLoss"]


class _Workspace:
    """Zero-initialised device scratch, one buffer per (device, stream) that grows to the largest size asked for
    (evaluating on variable-size images must not keep one buffer per distinct shape alive)."""

    def __init__(self):
        self._cache = {}

    def get(self, nbytes: int, device):
        """-> (buffer, fresh): fresh means newly allocated (zero-filled; nothing prepared in it yet)."""
        stream = torch._C._cuda_getCurrentRawStream(device.index if device.index is not None
                                                    else torch.cuda.current_device())
        k = (device.index, stream)
        buf = self._cache.get(k)
        if buf is None or buf.numel() < nbytes:
            buf = None                      # release the smaller one first
            self._cache.pop(k, None)
            buf = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._cache[k] = buf
            return buf, True
        return buf, False


class _VQForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, weight, beta, legacy, want_v1, flags, owner):
        _lib.require_cuda(z, weight)
        if z.dim() != 4:
            raise ValueError(f"expected z of shape [B, C, H, W], got {tuple(z.shape)}")
        B, D, H, W = z.shape
        K, D2 = weight.shape
        if D != D2:
            raise ValueError(f"z has {D} channels but the codebook has e_dim={D2}")
        lib = _lib.load()
        zc = z.detach().contiguous().float()
        wc = weight.detach().contiguous().float()
        N = B * H * W
        with _lib.on_device(z.device):
            z_q = torch.empty_like(zc)
            idx = torch.empty(N, dtype=torch.int64, device=z.device)
            loss = torch.empty((), dtype=torch.float32, device=z.device)
            onehot = torch.empty(N, K, dtype=torch.float32, device=z.device) if want_v1 else None
            ppl = torch.empty((), dtype=torch.float32, device=z.device) if want_v1 else None
            nbytes = lib.dcvic_vq_workspace_bytes(B, D, H, W, K)
            if nbytes == 0:
                raise RuntimeError("dcvic_vq_workspace_bytes rejected the shape")
            ws, fresh = owner._ws.get(nbytes, z.device)
            # what the prepared |e|^2 / FP16 codebook inside `ws` belongs to (the workspace layout depends on the
            # token count, so a different shape re-prepares)
            key = (wc.data_ptr(), weight._version, ws.data_ptr(), B * H * W, D, K)
            if owner._frozen and not fresh and owner._prep_key == key:
                flags |= _lib.VQ_REUSE_PREP
            rc = lib.dcvic_vq_forward(_lib.ptr(zc), _lib.ptr(wc), B, D, H, W, K, float(beta), int(bool(legacy)),
                                      _lib.ptr(z_q), _lib.ptr(idx), _lib.ptr(loss), _lib.ptr(onehot), _lib.ptr(ppl),
                                      int(flags), _lib.ptr(ws), ws.numel(), _lib.cur_stream())
            _lib.check(rc, "dcvic_vq_forward")
            owner._prep_key = key
        ctx.save_for_backward(zc, wc, idx)
        ctx.meta = (B, D, H, W, K, float(beta), bool(legacy))
        ctx.mark_non_differentiable(idx)
        if want_v1:
            ctx.mark_non_differentiable(onehot, ppl)
            return z_q, loss, idx, onehot, ppl
        return z_q, loss, idx

    @staticmethod
    def backward(ctx, g_zq, g_loss, *unused):
        zc, wc, idx = ctx.saved_tensors
        B, D, H, W, K, beta, legacy = ctx.meta
        lib = _lib.load()
        need_z, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        with _lib.on_device(zc.device):
            dz = torch.empty_like(zc) if need_z else None
            dE = torch.empty_like(wc) if need_w else None
            if dz is None and dE is None:
                return (None,) * 7
            gz = g_zq.contiguous().float() if g_zq is not None else None
            gl = g_loss.contiguous().float() if g_loss is not None else None
            rc = lib.dcvic_vq_backward(_lib.ptr(gz), _lib.ptr(gl), _lib.ptr(zc), _lib.ptr(wc), _lib.ptr(idx),
                                       B, D, H, W, K, beta, int(legacy), _lib.ptr(dz), _lib.ptr(dE),
                                       _lib.cur_stream())
            _lib.check(rc, "dcvic_vq_backward")
        return dz, dE, None, None, None, None, None


_CHECK_INDICES = os.environ.get("DCVIC_CHECK_INDICES", "0") == "1"


class _Gather(torch.autograd.Function):
    """Differentiable wrt the codebook like ``nn.Embedding`` (quantize.py:92-107 / :314-329 go through
    ``self.embedding``); the gradient is a scatter-add of the incoming rows (torch ``index_add_``: DC-VIC's codebook
    is frozen, so this is off the hot path)."""

    @staticmethod
    def forward(ctx, weight, indices, nchw):
        out = _gather_rows(indices, weight, nchw)
        ctx.save_for_backward(indices)
        ctx.meta = (tuple(weight.shape), bool(nchw))
        return out

    @staticmethod
    def backward(ctx, g):
        (indices,) = ctx.saved_tensors
        (K, D), nchw = ctx.meta
        rows = g.permute(0, 2, 3, 1).reshape(-1, D) if nchw else g.reshape(-1, D)
        dE = torch.zeros(K, D, dtype=g.dtype, device=g.device).index_add_(0, indices.reshape(-1).clamp(0, K - 1), rows)
        return dE, None, None


def codebook_lookup(indices: torch.Tensor, weight: torch.Tensor, nchw: bool = True) -> torch.Tensor:
    """``embedding(indices)`` [+ 'b h w c -> b c h w'] in one gather kernel.

    ``indices`` [B, H, W] (nchw=True -> [B, D, H, W]) or any shape (nchw=False -> [..., D]).
    Mirrors vq_indices_to_latent (src/models/comp_model/hyperprior_vic_model.py:165-168).
    Out-of-range indices are clamped by the kernel (``nn.Embedding`` raises); set DCVIC_CHECK_INDICES=1 to read the
    kernel's out-of-range counter back (one host sync) and raise IndexError.
    """
    if weight.requires_grad and torch.is_grad_enabled():
        return _Gather.apply(weight, indices, nchw)
    return _gather_rows(indices, weight, nchw)


def _gather_rows(indices: torch.Tensor, weight: torch.Tensor, nchw: bool) -> torch.Tensor:
    _lib.require_cuda(indices, weight)
    lib = _lib.load()
    K, D = weight.shape
    idx = indices.contiguous().long()
    wc = weight.detach().contiguous().float()
    with _lib.on_device(weight.device):
        bad = torch.zeros(1, dtype=torch.int32, device=weight.device) if _CHECK_INDICES else None
        if nchw:
            if idx.dim() != 3:
                raise ValueError("nchw lookup expects indices of shape [B, H, W]")
            B, H, W = idx.shape
            out = torch.empty(B, D, H, W, dtype=torch.float32, device=weight.device)
            rc = lib.dcvic_codebook_gather(_lib.ptr(idx), _lib.ptr(wc), B, H * W, D, K, 1, _lib.ptr(out),
                                           _lib.ptr(bad), _lib.cur_stream())
        else:
            n = idx.numel()
            out = torch.empty(*idx.shape, D, dtype=torch.float32, device=weight.device)
            rc = lib.dcvic_codebook_gather(_lib.ptr(idx), _lib.ptr(wc), 1, n, D, K, 0, _lib.ptr(out), _lib.ptr(bad),
                                           _lib.cur_stream())
        _lib.check(rc, "dcvic_codebook_gather")
        if bad is not None and int(bad) != 0:
            raise IndexError(f"codebook_lookup: {int(bad)} indices outside [0, {K})")
    return out


def onehot_feature(indices_bhw: torch.Tensor, n_embed: int) -> torch.Tensor:
    """``F.one_hot(idx, K).permute(0,3,1,2).float()`` (hyperprior_vic_model.py:268-271), one kernel."""
    _lib.require_cuda(indices_bhw)
    lib = _lib.load()
    idx = indices_bhw.contiguous().long()
    B, H, W = idx.shape
    with _lib.on_device(idx.device):
        out = torch.empty(B, n_embed, H, W, dtype=torch.float32, device=idx.device)
        rc = lib.dcvic_onehot_nchw(_lib.ptr(idx), B, H * W, int(n_embed), _lib.ptr(out), _lib.cur_stream())
        _lib.check(rc, "dcvic_onehot_nchw")
    return out


def _token_decode(logits, weight, gt, want_latent, pq_w, pq_b, gamma, want_lse, want_loss):
    _lib.require_cuda(logits, weight, gt, pq_w, pq_b)
    if logits.dim() != 4:
        raise ValueError(f"expected logits of shape [B, K, H, W], got {tuple(logits.shape)}")
    lib = _lib.load()
    B, K, H, W = logits.shape
    K2, D = weight.shape
    if K != K2:
        raise ValueError(f"logits have {K} classes but the codebook has {K2} entries")
    lg = logits.detach().contiguous().float()
    wc = weight.detach().contiguous().float()
    gtc = None if gt is None else gt.contiguous().long()
    D_out = D
    if pq_w is not None:
        pq_w = pq_w.detach().reshape(pq_w.shape[0], -1).contiguous().float()
        if pq_w.shape[1] != D:
            raise ValueError(f"post_quant_conv expects {pq_w.shape[1]} input channels, the codebook has e_dim={D}")
        D_out = pq_w.shape[0]
        pq_b = None if pq_b is None else pq_b.detach().contiguous().float()
    with _lib.on_device(lg.device):
        idx = torch.empty(B, H, W, dtype=torch.int64, device=lg.device)
        latent = torch.empty(B, D_out, H, W, dtype=torch.float32, device=lg.device) if want_latent else None
        count = torch.empty(1, dtype=torch.int32, device=lg.device) if gtc is not None else None
        lse = torch.empty(B, H, W, dtype=torch.float32, device=lg.device) if want_lse else None
        sums = torch.empty(2, dtype=torch.float64, device=lg.device) if want_loss else None
        rc = lib.dcvic_token_decode_ex(_lib.ptr(lg), _lib.ptr(wc), _lib.ptr(gtc), B, K, H * W, D, _lib.ptr(pq_w),
                                       _lib.ptr(pq_b), D_out, float(gamma), _lib.ptr(idx), _lib.ptr(latent),
                                       _lib.ptr(count), _lib.ptr(lse), _lib.ptr(sums), _lib.cur_stream())
        _lib.check(rc, "dcvic_token_decode_ex")
    return idx, latent, count, lse, sums, lg, gtc


def decode_tokens(logits: torch.Tensor, weight: torch.Tensor, gt_indices: Optional[torch.Tensor] = None,
                  want_latent: bool = True, post_quant_conv: Optional[nn.Module] = None):
    """Decoder-side token path of DC-VIC in one kernel (hyperprior_dc_vic_model.py:250-260):

        out_vq_indices = torch.argmax(out_vq_logits, dim=1)                       # [B, H, W]
        vq_accuracy    = (out_vq_indices == gt_vq_indices).float().mean()         # if gt_indices is given
        vq_latent      = vq_indices_to_latent(out_vq_indices)                     # [B, D, H, W]
        vq_latent      = vq_model.post_quant_conv(vq_latent)                      # if post_quant_conv (a 1x1 Conv2d)

    Returns ``(indices, latent | None, accuracy | None)``.  No gradient flows through argmax in the reference either
    (the logits are trained through the code CE / MSE losses on ``out_vq_logits`` itself).
    """
    pq_w = pq_b = None
    if post_quant_conv is not None:
        if tuple(post_quant_conv.weight.shape[2:]) != (1, 1):
            raise ValueError("post_quant_conv must be a 1x1 convolution (ldm/models/autoencoder.py:41)")
        pq_w, pq_b = post_quant_conv.weight, post_quant_conv.bias
    B, K, H, W = logits.shape
    idx, latent, count, _, _, _, _ = _token_decode(logits, weight, gt_indices, want_latent, pq_w, pq_b, 0.0, False, False)
    acc = None if count is None else (count.float() / float(B * H * W)).reshape(())
    return idx, latent, acc


class _CodeCE(torch.autograd.Function):
    """Mean (or summed) code cross entropy / focal cross entropy over [B, K, H, W] logits, one pass forward (riding
    on the arg-max scan), one pass backward."""

    @staticmethod
    def forward(ctx, logits, target, gamma, loss_weight, reduction, weight_for_shape):
        idx, _, _, lse, sums, lg, gtc = _token_decode(logits, weight_for_shape, target, False, None, None, gamma, True,
                                                      True)
        B, K, H, W = logits.shape
        n = B * H * W
        scale = loss_weight / n if reduction == "mean" else loss_weight
        ctx.save_for_backward(lg, gtc, lse)
        ctx.meta = (B, K, H * W, float(gamma), float(scale))
        return (sums[1 if gamma != 0.0 else 0] * scale).float()

    @staticmethod
    def backward(ctx, g):
        lg, gtc, lse = ctx.saved_tensors
        B, K, HW, gamma, scale = ctx.meta
        with _lib.on_device(lg.device):
            d = torch.empty_like(lg)
            gl = g.detach().reshape(1).contiguous().float()
            rc = _lib.load().dcvic_token_ce_backward(_lib.ptr(lg), _lib.ptr(gtc), _lib.ptr(lse), _lib.ptr(gl), B, K, HW,
                                                     gamma, scale, _lib.ptr(d), _lib.cur_stream())
            _lib.check(rc, "dcvic_token_ce_backward")
        return d, None, None, None, None, None


class CrossEntropyLoss(nn.Module):
    """src/losses/cross_entropy_loss.py:9-31 (the code CE loss on the vq_estimator logits), default ``ce_kwargs``."""

    def __init__(self, loss_weight: float, ce_kwargs: Optional[dict] = None) -> None:
        super().__init__()
        if ce_kwargs:
            raise NotImplementedError("class weights / label smoothing / ignore_index are not used by DC-VIC's configs")
        self.loss_weight = loss_weight

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        dummy = torch.empty(input.shape[1], 1, device=input.device)
        return _CodeCE.apply(input, target, 0.0, float(self.loss_weight), "mean", dummy)


class FocalCrossEntropyLoss(nn.Module):
    """src/losses/cross_entropy_loss.py:33-52."""

    def __init__(self, loss_weight: float, gamma: float, reduction: str = "mean", **kwargs) -> None:
        super().__init__()
        if kwargs:
            raise NotImplementedError("extra nn.CrossEntropyLoss arguments are not used by DC-VIC's configs")
        if reduction not in ("mean", "sum"):
            raise NotImplementedError("reduction='none' is not implemented (the trainers use 'mean')")
        self.loss_weight, self.gamma, self.reduction = loss_weight, gamma, reduction

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        dummy = torch.empty(input.shape[1], 1, device=input.device)
        if self.gamma == 0:
            return _CodeCE.apply(input, target, 0.0, float(self.loss_weight), self.reduction, dummy)
        return _CodeCE.apply(input, target, float(self.gamma), float(self.loss_weight), self.reduction, dummy)


class _QuantizerBase(nn.Module):
    def _init_common(self, n_e, e_dim, beta):
        self.n_e = n_e
        self.e_dim = e_dim
        self.beta = beta
        self.embedding = nn.Embedding(self.n_e, self.e_dim)
        self.embedding.weight.data.uniform_(-1.0 / self.n_e, 1.0 / self.n_e)
        self._ws = _Workspace()
        self._frozen = False
        self._prep_key = None
        self.search = "auto"   # "auto" | "exact" (FP32 SIMT scan) | "tensor" (tcgen05, error if unsupported)

    def freeze_codebook(self, frozen: bool = True):
        """Declare the codebook constant (DC-VIC always freezes the VQGAN,
        src/trainer/rate_distortion_vq_code_trainer.py:62): |e|^2 and the FP16 copy used by the
        tensor-core search are then prepared once and reused while the weight tensor, its ``_version`` and the
        workspace are unchanged.  In-place writes THROUGH ``weight.data`` (``.data.copy_``, EMA updates, loading a
        state dict into an existing parameter bumps ``_version``, ``.data`` writes do not) are invisible to that
        check: call ``invalidate_codebook()`` after them."""
        self._frozen = bool(frozen)
        self._prep_key = None
        return self

    def invalidate_codebook(self):
        """Forget the prepared copies of a frozen codebook (next forward prepares them again)."""
        self._prep_key = None
        return self

    @property
    def codebook_frozen(self) -> bool:
        return self._frozen

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._prep_key = None

    def _flags(self) -> int:
        return {"auto": 0, "exact": _lib.VQ_FORCE_EXACT, "tensor": _lib.VQ_FORCE_TENSOR}[self.search]

    def search_path(self) -> str:
        code = _lib.load().dcvic_vq_path(self.e_dim, self.n_e, self._flags())
        return {0: "narrow-simt", 1: "exact-simt", 2: "tcgen05"}.get(code, f"unsupported({code})")


class VectorQuantizer(_QuantizerBase):
    """taming ``VectorQuantizer`` (quantize.py:9-107): returns
    ``(z_q, loss, (perplexity, min_encodings[N,K], min_encoding_indices[N,1]))``."""

    def __init__(self, n_e, e_dim, beta):
        super().__init__()
        self._init_common(n_e, e_dim, beta)

    def forward(self, z):
        z_q, loss, idx, onehot, ppl = _VQForward.apply(z, self.embedding.weight, self.beta, True, True,
                                                       self._flags(), self)
        return z_q, loss, (ppl, onehot, idx.unsqueeze(1))

    def get_codebook_entry(self, indices, shape):
        # shape specifying (batch, height, width, channel)
        if shape is not None:
            return codebook_lookup(indices.reshape(shape[0], shape[1], shape[2]), self.embedding.weight, nchw=True)
        return codebook_lookup(indices.reshape(-1), self.embedding.weight, nchw=False)


class VectorQuantizer2(_QuantizerBase):
    """taming ``VectorQuantizer2`` (quantize.py:213-329): returns
    ``(z_q, loss, (None, None, min_encoding_indices))``; indices are [N] or, with
    ``sane_index_shape``, [B, H, W]."""

    def __init__(self, n_e, e_dim, beta, remap=None, unknown_index="random", sane_index_shape=False, legacy=True):
        super().__init__()
        self._init_common(n_e, e_dim, beta)
        self.legacy = legacy
        if remap is not None:
            # quantize.py:221-269 (index remapping onto a "used" subset): no DC-VIC config sets it
            raise NotImplementedError("VectorQuantizer2(remap=...) is not part of the DC-VIC hot path and is not "
                                      "implemented; build the quantizer with remap=None")
        self.remap = None
        self.unknown_index = unknown_index
        self.re_embed = n_e
        self.sane_index_shape = sane_index_shape

    def forward(self, z, temp=None, rescale_logits=False, return_logits=False):
        assert temp is None or temp == 1.0, "Only for interface compatible with Gumbel"
        assert rescale_logits is False, "Only for interface compatible with Gumbel"
        assert return_logits is False, "Only for interface compatible with Gumbel"
        z_q, loss, idx = _VQForward.apply(z, self.embedding.weight, self.beta, self.legacy, False, self._flags(), self)
        if self.sane_index_shape:
            idx = idx.reshape(z_q.shape[0], z_q.shape[2], z_q.shape[3])
        return z_q, loss, (None, None, idx)

    def get_codebook_entry(self, indices, shape):
        # shape specifying (batch, height, width, channel)
        if shape is not None:
            return codebook_lookup(indices.reshape(shape[0], shape[1], shape[2]), self.embedding.weight, nchw=True)
        return codebook_lookup(indices.reshape(-1), self.embedding.weight, nchw=False)

    @classmethod
    def from_reference(cls, ref: nn.Module) -> "VectorQuantizer2":
        """Build from a taming ``VectorQuantizer2`` instance (same codebook tensor values)."""
        new = cls(ref.n_e, ref.e_dim, ref.beta, remap=getattr(ref, "remap", None),
                  unknown_index=getattr(ref, "unknown_index", "random"),
                  sane_index_shape=getattr(ref, "sane_index_shape", False), legacy=getattr(ref, "legacy", True))
        new.embedding.weight.data = ref.embedding.weight.data.clone()
        new.embedding.weight.requires_grad_(ref.embedding.weight.requires_grad)
        if not ref.embedding.weight.requires_grad:
            new.freeze_codebook()      # DC-VIC: vq_model.requires_grad_(False) (rate_distortion_vq_code_trainer.py:62)
        return new.to(ref.embedding.weight.device)


def swap_quantizer(vq_model: nn.Module) -> nn.Module:
    """``vq_model.quantize = VectorQuantizer2.from_reference(vq_model.quantize)`` -- the VQ plugin
    boundary of DC-VIC is this attribute (ldm/models/autoencoder.py:39-41).  A codebook that does not require grad
    (DC-VIC always freezes the VQGAN) is declared frozen, so its |e|^2 / FP16 copies are prepared once."""
    vq_model.quantize = VectorQuantizer2.from_reference(vq_model.quantize)
    return vq_model
