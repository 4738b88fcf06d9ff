"""CPU oracle for the VQGAN codebook quantizer -- TEST INFRASTRUCTURE, NOT PRODUCT.

Restates, as plain functions over CPU FP32 tensors, what the reference's two quantizer
modules compute (all citations into /root/reference):

* ``VectorQuantizer.forward``      taming/modules/vqvae/quantize.py:34-90   ("V1")
* ``VectorQuantizer2.forward``     taming/modules/vqvae/quantize.py:271-312 ("V2", the one
  DC-VIC instantiates: ldm/models/autoencoder.py:6,39-41, legacy=True, beta=0.25)
* ``get_codebook_entry``           quantize.py:92-107 / :314-329
* the one-hot feature              src/models/comp_model/hyperprior_vic_model.py:268-271

The op ORDER matters for index parity and is kept: distances are formed as
``(sum(z^2) + sum(E^2)) - 2 * (z @ E^T)`` in FP32 and ``argmin`` returns the lowest index
among equal minima.  Everything is written with differentiable torch ops so that
``torch.autograd`` on the oracle is also the oracle for the backward pass.

PINNED against ``tests/golden/vq_*.pt``, which ``tests/golden/make_golden.py`` generates by
importing the vendored reference file itself (``tests/test_oracle_vq.py``).
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch
import torch.nn.functional as F


class VQOut(NamedTuple):
    z_q: torch.Tensor                    # [B, D, H, W] (straight-through value)
    loss: torch.Tensor                   # scalar
    perplexity: Optional[torch.Tensor]   # scalar (V1) / None (V2)
    min_encodings: Optional[torch.Tensor]  # [N, K] one-hot fp32 (V1) / None (V2)
    indices: torch.Tensor                # V1: [N, 1]; V2: [N] or [B, H, W]


def default_codebook(n_e: int, e_dim: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Default init of ``embedding.weight``: U(-1/n_e, 1/n_e) (quantize.py:31-32, :229-230)."""
    w = torch.empty(n_e, e_dim, dtype=torch.float32)
    w.uniform_(-1.0 / n_e, 1.0 / n_e, generator=generator)
    return w


def token_rows(z: torch.Tensor) -> torch.Tensor:
    """NCHW -> [N, D] token-major rows (quantize.py:45-46 / :276-277)."""
    b, d, h, w = z.shape
    return z.permute(0, 2, 3, 1).contiguous().view(b * h * w, d)


def distances(rows: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    """FP32 distance matrix in the reference's association (quantize.py:49-51 / :280-282)."""
    zz = (rows ** 2).sum(dim=1, keepdim=True)
    ee = (codebook ** 2).sum(dim=1)
    cross = rows @ codebook.t()
    return zz + ee - 2 * cross


def top2_relative_gap(dist: torch.Tensor) -> torch.Tensor:
    """(d2 - d1) / |d1| per row: the quantity of BASELINE.json's near-tie clause (< 1e-6)."""
    two = torch.topk(dist, 2, dim=1, largest=False).values
    return (two[:, 1] - two[:, 0]) / two[:, 0].abs().clamp_min(1e-30)


def _commitment_loss(zq_rows_shaped, z_nhwc, beta: float, legacy: bool) -> torch.Tensor:
    a = ((zq_rows_shaped.detach() - z_nhwc) ** 2).mean()
    b = ((zq_rows_shaped - z_nhwc.detach()) ** 2).mean()
    # legacy=True keeps taming's "beta on the wrong term" behaviour (quantize.py:291-295)
    return a + beta * b if legacy else beta * a + b


def vq2_forward(z: torch.Tensor, codebook: torch.Tensor, beta: float = 0.25, legacy: bool = True,
                sane_index_shape: bool = False) -> VQOut:
    """``VectorQuantizer2.forward`` (quantize.py:271-312), remap=None."""
    z_nhwc = z.permute(0, 2, 3, 1).contiguous()
    rows = z_nhwc.view(-1, codebook.shape[1])
    idx = torch.argmin(distances(rows, codebook), dim=1)
    picked = F.embedding(idx, codebook).view(z_nhwc.shape)
    loss = _commitment_loss(picked, z_nhwc, beta, legacy)
    ste = z_nhwc + (picked - z_nhwc).detach()
    z_q = ste.permute(0, 3, 1, 2).contiguous()
    if sane_index_shape:
        idx = idx.reshape(z_q.shape[0], z_q.shape[2], z_q.shape[3])
    return VQOut(z_q, loss, None, None, idx)


def vq1_forward(z: torch.Tensor, codebook: torch.Tensor, beta: float = 0.25) -> VQOut:
    """``VectorQuantizer.forward`` (quantize.py:34-90): one-hot matmul gather + perplexity."""
    n_e = codebook.shape[0]
    z_nhwc = z.permute(0, 2, 3, 1).contiguous()
    rows = z_nhwc.view(-1, codebook.shape[1])
    idx = torch.argmin(distances(rows, codebook), dim=1).unsqueeze(1)
    onehot = torch.zeros(idx.shape[0], n_e).to(z_nhwc)
    onehot.scatter_(1, idx, 1)
    picked = (onehot @ codebook).view(z_nhwc.shape)
    loss = _commitment_loss(picked, z_nhwc, beta, legacy=True)
    ste = z_nhwc + (picked - z_nhwc).detach()
    usage = onehot.mean(dim=0)
    perplexity = torch.exp(-(usage * torch.log(usage + 1e-10)).sum())
    return VQOut(ste.permute(0, 3, 1, 2).contiguous(), loss, perplexity, onehot, idx)


def codebook_entry(indices: torch.Tensor, codebook: torch.Tensor, shape=None) -> torch.Tensor:
    """``get_codebook_entry`` (quantize.py:314-329): gather, optional (B,H,W,D) -> NCHW."""
    out = F.embedding(indices.reshape(-1), codebook)
    if shape is not None:
        out = out.view(shape).permute(0, 3, 1, 2).contiguous()
    return out


def indices_to_latent(indices_bhw: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    """``vq_indices_to_latent`` (hyperprior_vic_model.py:165-168)."""
    return F.embedding(indices_bhw, codebook).permute(0, 3, 1, 2).contiguous()


def decode_tokens(logits: torch.Tensor, codebook: torch.Tensor, gt_indices: Optional[torch.Tensor] = None):
    """Decoder-side token path (hyperprior_dc_vic_model.py:250-260): argmax over the class dimension,
    accuracy against the encoder's indices, codebook lookup."""
    idx = torch.argmax(logits, dim=1)
    acc = None if gt_indices is None else (idx == gt_indices).float().mean()
    return idx, indices_to_latent(idx, codebook), acc


def post_quant_latent(idx: torch.Tensor, codebook: torch.Tensor, pq_weight: torch.Tensor, pq_bias) -> torch.Tensor:
    """Lookup followed by ldm VQModel.post_quant_conv, a 1x1 Conv2d (hyperprior_dc_vic_model.py:258-260;
    ldm/models/autoencoder.py:41)."""
    return F.conv2d(indices_to_latent(idx, codebook), pq_weight.reshape(pq_weight.shape[0], -1, 1, 1), pq_bias)


def code_cross_entropy(logits: torch.Tensor, target: torch.Tensor, loss_weight: float = 1.0) -> torch.Tensor:
    """src/losses/cross_entropy_loss.py:9-31 (CrossEntropyLoss, default ce_kwargs)."""
    return loss_weight * F.cross_entropy(logits, target)


def code_focal_cross_entropy(logits: torch.Tensor, target: torch.Tensor, loss_weight: float, gamma: float,
                             reduction: str = "mean") -> torch.Tensor:
    """src/losses/cross_entropy_loss.py:33-52 (FocalCrossEntropyLoss)."""
    ce_loss = F.cross_entropy(logits, target, reduction="none")
    pt = F.softmax(logits, dim=1).gather(1, target.unsqueeze(1)).squeeze(1)
    focal = ((1 - pt) ** gamma) * ce_loss
    val = focal.mean() if reduction == "mean" else (focal.sum() if reduction == "sum" else focal)
    return loss_weight * val


def onehot_feature(indices_bhw: torch.Tensor, n_embed: int) -> torch.Tensor:
    """``onehot_indices`` encoder feature (hyperprior_vic_model.py:268-271): [B,K,H,W] fp32."""
    return F.one_hot(indices_bhw, num_classes=n_embed).permute(0, 3, 1, 2).float()


def allowed_index_mismatch(z: torch.Tensor, codebook: torch.Tensor, idx_test: torch.Tensor,
                           rel_gap: float = 1e-6, chunk: int = 8192):
    """Apply BASELINE.json's parity rule.

    Returns (n_mismatch, n_outside_clause, n_near_tie_rows).  A mismatching row is *inside*
    the clause when the oracle distance of the index under test is within ``rel_gap``
    (relative) of the oracle minimum, i.e. the two candidates are a documented near-tie.
    """
    rows = token_rows(z)
    test = idx_test.reshape(-1).long().cpu()
    n_mis = n_out = n_tie = 0
    for s in range(0, rows.shape[0], chunk):
        d = distances(rows[s:s + chunk], codebook)
        best, arg = d.min(dim=1)
        t = test[s:s + chunk]
        mis = arg != t
        d_t = d.gather(1, t[:, None]).squeeze(1)
        gap = (d_t - best) / best.abs().clamp_min(1e-30)
        n_mis += int(mis.sum())
        n_out += int((mis & ~(gap < rel_gap)).sum())
        n_tie += int((top2_relative_gap(d) < rel_gap).sum())
    return n_mis, n_out, n_tie
