"""Batched tiling driver for images beyond 1024 px.

The reference handles large images by serial tiling on one device: 512 x 512 windows with stride 256, the network
runs on ONE window per Python-loop iteration, and only the central part of every result is kept
(``_vq_encode_split`` src/models/comp_model/hyperprior_vic_model.py:190-246; ``decode_split`` :413-473, which
stitches on the CPU).  ``_vq_quantize_split`` (:170-188) exists only to cap the N x K distance matrix and is not
needed at all here: the quantizer never materialises that matrix, so the whole latent is quantized in one call.

``TilePlan`` reproduces the reference's window / keep-window arithmetic; ``gather`` puts the windows of all tiles into
one batch (one copy kernel), the caller's network runs once (or in a few chunks) on that batch, ``stitch`` writes every
tile's keep-window into the full-size output (one copy kernel, on the device)."""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch

from . import _lib

__all__ = ["TilePlan", "encode_split", "decode_split", "SPLIT_WINDOW_SIZE", "SPLIT_STRIDE"]

SPLIT_WINDOW_SIZE, SPLIT_STRIDE = 512, 256        # hyperprior_vic_model.py:25-26


def _starts(size: int, patch: int, stride: int) -> List[int]:
    """left_list / top_list of the reference (:196-212, :421-437)."""
    out = []
    for lo in range(0, size + stride, stride):
        if lo + patch < size:
            out.append(lo)
        else:
            out.append(size - patch)
            break
    return out


class TilePlan:
    """Windows of `patch` with `stride` over an H x W input (input coordinates), and for an output that is the input
    scaled by num/den (VQGAN encoder: 1/8; decoder from y_hat: 16/1) the keep-window of every tile."""

    def __init__(self, H: int, W: int, patch: int, stride: int, num: int = 1, den: int = 1):
        if patch > H or patch > W:
            raise ValueError(f"{H}x{W} is smaller than one {patch}x{patch} window: call the network directly")
        if (patch * num) % den or (stride // 2 * num) % den or (stride * num) % den:
            raise ValueError("window / stride do not map onto whole output pixels")
        self.H, self.W, self.patch, self.stride, self.num, self.den = H, W, patch, stride, num, den
        tops, lefts = _starts(H, patch, stride), _starts(W, patch, stride)
        sc = lambda v: v * num // den                     # noqa: E731
        self.out_H, self.out_W, self.out_patch = sc(H), sc(W), sc(patch)
        off, keep = sc(stride // 2), sc(stride)
        self.origins: List[Tuple[int, int]] = []
        self.windows: List[Tuple[int, int, int, int, int, int]] = []

        def spans(starts, out_size):
            """Keep-range [lo, hi) of every window start along one axis (:224-241 / :451-466).  The last window is
            pulled back to fit (start = size - patch), so its range can overlap its predecessor's: the reference's loop
            simply overwrites, i.e. the LATER window wins - the earlier range is cut where the next one begins, which
            makes the ranges a partition and lets one kernel write them all at once."""
            out = []
            for s0 in starts:
                o = sc(s0)
                out.append([o + off if s0 > 0 else 0, o + off + keep if s0 < starts[-1] else out_size])
            for i in range(len(out) - 1):
                out[i][1] = min(out[i][1], out[i + 1][0])
            return out

        rows, cols = spans(tops, self.out_H), spans(lefts, self.out_W)
        for y0, (t, b) in zip(tops, rows):
            for x0, (l, r) in zip(lefts, cols):
                self.origins.append((y0, x0))
                self.windows.append((sc(y0), sc(x0), t, b, l, r))
        self.T = len(self.origins)
        self._dev = {}

    def _tables(self, device):
        hit = self._dev.get(str(device))
        if hit is None:
            hit = (torch.tensor(self.origins, dtype=torch.int32, device=device).contiguous(),
                   torch.tensor(self.windows, dtype=torch.int32, device=device).contiguous())
            self._dev[str(device)] = hit
        return hit

    def gather(self, x: torch.Tensor) -> torch.Tensor:
        """[N, C, H, W] -> [T * N, C, patch, patch] (tile-major)."""
        _lib.require_cuda(x)
        N, C, H, W = x.shape
        assert (H, W) == (self.H, self.W)
        xc = x.detach().contiguous().float()
        origins, _ = self._tables(x.device)
        vec = int(all(o[1] % 4 == 0 for o in self.origins))
        with _lib.on_device(x.device):
            out = torch.empty(self.T * N, C, self.patch, self.patch, dtype=torch.float32, device=x.device)
            rc = _lib.load().dcvic_tile_gather(_lib.ptr(xc), N, C, H, W, _lib.ptr(origins), self.T, self.patch,
                                               self.patch, vec, _lib.ptr(out), _lib.cur_stream())
            _lib.check(rc, "dcvic_tile_gather")
        return out

    def stitch(self, tiles: torch.Tensor, N: int) -> torch.Tensor:
        """[T * N, C', out_patch, out_patch] -> [N, C', out_H, out_W]."""
        _lib.require_cuda(tiles)
        TN, C, ph, pw = tiles.shape
        assert TN == self.T * N and (ph, pw) == (self.out_patch, self.out_patch), (tiles.shape, self.T, N, self.out_patch)
        tc = tiles.detach().contiguous().float()
        _, windows = self._tables(tiles.device)
        vec = int(all(w[1] % 4 == 0 and w[4] % 4 == 0 and w[5] % 4 == 0 for w in self.windows))
        with _lib.on_device(tiles.device):
            out = torch.empty(N, C, self.out_H, self.out_W, dtype=torch.float32, device=tiles.device)
            rc = _lib.load().dcvic_tile_stitch(_lib.ptr(tc), N, C, ph, pw, _lib.ptr(windows), self.T, vec,
                                               _lib.ptr(out), self.out_H, self.out_W, _lib.cur_stream())
            _lib.check(rc, "dcvic_tile_stitch")
        return out


def _run_batched(fn: Callable[[torch.Tensor], torch.Tensor], tiles: torch.Tensor, max_tiles: Optional[int]):
    if not max_tiles or tiles.shape[0] <= max_tiles:
        return fn(tiles)
    return torch.cat([fn(tiles[i:i + max_tiles]) for i in range(0, tiles.shape[0], max_tiles)], 0)


def encode_split(real_images: torch.Tensor, encode: Callable[[torch.Tensor], torch.Tensor], df: int = 8,
                 max_tiles_per_call: Optional[int] = None) -> torch.Tensor:
    """``_vq_encode_split`` (:190-246): ``encode`` maps [M, 3, 512, 512] -> [M, C_z, 512 / df, 512 / df] (the VQGAN
    encoder + quant_conv) and is called on all windows at once."""
    N, _, H, W = real_images.shape
    plan = TilePlan(H, W, SPLIT_WINDOW_SIZE, SPLIT_STRIDE, 1, df)
    return plan.stitch(_run_batched(encode, plan.gather(real_images), max_tiles_per_call), N)


def decode_split(y_hat: torch.Tensor, decode: Callable[[torch.Tensor], torch.Tensor], df: int = 16,
                 max_tiles_per_call: Optional[int] = None) -> torch.Tensor:
    """``decode_split`` (:413-473): ``decode`` maps latent windows [M, C, 32, 32] -> images [M, 3, 512, 512]; the
    result stays on the device (the reference assembles it on the CPU and copies it back)."""
    N, _, yH, yW = y_hat.shape
    plan = TilePlan(yH, yW, SPLIT_WINDOW_SIZE // df, SPLIT_STRIDE // df, df, 1)
    return plan.stitch(_run_batched(decode, plan.gather(y_hat), max_tiles_per_call), N)
