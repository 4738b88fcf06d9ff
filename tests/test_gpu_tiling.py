"""GPU: the batched tiling driver (gather -> one network call -> stitch) against the reference's serial loops (oracle)."""
import pytest
import torch
import torch.nn.functional as F

import dc_vic_b200 as D
from dc_vic_b200 import tiling
from oracle import tiling_oracle as TO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def enc(crop):
    return torch.cat([F.avg_pool2d(crop, 8), F.max_pool2d(crop[:, :1], 8)], 1)


def dec(crop):
    return F.interpolate(crop[:, :3], scale_factor=16, mode="nearest") + 0.001 * crop[:, 3:4].amax()


@pytest.mark.parametrize("N,H,W", [(1, 1088, 1536), (2, 1408, 2048), (1, 1152, 1152), (1, 520, 1284)])
def test_encode_and_decode_split_match_the_serial_loops(N, H, W):
    g = torch.Generator().manual_seed(H * 3 + W)
    img = torch.randn(N, 3, H, W, generator=g)
    z = tiling.encode_split(img.to(DEV), enc, df=8)
    assert torch.equal(z.cpu(), TO.vq_encode_split(img, enc, 8, 4))
    z2 = tiling.encode_split(img.to(DEV), enc, df=8, max_tiles_per_call=5)     # chunked network calls
    assert torch.equal(z2, z)
    if H % 16 == 0 and W % 16 == 0:
        y_hat = torch.randn(N, 6, H // 16, W // 16, generator=g)
        pe = lambda c: F.interpolate(c[:, :3], scale_factor=16, mode="nearest") + 0.001 * c[:, 3:4]   # noqa: E731
        pe16 = lambda c: F.interpolate(c[:, :3], scale_factor=16, mode="nearest") + \
            0.001 * F.interpolate(c[:, 3:4], scale_factor=16, mode="nearest")                      # noqa: E731
        out = tiling.decode_split(y_hat.to(DEV), pe16, df=16)
        assert out.is_cuda and torch.equal(out.cpu(), TO.decode_split(y_hat, pe16))
        del pe


def test_whole_latent_quantization_replaces_the_quantize_split_loop():
    """`_vq_quantize_split` (hyperprior_vic_model.py:170-188) crops 64 x 64 tokens to cap the N x K matrix; the fused
    quantizer has no such matrix: one call on the whole 2K latent gives the same indices as the crop loop."""
    z = torch.randn(1, 4, 176, 256, generator=torch.Generator().manual_seed(1))
    m = D.VectorQuantizer2(16384, 4, 0.25, sane_index_shape=True).to(DEV)
    with torch.no_grad():
        zq, _, (_, _, idx) = m(z.to(DEV))
        idx_crops = torch.full((1, 176, 256), -1, dtype=torch.long, device=DEV)
        zq_crops = torch.zeros_like(zq)
        for h in range(0, 176, 64):
            for w in range(0, 256, 64):
                a, _, (_, _, i) = m(z[:, :, h:h + 64, w:w + 64].contiguous().to(DEV))
                zq_crops[:, :, h:h + 64, w:w + 64] = a
                idx_crops[:, h:h + 64, w:w + 64] = i
    assert torch.equal(idx, idx_crops) and torch.equal(zq, zq_crops)
