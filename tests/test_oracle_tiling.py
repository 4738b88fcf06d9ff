"""CPU: the tiling plan (dc_vic_b200.tiling.TilePlan, host arithmetic) and the oracle loops against the reference's
own `_vq_encode_split` / `decode_split` methods (called with a stand-in `self`) when /root/reference exists."""
import types

import pytest
import torch
import torch.nn.functional as F

from dc_vic_b200.tiling import TilePlan
from oracle import tiling_oracle as TO
from ref_shims import have_reference, import_reference

SIZES = [(1088, 1536), (1408, 2048), (1152, 1152), (1024 + 64, 512), (512, 1280), (1600, 1600)]


def enc(crop):          # a stand-in "VQGAN encoder": per-window, deterministic, 8x down, 4 channels
    return torch.cat([F.avg_pool2d(crop, 8), F.max_pool2d(crop[:, :1], 8)], 1)


def dec(crop):          # a stand-in decoder: 16x up, 3 channels
    return F.interpolate(crop[:, :3], scale_factor=16, mode="nearest") + 0.001 * crop[:, 3:4].mean()


@pytest.mark.parametrize("H,W", SIZES)
def test_plan_windows_partition_the_output(H, W):
    for patch, stride, num, den in ((512, 256, 1, 8), (32, 16, 16, 1)):
        h, w = (H, W) if num == 1 else (H // 16, W // 16)
        plan = TilePlan(h, w, patch, stride, num, den)
        cover = torch.zeros(plan.out_H, plan.out_W, dtype=torch.int32)
        for (y0, x0), (_y0, _x0, t, b, l, r) in zip(plan.origins, plan.windows):
            assert 0 <= y0 <= h - patch and 0 <= x0 <= w - patch
            assert _y0 <= t <= b <= _y0 + plan.out_patch and _x0 <= l <= r <= _x0 + plan.out_patch
            cover[t:b, l:r] += 1
        assert bool((cover == 1).all()), "keep-windows must cover every output pixel exactly once"


@pytest.mark.skipif(not have_reference(), reason="/root/reference is not present on this box")
@pytest.mark.parametrize("H,W", SIZES[:4])
def test_oracle_loops_match_the_reference_methods(H, W):
    import_reference()
    import src.models.comp_model.hyperprior_vic_model as M
    g = torch.Generator().manual_seed(H + W)
    img = torch.randn(1, 3, H, W, generator=g)
    fake = types.SimpleNamespace(vq_model=types.SimpleNamespace(encode=enc, embed_dim=4,
                                                                encoder=types.SimpleNamespace(num_resolutions=4)))
    ref_z = M.HyperpriorVicModel._vq_encode_split(fake, img)
    assert torch.equal(ref_z, TO.vq_encode_split(img, enc, 8, 4))
    y_hat = torch.randn(1, 6, H // 16, W // 16, generator=g)
    fake2 = types.SimpleNamespace(_decode=lambda c, w=None: dec(c))
    ref_img = M.HyperpriorVicModel.decode_split(fake2, y_hat, None)
    assert torch.equal(ref_img, TO.decode_split(y_hat, dec))
