"""TEST INFRASTRUCTURE (oracle): the reference's serial tiling loops for images beyond 1024 px, restated with the
per-window network as a parameter.  `vq_encode_split` follows src/models/comp_model/hyperprior_vic_model.py:190-246,
`decode_split` :413-473 (SPLIT_WINDOW_SIZE = 512, SPLIT_STRIDE = 256, :26-27).  Pinned in tests/test_oracle_tiling.py
against the reference's own methods (called with a stand-in `self`) when /root/reference exists."""
import torch


def vq_encode_split(real_images, encode, df: int, ndim: int):
    N, _, H, W = real_images.size()
    stride, patch_size = 256, 512
    left_list = []
    for i in range(W // stride + 1):
        left = i * stride
        if left + patch_size < W:
            left_list.append(left)
        else:
            left_list.append(W - patch_size)
            break
    top_list = []
    for i in range(H // stride + 1):
        top = i * stride
        if top + patch_size < H:
            top_list.append(top)
        else:
            top_list.append(H - patch_size)
            break
    z_out = torch.zeros(N, ndim, H // df, W // df)
    for y0 in top_list:
        for x0 in left_list:
            z = encode(real_images[:, :, y0:y0 + patch_size, x0:x0 + patch_size])
            offset = (stride // 2) // df
            _x0, _y0 = x0 // df, y0 // df
            l = _x0 + offset if x0 > 0 else 0
            t = _y0 + offset if y0 > 0 else 0
            r = _x0 + offset + stride // df if x0 < left_list[-1] else W // df
            b = _y0 + offset + stride // df if y0 < top_list[-1] else H // df
            z_out[:, :, t:b, l:r] = z[:, :, t - _y0:b - _y0, l - _x0:r - _x0]
    return z_out


def decode_split(y_hat, decode, df: int = 16):
    N, _, yH, yW = y_hat.size()
    stride, patch_size = 256 // df, 512 // df
    left_list = []
    for l in range(0, yW, stride):
        if l + patch_size < yW:
            left_list.append(l)
        else:
            left_list.append(yW - patch_size)
            break
    top_list = []
    for t in range(0, yH, stride):
        if t + patch_size < yH:
            top_list.append(t)
        else:
            top_list.append(yH - patch_size)
            break
    fake_images = torch.zeros((N, 3, yH * df, yW * df)).fill_(-100.0)
    for y0 in top_list:
        for x0 in left_list:
            out_patch = decode(y_hat[:, :, y0:y0 + patch_size, x0:x0 + patch_size])
            offset = (stride // 2) * df
            _x0, _y0 = x0 * df, y0 * df
            l = _x0 + offset if x0 > 0 else 0
            t = _y0 + offset if y0 > 0 else 0
            r = _x0 + offset + stride * df if x0 < left_list[-1] else yW * df
            b = _y0 + offset + stride * df if y0 < top_list[-1] else yH * df
            fake_images[:, :, t:b, l:r] = out_patch[:, :, t - _y0:b - _y0, l - _x0:r - _x0]
    return fake_images
