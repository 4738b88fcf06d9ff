"""Timing of the narrow-codebook quantizer (the codebooks DC-VIC itself uses) and the small entropy calls.
    python tools/time_narrow.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import dc_vic_b200 as D

dev = "cuda:0"
def timeit(fn, iters=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3   # us

for (B, Dm, H, W, K) in ((1, 4, 64, 96, 256), (1, 4, 176, 256, 256), (8, 4, 176, 256, 256), (1, 4, 176, 256, 16384), (1, 8, 88, 128, 16384)):
    m = D.VectorQuantizer2(K, Dm, 0.25, sane_index_shape=True).to(dev)
    m.freeze_codebook()
    z = torch.randn(B, Dm, H, W, device=dev)
    with torch.no_grad():
        us = timeit(lambda: m(z))
    N = B * H * W
    print(f"VQ narrow K={K:6d} D={Dm} N={N:7d}: {us:8.1f} us  {N/us:8.1f} Mtok/s  {N*(8*Dm+8)/us/1e3:7.1f} GB/s  path={m.search_path()}")

g = D.SteGaussianMeanScaleConditional(scale_bound=0.11).to(dev)
for (B, C, H, W) in ((1, 32, 32, 48), (6, 32, 16, 16), (1, 32, 128, 88)):
    y = torch.randn(B, C, H, W, device=dev); p = torch.randn(B, 2 * C, H, W, device=dev).abs() + 0.1
    with torch.no_grad():
        us = timeit(lambda: g(y, p, is_train=False))
    print(f"GC slice {B}x{C}x{H}x{W}: {us:7.1f} us per call")
