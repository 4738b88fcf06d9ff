"""EntropyBottleneck at 64x192x64x64 (50 M latents): us per module call, evaluation (look-up kernel) and training
(direct kernel) forward.  (Tried: two tanh for one reciprocal - 36 instead of 49 MUFU operations per latent - 820 us
against 790: the direct kernel is bound by its instruction count, not by the MUFU pipe.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dc_vic_b200 as D
dev = "cuda:0"
eb = D.SteEntropyBottleneck(channels=192).to(dev)
x = 3 * torch.randn(64, 192, 64, 64, device=dev)
with torch.no_grad():
    for _ in range(3):
        eb(x, is_train=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        eb(x, is_train=False)
    b.record()
    torch.cuda.synchronize()
us = a.elapsed_time(b) * 100
print(f"lib={os.environ.get('DCVIC_B200_LIB', 'default')} eb eval forward {us:.1f} us = {12 * x.numel() / us / 1e6:.2f} TB/s")
nz = torch.rand_like(x) - 0.5
with torch.no_grad():
    for _ in range(2):
        eb(x, is_train=True, noise=nz)
    torch.cuda.synchronize()
    a.record()
    for _ in range(5):
        eb(x, is_train=True, noise=nz)
    b.record()
    torch.cuda.synchronize()
print(f"eb train forward {a.elapsed_time(b) * 200:.1f} us")
