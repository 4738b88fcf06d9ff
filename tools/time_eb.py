"""EntropyBottleneck evaluation forward at 64x192x64x64 (50 M latents): us per call through the C ABI path of the module."""
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
