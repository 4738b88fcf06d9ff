"""Device time and achieved HBM bandwidth (algorithmic bytes of SURVEY 8d) of the secondary kernels on the C2 / C3
shapes: VQ backward, codebook gather, one-hot feature, token decode, entropy bottleneck, rate sum, codec step.
    python tools/time_secondary.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import dc_vic_b200 as D  # noqa: E402

dev = "cuda:0"
ROT = 4


def timeit(make, iters=40):
    """make(i) runs one call on buffer set i % ROT (inputs larger than L2 in total)."""
    for i in range(5):
        make(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        make(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3   # seconds


def report(name, sec, nbytes):
    print(f"{name:44s} {sec * 1e6:8.1f} us  {nbytes / sec / 1e9:8.0f} GB/s (algorithmic {nbytes / 1e6:.1f} MB)")


B, Dm, H, W, K = 64, 256, 32, 32, 1024
N = B * H * W
m = D.VectorQuantizer2(K, Dm, 0.25, sane_index_shape=True).to(dev)
zs = [torch.randn(B, Dm, H, W, device=dev, requires_grad=True) for _ in range(ROT)]
outs = [m(z) for z in zs]
gz = [torch.randn(B, Dm, H, W, device=dev) for _ in range(ROT)]


def bwd(i):
    j = i % ROT
    zq, loss, _ = outs[j]
    torch.autograd.grad([zq, loss], [zs[j], m.embedding.weight], [gz[j], torch.ones_like(loss)], retain_graph=True)


report("VQ backward (dz + dE), C2", timeit(bwd), N * (3 * 4 * Dm + 8) + 4 * K * Dm)


def bwd_z(i):
    j = i % ROT
    zq, loss, _ = outs[j]
    torch.autograd.grad([zq, loss], [zs[j]], [gz[j], torch.ones_like(loss)], retain_graph=True)


report("VQ backward (dz only: frozen codebook), C2", timeit(bwd_z), N * (3 * 4 * Dm + 8))

# the same through the C ABI (no autograd engine in the way)
import ctypes as C  # noqa: E402
from dc_vic_b200 import _lib  # noqa: E402
lib = _lib.load()
dzs = [torch.empty(B, Dm, H, W, device=dev) for _ in range(ROT)]
dE = torch.empty(K, Dm, device=dev)
one = torch.ones((), device=dev)
zd = [z.detach() for z in zs]
idx64 = [o[2][2].reshape(-1) for o in outs]
Ew = m.embedding.weight.detach()


def bwd_abi(i, with_dE):
    j = i % ROT
    rc = lib.dcvic_vq_backward(_lib.ptr(gz[j]), _lib.ptr(one), _lib.ptr(zd[j]), _lib.ptr(Ew), _lib.ptr(idx64[j]),
                               B, Dm, H, W, K, 0.25, 1, _lib.ptr(dzs[j]), _lib.ptr(dE) if with_dE else None,
                               _lib.cur_stream())
    assert rc == 0


report("  C ABI: dz + dE", timeit(lambda i: bwd_abi(i, True)), N * (3 * 4 * Dm + 8) + 4 * K * Dm)
report("  C ABI: dz only", timeit(lambda i: bwd_abi(i, False)), N * (3 * 4 * Dm + 8))

idx = [o[2][2] for o in outs]
with torch.no_grad():
    report("codebook gather -> NCHW, C2", timeit(lambda i: D.codebook_lookup(idx[i % ROT], m.embedding.weight)),
           N * (8 + 4 * Dm))
    idx_small = [torch.randint(0, 256, (8, 64, 96), device=dev) for _ in range(ROT)]
    report("one-hot feature, 8 x 64x96 tokens, K=256", timeit(lambda i: D.onehot_feature(idx_small[i % ROT], 256)),
           8 * 64 * 96 * (8 + 4 * 256))
    logits = [torch.randn(8, 256, 64, 96, device=dev) for _ in range(ROT)]
    cb = torch.randn(256, 4, device=dev)
    report("token decode (argmax + gather), 8x256x64x96", timeit(lambda i: D.decode_tokens(logits[i % ROT], cb)),
           8 * 64 * 96 * (4 * 256 + 8 + 16))

    eb = D.SteEntropyBottleneck(channels=192).to(dev)
    zh = [3 * torch.randn(64, 192, 8, 8, device=dev) for _ in range(ROT)]
    report("entropy bottleneck eval forward, 64x192x8x8", timeit(lambda i: eb(zh[i % ROT], is_train=False)),
           zh[0].numel() * 12)
    zbig = [3 * torch.randn(64, 192, 64, 64, device=dev) for _ in range(ROT)]
    report("entropy bottleneck eval forward, 64x192x64x64", timeit(lambda i: eb(zbig[i % ROT], is_train=False)),
           zbig[0].numel() * 12)

    lik = [torch.rand(64, 320, 32, 32, device=dev) * 0.9 + 0.05 for _ in range(ROT)]
    report("rate sum (batch_bits), C3", timeit(lambda i: D.batch_bits(lik[i % ROT])), lik[0].numel() * 4)
    y = [torch.randn(64, 320, 32, 32, device=dev) for _ in range(2)]
    p = [torch.randn(64, 640, 32, 32, device=dev).abs() + 0.05 for _ in range(2)]
    tab = D.get_scale_table().to(dev)
    report("codec step (y_hat, lik, symbols, indexes), C3",
           timeit(lambda i: D.gaussian_codec_step(y[i % 2], p[i % 2], tab), iters=20), y[0].numel() * 28)
