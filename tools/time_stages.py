"""Stage timing of the wide VQ forward on C2 (64x256x32x32, K=1024) through the C ABI: whole step, prepare only,
prepare + search; the finish is the difference.  DCVIC_B200_LIB selects an experimental build of the library.
    python tools/time_stages.py [steps]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from dc_vic_b200 import _lib  # noqa: E402
from synth import vq_inputs  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
lib = _lib.load()
dev = "cuda:0"
B, Dm, H, W, K = 64, 256, 32, 32, 1024
N = B * H * W
ROT = 4
z0, E = vq_inputs(0, "D0", B, Dm, H, W, K)
Ec = E.to(dev)
zs = [z0.to(dev)] + [torch.randn(B, Dm, H, W, device=dev) for _ in range(ROT - 1)]
zqs = [torch.empty_like(zs[0]) for _ in range(ROT)]
idxs = [torch.empty(N, dtype=torch.int64, device=dev) for _ in range(ROT)]
loss = torch.empty((), device=dev)
ws = torch.zeros(lib.dcvic_vq_workspace_bytes(B, Dm, H, W, K), dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream()


def step(i, flags=0):
    j = i % ROT
    rc = lib.dcvic_vq_forward(_lib.ptr(zs[j]), _lib.ptr(Ec), B, Dm, H, W, K, 0.25, 1, _lib.ptr(zqs[j]),
                              _lib.ptr(idxs[j]), _lib.ptr(loss), None, None, flags, _lib.ptr(ws), ws.numel(),
                              C.c_void_p(stream.cuda_stream))
    _lib.check(rc, "dcvic_vq_forward")


def timed(flags):
    for i in range(10):
        step(i, flags)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        step(10 + i, flags)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps * 1e3


full = timed(0)
frozen = timed(_lib.VQ_REUSE_PREP)
prep = timed(_lib.VQ_STAGE_PREP_ONLY)
ps = timed(_lib.VQ_STAGE_SEARCH_ONLY)
print(f"lib={os.environ.get('DCVIC_B200_LIB', 'default')} full={full:.1f}us frozen={frozen:.1f}us prep={prep:.1f}us "
      f"search={ps - prep:.1f}us finish={full - ps:.1f}us")
