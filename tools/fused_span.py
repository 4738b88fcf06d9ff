"""Wall-clock anatomy of one single-pass VQ launch in a back-to-back chain (needs the marks build:
`python -m dc_vic_b200.build --variant gmarks -DDCVIC_FZ_DEBUG -DDCVIC_FZ_MARKS_ONLY -DDCVIC_FZ_NO_CMARKS`).
    python tools/fused_span.py [steps]
Every warp writes %globaltimer at kernel entry, after the set-up, after the predecessor completed (PDL wait), when it
leaves its role loop and at kernel exit; the last launch's marks stay in the mapped buffer.  Printed: the spread over
all CTAs of each mark (ns since the first CTA's entry) and the launch period measured with CUDA events."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("DCVIC_B200_LIB", os.path.join(ROOT, "dc_vic_b200", "lib", "libdcvic_b200_gmarks.so"))
import torch  # noqa: E402
from dc_vic_b200 import _lib  # noqa: E402
from synth import vq_inputs  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lib = _lib.load()
REC = 48
prog = torch.zeros(148 * 32 * REC, dtype=torch.int32, device="cuda:0")   # device memory: a mark is an L2 write
lib.dcvic_debug_set_fz_progress.restype = C.c_int
lib.dcvic_debug_set_fz_progress.argtypes = [C.c_void_p]
assert lib.dcvic_debug_set_fz_progress(C.c_void_p(prog.data_ptr())) == 0
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
REUSE = 1 << 0 if not hasattr(_lib, "VQ_REUSE_PREP") else _lib.VQ_REUSE_PREP


def step(i, flags):
    j = i % ROT
    rc = lib.dcvic_vq_forward(_lib.ptr(zs[j]), _lib.ptr(Ec), B, Dm, H, W, K, 0.25, 1, _lib.ptr(zqs[j]),
                              _lib.ptr(idxs[j]), _lib.ptr(loss), None, None, flags, _lib.ptr(ws), ws.numel(),
                              C.c_void_p(stream.cuda_stream))
    _lib.check(rc, "dcvic_vq_forward")


step(0, 0)
for i in range(5):
    step(i, REUSE)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(steps):
    step(i, REUSE)          # the last one runs on z0 (D0) when steps % ROT == 1
b.record()
torch.cuda.synchronize()
print("period %.2f us per launch (marks build)" % (a.elapsed_time(b) / steps * 1e3))
m = prog.cpu().view(148, 32, REC)[:, :, 8:].to(torch.int64) & 0xFFFFFFFF
m = m[m[:, 0, 33] != 0]                      # CTAs that ran
print("CTAs", m.shape[0])
t0 = int(m[:, :, 33].min())
names = {33: "kernel entry", 34: "set-up done (cluster sync)", 35: "predecessor complete (PDL wait)",
         36: "role loop left", 37: "kernel exit"}
for k, n in names.items():
    v = (m[:, :, k] - t0) & 0xFFFFFFFF
    per_cta = v.max(dim=1).values.float()
    print(f"  {n:34s} first CTA {int(per_cta.min()):7d} ns  median {int(per_cta.median()):7d}  last CTA {int(per_cta.max()):7d}")
roles = {"tmaB": [0], "zload": [1], "fin": [2], "mma": [3], "conv": [4, 5, 6, 7], "epi": list(range(8, 24)),
         "cons": list(range(24, 32))}
for r, ws_ in roles.items():
    v = ((m[:, ws_, 36] - t0) & 0xFFFFFFFF).max(dim=1).values.float()
    print(f"  role loop left, {r:5s}: min {int(v.min()):7d}  median {int(v.median()):7d}  max {int(v.max()):7d}")
span = ((m[:, :, 37] - m[:, :, 33]) & 0xFFFFFFFF).max(dim=1).values.float()
print("  per-CTA entry->exit: min %d median %d max %d ns" % (int(span.min()), int(span.median()), int(span.max())))
if os.environ.get("FZ_SPAN_TABLE"):
    # per CTA pair: groups of 32 tokens per CTA, wall-clock of the main loop (ns), per role
    num_gp = (N + 63) // 64
    npairs = m.shape[0] // 2
    rows = []
    for p in range(npairs):
        groups = (p + 1) * num_gp // npairs - p * num_gp // npairs
        c = m[2 * p]
        st = int(((c[:, 35] - t0) & 0xFFFFFFFF).max())
        rows.append((p, groups, st, *[int(((c[ws_, 36] - t0) & 0xFFFFFFFF).max()) for ws_ in roles.values()]))
    print("pair groups start " + " ".join(roles))
    for r in rows:
        print(" ".join(f"{x:6d}" for x in r))
