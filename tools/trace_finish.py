"""Phase timing of the TMA finish kernel (needs `python -m dc_vic_b200.build --trace` first; add -DDCVIC_FIN_NCW=n
to try another consumer-warp count).
    python tools/trace_finish.py [D0|D1b]
Consumer warps: cycles per phase summed over the warp's units.  Producer thread: cycles waiting / issuing.
"""
import ctypes as C
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("DCVIC_B200_LIB", os.path.join(ROOT, "dc_vic_b200", "lib", "libdcvic_b200_trace.so"))
import torch  # noqa: E402
import dc_vic_b200 as D  # noqa: E402
from dc_vic_b200 import _lib  # noqa: E402
from synth import vq_inputs  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "D0"
z, E = vq_inputs(0, kind, 64, 256, 32, 32, 1024)
m = D.VectorQuantizer2(1024, 256, 0.25, sane_index_shape=True).to("cuda:0")
m.embedding.weight.data.copy_(E)
zc = z.to("cuda:0")
with torch.no_grad():
    for _ in range(3):
        m(zc)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * (148 * 32 * 8))()
lib.dcvic_debug_read_ftma_trace.restype = C.c_int
assert lib.dcvic_debug_read_ftma_trace(buf) == 0
rows = [[[buf[(b * 32 + w) * 8 + i] for i in range(8)] for w in range(32)] for b in range(148)]
NEW = int(os.environ.get("DCVIC_FIN_NEW", "2"))             # expander warps of the traced build
last = max(w for w in range(32) if rows[0][w][7] > 0)       # the producer is the last warp that wrote
ncw = last - NEW
cons = [rows[b][w] for b in range(148) for w in range(ncw)]
expd = [rows[b][w] for b in range(148) for w in range(ncw, last)]
prod = [rows[b][last] for b in range(148)]
cn = ["pdl wait", "draw unit + wait cands", "issue rows", "wait tile", "re-rank", "z_q + arrive", "-", "total"]
en = ["pdl wait", "loads + wait stage", "expand + arrive", "-", "-", "-", "-", "total"]
pn = ["prologue loads", "wait done", "wait store read", "issue", "drain", "-", "-", "total"]
print("consumer warps:", ncw)
for i, n in enumerate(cn):
    if n != "-":
        col = [r[i] for r in cons]
        print(f"  {n:26s} median {statistics.median(col):9.0f}  max {max(col):9.0f}  min {min(col):9.0f}")
print("expander warps:", NEW)
for i, n in enumerate(en):
    if n != "-":
        col = [r[i] for r in expd]
        print(f"  {n:26s} median {statistics.median(col):9.0f}  max {max(col):9.0f}  min {min(col):9.0f}")
print("producer thread")
for i, n in enumerate(pn):
    if n != "-":
        col = [r[i] for r in prod]
        print(f"  {n:26s} median {statistics.median(col):9.0f}  max {max(col):9.0f}  min {min(col):9.0f}")

# balance: work (total - pdl wait) per warp
work = [[rows[b][w][7] - rows[b][w][0] for w in range(ncw)] for b in range(148)]
cta_max = [max(x) for x in work]
cta_mean = [sum(x) / len(x) for x in work]
print(f"work per warp (total - pdl wait): mean over all {statistics.mean([v for x in work for v in x]):.0f}")
print(f"  per CTA: slowest warp median {statistics.median(cta_max):.0f} max {max(cta_max):.0f} min {min(cta_max):.0f};"
      f" mean-warp median {statistics.median(cta_mean):.0f} max {max(cta_mean):.0f} min {min(cta_mean):.0f}")
pw = [rows[b][last][7] for b in range(148)]
print(f"  producer total: median {statistics.median(pw):.0f} max {max(pw):.0f} min {min(pw):.0f}")
