"""Role timing of the tcgen05 search kernel (needs `python -m dc_vic_b200.build --trace` first).
    DCVIC_B200_LIB=dc_vic_b200/lib/libdcvic_b200_trace.so python tools/trace_run.py [D0|D1b]
Prints, per role, the cycles spent waiting / working summed over the launch (median and max over CTAs).
Only the MMA issuer's rows are filled unless the library was built with -DDCVIC_TRACE_ALL as well
(`python -m dc_vic_b200.build --trace -DDCVIC_TRACE_ALL`); any tracing slows this kernel by 30 %: its warps have no
registers to spare, so read the proportions, not the totals.
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
buf = (C.c_ulonglong * (296 * 16))()
lib.dcvic_debug_read_trace.restype = C.c_int
assert lib.dcvic_debug_read_trace(buf) == 0
names = ["A wait_empty", "A load+cvt", "E0 wait_zz", "E0 wait_full", "E0 process", "E0 reinit", "E1 wait_zz",
         "E1 wait_full", "E1 process", "E1 reinit", "MMA wait_t_empty", "MMA wait_a_full", "MMA wait_b_full", "MMA total"]
rows = [[buf[b * 16 + i] for i in range(16)] for b in range(148)]
for i, n in enumerate(names):
    col = [r[i] for r in rows if (i < 10 or r[13] > 0)]
    print(f"{n:18s} median {statistics.median(col):10.0f}  max {max(col):10.0f}  min {min(col):10.0f}")

# (the finish kernel has its own tool: tools/trace_finish.py)
