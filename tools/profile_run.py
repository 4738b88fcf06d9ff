"""Small driver for ncu captures: runs the hot-path kernels a few times on the benchmark shapes.
    python tools/profile_run.py vq [D0|D1b] [iters]      python tools/profile_run.py gc [iters]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import dc_vic_b200 as D  # noqa: E402
from synth import vq_inputs, entropy_inputs  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "vq"
dev = "cuda:0"
if what == "vq":
    kind = sys.argv[2] if len(sys.argv) > 2 else "D0"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    z, E = vq_inputs(0, kind, 64, 256, 32, 32, 1024)
    m = D.VectorQuantizer2(1024, 256, 0.25, sane_index_shape=True).to(dev)
    m.embedding.weight.data.copy_(E)
    zc = z.to(dev)
    with torch.no_grad():
        for _ in range(iters):
            out = m(zc)
    torch.cuda.synchronize()
    ws = next(iter(m._ws._cache.values()))
    ctr = ws[:64 * 4].view(torch.int32).cpu()
    print("path", m.search_path(), "overflow rows", int(ctr[1]), "re-ranked rows", int(ctr[3]), "(cumulative over", iters, "iters)")
else:
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    y, p = entropy_inputs(2)
    g = D.SteGaussianMeanScaleConditional(scale_bound=0.11).to(dev)
    yc, pc = y.to(dev), p.to(dev)
    with torch.no_grad():
        for _ in range(iters):
            a, b = g(yc, pc, is_train=False)
            D.batch_bits(b)
    torch.cuda.synchronize()
    print("gc done")
