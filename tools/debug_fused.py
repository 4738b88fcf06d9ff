"""Where does the single-pass VQ kernel stand?  (needs `python -m dc_vic_b200.build --variant fzdbg -DDCVIC_FZ_DEBUG`)
    python tools/debug_fused.py [B H W D K] [kind]
Launches one forward, waits a few seconds, and - finished or hung - prints every warp's last progress mark
(code, value), read from a mapped host buffer.  Exits by os._exit so that a hung kernel cannot hold the process."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("DCVIC_B200_LIB", os.path.join(ROOT, "dc_vic_b200", "lib", "libdcvic_b200_fzdbg.so"))
import torch  # noqa: E402
import dc_vic_b200 as D  # noqa: E402
from dc_vic_b200 import _lib  # noqa: E402
from synth import vq_inputs  # noqa: E402

args = [int(a) for a in sys.argv[1:6]] if len(sys.argv) >= 6 else [1, 16, 16, 256, 1024]
kind = sys.argv[6] if len(sys.argv) > 6 else "D1b"
frozen = len(sys.argv) > 7 and sys.argv[7] == "frozen"      # second call on a frozen codebook (no prepare kernel)
B, H, W, Dm, K = args
z, E = vq_inputs(3, kind, B, Dm, H, W, K)
lib = _lib.load()
REC = 48
prog = torch.zeros(148 * 32 * REC, dtype=torch.int32).pin_memory()
if hasattr(lib, "dcvic_debug_set_fz_progress"):       # (timing runs use the normal library: DCVIC_B200_LIB=...)
    lib.dcvic_debug_set_fz_progress.restype = C.c_int
    lib.dcvic_debug_set_fz_progress.argtypes = [C.c_void_p]
    assert lib.dcvic_debug_set_fz_progress(C.c_void_p(prog.data_ptr())) == 0
m = D.VectorQuantizer2(K, Dm, 0.25, sane_index_shape=True).to("cuda:0")
m.embedding.weight.data.copy_(E)
zc = z.to("cuda:0")
torch.cuda.synchronize()
with torch.no_grad():
    if frozen:
        m.freeze_codebook()
        m(zc)
    out = m(zc)
ev = torch.cuda.Event()
ev.record()
t0 = time.time()
while not ev.query() and time.time() - t0 < 5.0:
    time.sleep(0.05)
done = ev.query()
NAMES = {0: "-", 1: "tmaB wait B_EMPTY", 2: "zload wait Z_EMPTY", 3: "fin wait converted", 4: "fin wait F_DONE",
         5: "mma wait T_EMPTY", 6: "mma wait A_FULL", 7: "mma wait B_FULL", 8: "conv wait A_EMPTY", 9: "conv wait Z_FULL",
         10: "epi wait ZZ", 11: "epi wait T_FULL", 12: "epi tile-end barrier", 13: "epi wait C_EMPTY", 14: "epi tile done",
         15: "cons wait C_FULL", 16: "cons wait F_FULL", 17: "cons work", 20: "epi pdl", 21: "epi after pdl",
         22: "cons setmaxnreg", 23: "cons after setmaxnreg", 30: "at exit"}
print("finished" if done else "HUNG", "after", round(time.time() - t0, 2), "s")
grid = min(148, 2 * ((B * H * W + 255) // 256))
ROLES = {0: ("tmaB", ["wait B_EMPTY"]), 1: ("zload", ["wait Z_EMPTY"]),
         2: ("fin", ["wait converted", "wait F_DONE", "wait read-out", "drain"]),
         3: ("mma", ["wait T_EMPTY", "wait A_FULL", "wait B_FULL"]),
         4: ("conv", ["wait A_EMPTY", "wait Z_FULL"]),
         8: ("epi", ["wait ZZ", "wait T_FULL", "tcgen05.ld+release", "wait C_EMPTY", "flags+emit", "tile end"]),
         24: ("cons", ["wait C_FULL", "rows issue", "wait F_FULL", "re-rank", "z_q+arrive"])}
for b in range(min(grid, 2)):
    print(f"CTA {b}:")
    for w in range(32):
        rec = [int(x) for x in prog[(b * 32 + w) * REC:(b * 32 + w) * REC + 8]]
        v = rec[0]
        role = ROLES[max(k for k in ROLES if k <= w)]
        cyc = "  ".join(f"{n} {rec[1 + i] * 8}" for i, n in enumerate(role[1]))
        print(f"  warp {w:2d} {role[0]:5s}: {NAMES.get(v >> 20, v >> 20):20s} {v & 0xFFFFF:6d} | total {rec[7] * 8}  {cyc}")
if done and os.environ.get("FZ_MARKS"):
    # tile timeline of CTA 0 (cycles since the first mark): per warp and tile the three marks of its role
    LEG = {"mma": "A chunk0 ready / A tile ready / tile issued", "conv": "A buffer free / first z chunk in / tile converted",
           "epi": "first T_FULL / last N-tile done / candidates published", "cons": "first unit starts / - / last unit done",
           "fin": "first load / first store / last store", "zload": "- / - / tile issued", "tmaB": "- / - / -"}
    for b in (0, 1):
        marks = prog[(b * 32) * REC:(b * 32 + 32) * REC].view(32, REC)[:, 8:].tolist()
        base = min(m[0] for m in marks if m[0])
        print(f"CTA {b}: marks in cycles since kernel start (after cluster sync); exit marks:",
              sorted(set((m[39] - base) & 0xFFFFFFFF for m in marks))[-1])
        for w in (1, 2, 3, 4, 8, 12, 16, 20, 24, 25, 26, 27, 28, 29, 30, 31):
            role = ROLES[max(k for k in ROLES if k <= w)][0]
            row = []
            for it in range(4):
                row.append("/".join(f"{((marks[w][1 + it * 4 + i] - base) & 0xFFFFFFFF):6d}" if marks[w][1 + it * 4 + i] else "     -" for i in range(3)))
            print(f"  warp {w:2d} {role:5s} " + "  |  ".join(row))
        for r, t in LEG.items():
            print(f"    {r}: {t}")
if not done:
    shown = 0
    for b in range(grid):
        for w in range(32):
            rec = [int(x) for x in prog[(b * 32 + w) * REC:(b * 32 + w) * REC + 8]]
            if (rec[0] >> 20) != 30 and shown < 80:
                role = ROLES[max(k for k in ROLES if k <= w)]
                print(f"  STUCK CTA {b:3d} warp {w:2d} {role[0]:5s}: {NAMES.get(rec[0] >> 20, rec[0] >> 20):20s} {rec[0] & 0xFFFFF}")
                shown += 1
if done:
    ws = list(m._ws._cache.values())[0]
    ctr = ws[:64].view(torch.int32).cpu().tolist()
    print("counters: overflow/full scans", ctr[1], "re-ranked tokens", ctr[3], "of", B * H * W,
          "| list overflow", ctr[6], "more than CK_MAX candidates", ctr[7], "no candidate", ctr[8], "unsafe", ctr[9])
    with torch.no_grad():
        for _ in range(3):
            m(zc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            m(zc)
        e1.record()
        torch.cuda.synchronize()
    print("forward: %.1f us" % (e0.elapsed_time(e1) * 100))
    from oracle import vq_oracle as O
    n_mis, n_out, n_tie = O.allowed_index_mismatch(z, E, out[2][2].cpu())
    print("mismatches", n_mis, "outside clause", n_out, "near ties", n_tie)
sys.stdout.flush()
os._exit(0 if done else 1)
