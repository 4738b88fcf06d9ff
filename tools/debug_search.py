"""Stage-by-stage check of the tensor search against torch (debug aid).
    python tools/debug_search.py [kind] [B]
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

kind = sys.argv[1] if len(sys.argv) > 1 else "D1"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
H = W = 16 if B == 1 else 32
D, K = 256, 1024
dev = torch.device("cuda:0")
lib = _lib.load()
z, E = vq_inputs(1, kind, B, D, H, W, K)
zc, Ec = z.to(dev), E.to(dev)
N = B * H * W
zq = torch.empty_like(zc); idx = torch.empty(N, dtype=torch.int64, device=dev); loss = torch.empty((), device=dev)
nbytes = lib.dcvic_vq_workspace_bytes(B, D, H, W, K)
ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(flags, what):
    rc = lib.dcvic_vq_forward(_lib.ptr(zc), _lib.ptr(Ec), B, D, H, W, K, 0.25, 1, _lib.ptr(zq), _lib.ptr(idx),
                              _lib.ptr(loss), None, None, flags, _lib.ptr(ws), ws.numel(), st)
    try:
        torch.cuda.synchronize()
        print(what, "rc", rc, "sync ok")
    except Exception as e:  # noqa: BLE001
        print(what, "rc", rc, "SYNC FAILED", str(e).splitlines()[0])
        sys.exit(1)


call(_lib.VQ_STAGE_SEARCH_ONLY, "prep+search")
# workspace layout (vq_common.cuh)
def align(x): return (x + 255) // 256 * 256
o = 0
def take(n):
    global o
    at = o; o = align(o + n); return at
off_counters = take(64 * 4); off_ee = take(K * 4); off_nhee = take(K * 4); off_emax = take(16)
off_partials = take((N // 32 + 2) * 8); off_hist = take(K * 4); off_cand = take(N * 4); off_meta = take(N * 16)
off_list = take(N * 2 * 16 * 8); off_cb16 = take(K * (D + 64) * 2)
meta_raw = ws[off_meta:off_meta + N * 16].cpu()
m_gpu = meta_raw.view(torch.float32).view(N, 4)[:, :2].contiguous()
meta = meta_raw.view(torch.int16).view(N, 8)[:, 4:6].to(torch.int32)
cb16 = ws[off_cb16:off_cb16 + K * (D + 64) * 2].view(torch.float16).view(K, D + 64).float().cpu()
ee = ws[off_ee:off_ee + K * 4].view(torch.float32).cpu()
print("pad cols vs -ee/2 max abs err:", float((cb16[:, D:D + 3].sum(1) + 0.5 * ee).abs().max()), " pad rest max:", float(cb16[:, D + 3:].abs().max()))
rows = z.permute(0, 2, 3, 1).reshape(N, D)
s = rows.half().float() @ E.half().float().t() - 0.5 * ee[None, :]
s4 = s.view(N, K // 256, 2, 128)          # [token][nt][column half][128]
m_ref = s4.amax(dim=(1, 3))               # per column half
err = (m_gpu - m_ref).abs()
print("running max per half: max abs err", float(err.max()), " (score scale", float(s.abs().max()), ")")
bad = (err > 1e-2 * s.abs().max()).nonzero()
print("bad rows:", bad[:10].tolist(), "count", len(bad))
for r in (0, 1, 2, 127, 128, 129, 255):
    if r < N:
        print("row", r, "m_gpu", m_gpu[r].tolist(), "m_ref", m_ref[r].tolist(), "max|z.e| half0", float((s4[r, :, 0] + 0.5 * ee.view(K // 256, 2, 128)[:, 0]).amax()))
print("n entries min/max:", int(meta.min()), int(meta.max()))
call(_lib.VQ_STAGE_FINISH_ONLY, "finish")
ref = (rows.pow(2).sum(1, keepdim=True) + E.pow(2).sum(1)[None] - 2 * rows @ E.t()).argmin(1)
print("idx mismatches vs torch fp32:", int((idx.cpu() != ref).sum()), "of", N)
