"""Generate the VQ golden vectors by running the REFERENCE's own vendored module.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

It imports ``/root/reference/taming/modules/vqvae/quantize.py`` unchanged
(``VectorQuantizer`` :9-107, ``VectorQuantizer2`` :213-329), runs it on seeded CPU FP32
inputs and stores inputs + outputs (+ autograd gradients) as small ``.pt`` files next to
this script.  The files are committed; nothing in tests/ reads /root/reference at run time.
Wide-codebook inputs are regenerated from their seed at test time and guarded by a SHA-256
of their bytes so the fixture stays small.
"""
import hashlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DCVIC_REFERENCE", "/root/reference")


sys.path.insert(0, os.path.dirname(HERE))
from synth import sha, vq_inputs as wide_inputs  # noqa: E402


def main():
    sys.path.insert(0, REF)
    from taming.modules.vqvae.quantize import VectorQuantizer, VectorQuantizer2  # noqa: E402

    torch.set_num_threads(1)  # fixes the SGEMM blocking -> reproducible last bits
    torch.manual_seed(0)
    meta = {"torch": torch.__version__, "threads": 1}

    def run_v2(z, E, beta=0.25, legacy=True, sane=True, with_grad=False, seed=0):
        m = VectorQuantizer2(E.shape[0], E.shape[1], beta, sane_index_shape=sane, legacy=legacy)
        m.embedding.weight.data.copy_(E)
        out = {}
        if with_grad:
            zz = z.clone().requires_grad_(True)
            z_q, loss, (_, _, idx) = m(zz)
            g = torch.Generator().manual_seed(1000 + seed)
            g_zq = torch.randn(z_q.shape, generator=g)
            g_loss = torch.tensor(0.7)
            (z_q * g_zq).sum().backward(retain_graph=True)
            dz_zq, dE_zq = zz.grad.clone(), (m.embedding.weight.grad.clone()
                                            if m.embedding.weight.grad is not None else torch.zeros_like(E))
            zz.grad = None
            m.embedding.weight.grad = None
            (loss * g_loss).backward()
            out.update(g_zq=g_zq, g_loss=g_loss, dz=dz_zq + zz.grad, dE=dE_zq + m.embedding.weight.grad,
                       dE_from_zq=dE_zq)
        else:
            with torch.no_grad():
                z_q, loss, (_, _, idx) = m(z)
        # oracle distances for the near-tie clause
        with torch.no_grad():
            rows = z.permute(0, 2, 3, 1).reshape(-1, E.shape[1])
            d = (rows ** 2).sum(1, keepdim=True) + (E ** 2).sum(1) - 2 * rows @ E.t()
            two = torch.topk(d, 2, dim=1, largest=False).values
            gap = (two[:, 1] - two[:, 0]) / two[:, 0].abs()
        out.update(z_q=z_q.detach(), loss=loss.detach(), idx=idx.to(torch.int32), rel_gap=gap,
                   beta=beta, legacy=legacy, sane=sane)
        return out

    # ---- narrow regime (the codebook DC-VIC actually uses: K=256, D=4) ----------------
    g = torch.Generator().manual_seed(2)
    E = torch.empty(256, 4).uniform_(-1 / 256, 1 / 256, generator=g)
    z = torch.randn(1, 4, 64, 96, generator=g)               # kodim03 768x512 / 8
    fx = run_v2(z, E, with_grad=True, seed=2)
    fx.update(z=z, E=E, meta=meta)
    torch.save(fx, os.path.join(HERE, "vq_v2_narrow_default.pt"))

    En = torch.randn(256, 4, generator=g)
    zn = torch.randn(2, 4, 24, 40, generator=g)
    for legacy in (True, False):
        fx = run_v2(zn, En, beta=0.25, legacy=legacy, sane=not legacy, with_grad=True, seed=3)
        fx.update(z=zn, E=En, meta=meta)
        torch.save(fx, os.path.join(HERE, f"vq_v2_narrow_randn_legacy{int(legacy)}.pt"))

    # D=8 (vq-f16 variants), ragged spatial size, K not a power of two
    E8 = torch.randn(1000, 8, generator=g)
    z8 = torch.randn(3, 8, 7, 13, generator=g)
    fx = run_v2(z8, E8, with_grad=True, seed=4)
    fx.update(z=z8, E=E8, meta=meta)
    torch.save(fx, os.path.join(HERE, "vq_v2_d8_ragged.pt"))

    # ---- V1 contract (one-hot + perplexity) --------------------------------------------
    g = torch.Generator().manual_seed(5)
    E1 = torch.randn(64, 8, generator=g) * 0.5
    z1 = torch.randn(2, 8, 6, 10, generator=g)
    m1 = VectorQuantizer(64, 8, 0.25)
    m1.embedding.weight.data.copy_(E1)
    zz = z1.clone().requires_grad_(True)
    z_q, loss, (ppl, onehot, idx) = m1(zz)
    g_zq = torch.randn(z_q.shape, generator=g)
    ((z_q * g_zq).sum() + 1.3 * loss).backward()
    torch.save(dict(z=z1, E=E1, beta=0.25, z_q=z_q.detach(), loss=loss.detach(), perplexity=ppl.detach(),
                    onehot_idx=onehot.argmax(1).to(torch.int32), onehot_rowsum=onehot.sum(1),
                    idx=idx.to(torch.int32), g_zq=g_zq, g_loss=torch.tensor(1.3), dz=zz.grad.clone(),
                    dE=m1.embedding.weight.grad.clone(), meta=meta),
               os.path.join(HERE, "vq_v1_small.pt"))
    # get_codebook_entry, both classes
    pick = torch.randint(0, 64, (2 * 6 * 10,), generator=g)
    m2 = VectorQuantizer2(64, 8, 0.25)
    m2.embedding.weight.data.copy_(E1)
    with torch.no_grad():
        e1 = m1.get_codebook_entry(pick, (2, 6, 10, 8))
        e2 = m2.get_codebook_entry(pick, (2, 6, 10, 8))
        e2_flat = m2.get_codebook_entry(pick, None)
    assert torch.equal(e1, e2)
    torch.save(dict(E=E1, pick=pick.to(torch.int32), shape=(2, 6, 10, 8), entry=e2, entry_flat=e2_flat, meta=meta),
               os.path.join(HERE, "vq_codebook_entry.pt"))

    # ---- wide regime (benchmark codebook 1024x256), inputs regenerated from seed ---------
    for kind, seed in (("D0", 0), ("D1", 1), ("D1b", 1)):
        z, E = wide_inputs(seed, kind, 1, 256, 16, 16, 1024)
        fx = run_v2(z, E)
        fx.update(kind=kind, seed=seed, shape=(1, 256, 16, 16, 1024), z_sha=sha(z), E_sha=sha(E), meta=meta)
        fx["z_q"] = fx["z_q"][:, :, :2, :].clone()   # keep two rows of tokens; full z_q == E[idx] check is in the test
        torch.save(fx, os.path.join(HERE, f"vq_v2_wide_{kind}.pt"))
        print(kind, "near-tie rows (<1e-6):", int((fx["rel_gap"] < 1e-6).sum()), "of", fx["rel_gap"].numel())

    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
