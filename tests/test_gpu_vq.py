"""GPU parity of the VQ quantizer path (through the modules -> ctypes -> C ABI -> CUDA) against
the CPU oracle and the reference-generated goldens.  Rule (BASELINE.json): indices equal to the
reference FP32 path except documented near-ties (oracle distance of our index within 1e-6
relative of the oracle minimum); z_q bit-exact given the index; loss within 1e-5 relative."""
import os

import pytest
import torch

import dc_vic_b200 as D
from oracle import vq_oracle as O
from synth import vq_inputs

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def load(name):
    return torch.load(os.path.join(G, name), weights_only=False)


def make(cls, E, **kw):
    m = cls(E.shape[0], E.shape[1], kw.pop("beta", 0.25), **kw).to(DEV)
    m.embedding.weight.data.copy_(E)
    return m


def check_against_oracle(z, E, out, idx_flat, beta=0.25, legacy=True, allow_near_ties=True):
    """Index rule + z_q bit-exactness (given OUR index) + loss tolerance."""
    z_q, loss = out[0].cpu(), out[1].cpu()
    idx = idx_flat.reshape(-1).cpu()
    n_mis, n_out, n_tie = O.allowed_index_mismatch(z, E, idx)
    assert n_out == 0, f"{n_out} index mismatches outside the near-tie clause ({n_mis} total)"
    if not allow_near_ties:
        assert n_mis == 0
    rows = O.token_rows(z)
    picked = E[idx]
    ste = (rows + (picked - rows)).view(z.shape[0], z.shape[2], z.shape[3], -1).permute(0, 3, 1, 2)
    assert torch.equal(z_q, ste), "z_q is not bit-exact z + (E[idx] - z)"
    m = ((picked - rows) ** 2).mean()
    ref_loss = m + beta * m if legacy else beta * m + m
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    return n_mis, n_tie


@pytest.mark.parametrize("name", ["vq_v2_narrow_default.pt", "vq_v2_narrow_randn_legacy0.pt",
                                  "vq_v2_narrow_randn_legacy1.pt", "vq_v2_d8_ragged.pt"])
@pytest.mark.parametrize("search", ["auto", "exact"])
def test_v2_golden_forward_backward(name, search):
    fx = load(name)
    z, E = fx["z"], fx["E"]
    m = make(D.VectorQuantizer2, E, beta=fx["beta"], legacy=fx["legacy"], sane_index_shape=fx["sane"])
    m.search = search
    zc = z.to(DEV).requires_grad_(True)
    z_q, loss, (ppl, enc, idx) = m(zc)
    assert ppl is None and enc is None and idx.dtype == torch.int64
    assert tuple(idx.shape) == ((z.shape[0], z.shape[2], z.shape[3]) if fx["sane"] else (z.shape[0] * z.shape[2] * z.shape[3],))
    n_mis, _ = check_against_oracle(z, E, (z_q, loss), idx, fx["beta"], fx["legacy"])
    # vs the reference's own outputs: every row whose reference top-2 gap is not a near-tie must agree
    same = idx.reshape(-1).cpu().to(torch.int32) == fx["idx"].reshape(-1)
    assert bool((same | (fx["rel_gap"] < 1e-6)).all())
    if n_mis == 0:
        assert torch.equal(z_q.detach().cpu(), fx["z_q"])
        assert abs(float(loss) - float(fx["loss"])) <= 1e-5 * abs(float(fx["loss"]))
        ((z_q * fx["g_zq"].to(DEV)).sum() + fx["g_loss"].to(DEV) * loss).backward()
        assert torch.allclose(zc.grad.cpu(), fx["dz"], rtol=1e-5, atol=1e-7)
        assert torch.allclose(m.embedding.weight.grad.cpu(), fx["dE"], rtol=1e-4, atol=1e-7)


def test_v1_golden_contract():
    fx = load("vq_v1_small.pt")
    z, E = fx["z"], fx["E"]
    m = make(D.VectorQuantizer, E, beta=fx["beta"])
    zc = z.to(DEV).requires_grad_(True)
    z_q, loss, (ppl, onehot, idx) = m(zc)
    assert tuple(idx.shape) == (z.shape[0] * z.shape[2] * z.shape[3], 1)
    assert tuple(onehot.shape) == (idx.shape[0], E.shape[0]) and onehot.dtype == torch.float32
    assert torch.equal(idx.squeeze(1).cpu().to(torch.int32), fx["idx"].squeeze(1))
    assert torch.equal(onehot.argmax(1).cpu().to(torch.int32), fx["onehot_idx"])
    assert torch.equal(onehot.sum(1).cpu(), fx["onehot_rowsum"])
    assert torch.equal(z_q.detach().cpu(), fx["z_q"])
    assert abs(float(loss) - float(fx["loss"])) <= 1e-5 * abs(float(fx["loss"]))
    assert abs(float(ppl) - float(fx["perplexity"])) <= 1e-5 * float(fx["perplexity"])
    ((z_q * fx["g_zq"].to(DEV)).sum() + fx["g_loss"].to(DEV) * loss).backward()
    assert torch.allclose(zc.grad.cpu(), fx["dz"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(m.embedding.weight.grad.cpu(), fx["dE"], rtol=1e-4, atol=1e-7)


def test_codebook_entry_and_onehot_feature():
    fx = load("vq_codebook_entry.pt")
    E, pick, shape = fx["E"], fx["pick"].long(), fx["shape"]
    for cls in (D.VectorQuantizer, D.VectorQuantizer2):
        m = make(cls, E)
        assert torch.equal(m.get_codebook_entry(pick.to(DEV), shape).cpu(), fx["entry"])
        assert torch.equal(m.get_codebook_entry(pick.to(DEV), None).cpu(), fx["entry_flat"])
    b, h, w, _ = shape
    idx = pick.view(b, h, w).to(DEV)
    assert torch.equal(D.codebook_lookup(idx, E.to(DEV)).cpu(), O.indices_to_latent(pick.view(b, h, w), E))
    assert torch.equal(D.onehot_feature(idx, E.shape[0]).cpu(), O.onehot_feature(pick.view(b, h, w), E.shape[0]))
    # ragged HW (not a multiple of 4) and a 256-entry alphabet
    idx2 = torch.randint(0, 256, (2, 7, 9), generator=torch.Generator().manual_seed(3))
    assert torch.equal(D.onehot_feature(idx2.to(DEV), 256).cpu(), O.onehot_feature(idx2, 256))


@pytest.mark.parametrize("kind", ["D0", "D1", "D1b"])
@pytest.mark.parametrize("search", ["auto", "exact"])
def test_wide_codebook_golden(kind, search):
    fx = load(f"vq_v2_wide_{kind}.pt")
    B, Dm, H, W, K = fx["shape"]
    z, E = vq_inputs(fx["seed"], kind, B, Dm, H, W, K)
    m = make(D.VectorQuantizer2, E, sane_index_shape=True)
    m.search = search
    with torch.no_grad():
        z_q, loss, (_, _, idx) = m(z.to(DEV))
    n_mis, n_tie = check_against_oracle(z, E, (z_q, loss), idx, allow_near_ties=(kind == "D0"))
    same = idx.reshape(-1).cpu().to(torch.int32) == fx["idx"].reshape(-1)
    assert bool((same | (fx["rel_gap"] < 1e-6)).all())
    if n_mis == 0:
        assert torch.equal(z_q[:, :, :2, :].cpu(), fx["z_q"])


@pytest.mark.parametrize("kind,shape", [("D1b", (2, 256, 24, 20, 1024)),      # HW not a multiple of the token tile
                                        ("D0", (3, 256, 16, 16, 1024)),
                                        ("D1b", (1, 128, 16, 16, 512)),
                                        ("D1", (1, 64, 9, 7, 256)),            # ragged, tiny
                                        ("D1b", (2, 96, 8, 8, 300))])          # e_dim / K off the tensor path's grid
def test_wide_shapes_vs_oracle(kind, shape):
    B, Dm, H, W, K = shape
    z, E = vq_inputs(7, kind, B, Dm, H, W, K)
    m = make(D.VectorQuantizer2, E, sane_index_shape=True)
    with torch.no_grad():
        out = m(z.to(DEV))
    check_against_oracle(z, E, out, out[2][2], allow_near_ties=(kind == "D0"))
    m.search = "exact"
    with torch.no_grad():
        out2 = m(z.to(DEV))
    check_against_oracle(z, E, out2, out2[2][2], allow_near_ties=(kind == "D0"))


def test_exact_ties_resolve_to_lowest_index():
    # duplicated codewords: every token has at least one exact tie
    g = torch.Generator().manual_seed(11)
    base = torch.randn(64, 256, generator=g)
    E = torch.cat([base, base, base, base], 0)            # K = 256, code k == code k+64 == ...
    z = torch.randn(1, 256, 8, 16, generator=g)
    for search in ("auto", "exact"):
        m = make(D.VectorQuantizer2, E)
        m.search = search
        with torch.no_grad():
            _, _, (_, _, idx) = m(z.to(DEV))
        assert int(idx.max()) < 64, "tie not resolved to the lowest index"
        ref = O.vq2_forward(z, E).indices
        n_mis, n_out, _ = O.allowed_index_mismatch(z, E, idx.cpu())
        assert n_out == 0
    En = torch.cat([torch.randn(40, 4, generator=g)] * 3, 0)
    zn = torch.randn(2, 4, 10, 10, generator=g)
    with torch.no_grad():
        _, _, (_, _, idn) = make(D.VectorQuantizer2, En)(zn.to(DEV))
    assert int(idn.max()) < 40


def test_full_size_properties_c2():
    """BASELINE config 2 (N=65,536, K=1024, D=256): size-independent properties."""
    z, E = vq_inputs(0, "D1b", 64, 256, 32, 32, 1024)
    m = make(D.VectorQuantizer2, E, sane_index_shape=True).freeze_codebook()
    zc = z.to(DEV)
    with torch.no_grad():
        z_q, loss, (_, _, idx) = m(zc)
        z_q2, loss2, (_, _, idx2) = m(zc)                      # idempotent / deterministic, prep reused
        assert torch.equal(idx, idx2) and torch.equal(z_q, z_q2) and torch.equal(loss, loss2)
        Ec = E.to(DEV)
        rows = zc.permute(0, 2, 3, 1).reshape(-1, 256)
        picked = Ec[idx.reshape(-1)]
        assert torch.equal(z_q.permute(0, 2, 3, 1).reshape(-1, 256), rows + (picked - rows))
        # re-quantizing the codewords themselves is the identity map (up to duplicate-free codebook)
        zz = Ec.t().reshape(1, 256, 32, 32).contiguous()
        _, l0, (_, _, self_idx) = m(zz)
        assert torch.equal(self_idx.reshape(-1), torch.arange(1024, device=DEV))
        assert float(l0) == 0.0
        # chosen code is at least as close as 64 random other codes (FP64 check)
        probe = torch.randint(0, 1024, (rows.shape[0], 64), device=DEV)
        d_best = ((rows.double() - picked.double()) ** 2).sum(1)
        for j in range(0, 64, 16):
            d_other = ((rows.double()[:, None, :] - Ec.double()[probe[:, j:j + 16]]) ** 2).sum(2)
            assert bool((d_best[:, None] <= d_other * (1 + 1e-6) + 1e-9).all())
    # oracle on a bounded sample of the same batch (first 2 images)
    n_mis, n_out, _ = O.allowed_index_mismatch(z[:2], E, idx[:2].cpu())
    assert n_out == 0


def test_backward_matches_oracle_autograd_wide():
    z, E = vq_inputs(5, "D1", 1, 256, 8, 8, 1024)
    m = make(D.VectorQuantizer2, E, legacy=False)
    zc = z.to(DEV).requires_grad_(True)
    z_q, loss, (_, _, idx) = m(zc)
    g = torch.randn(z.shape, generator=torch.Generator().manual_seed(1))
    ((z_q * g.to(DEV)).sum() + 2.5 * loss).backward()
    zo = z.clone().requires_grad_(True)
    Eo = E.clone().requires_grad_(True)
    oo = O.vq2_forward(zo, Eo, legacy=False)
    assert torch.equal(oo.indices, idx.cpu())
    ((oo.z_q * g).sum() + 2.5 * oo.loss).backward()
    assert torch.allclose(zc.grad.cpu(), zo.grad, rtol=1e-5, atol=1e-7)
    assert torch.allclose(m.embedding.weight.grad.cpu(), Eo.grad, rtol=1e-4, atol=1e-8)


def test_from_reference_swap_keeps_codebook():
    class FakeRef(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.n_e, self.e_dim, self.beta, self.legacy, self.sane_index_shape, self.remap = 32, 4, 0.25, True, True, None
            self.embedding = torch.nn.Embedding(32, 4)
    holder = torch.nn.Module()
    holder.quantize = FakeRef().to(DEV)
    w = holder.quantize.embedding.weight.data.clone()
    D.swap_quantizer(holder)
    assert isinstance(holder.quantize, D.VectorQuantizer2) and holder.quantize.sane_index_shape
    assert torch.equal(holder.quantize.embedding.weight.data, w)
    assert list(holder.state_dict()) == ["quantize.embedding.weight"]


def test_tensor_core_path_is_selected_and_agrees_with_exact_scan():
    """On sm_100 the benchmark codebook must run the tcgen05 search (no silent SIMT fallback) and
    produce the same indices as the FP32 scan wherever the scan itself is not at a near-tie."""
    assert torch.cuda.get_device_capability(0)[0] == 10
    for kind in ("D0", "D1b"):
        z, E = vq_inputs(3, kind, 4, 256, 32, 32, 1024)
        m = make(D.VectorQuantizer2, E)
        m.search = "tensor"                      # errors out instead of falling back
        assert m.search_path() == "tcgen05"
        with torch.no_grad():
            zq_t, l_t, (_, _, i_t) = m(z.to(DEV))
            m.search = "exact"
            assert m.search_path() == "exact-simt"
            zq_e, l_e, (_, _, i_e) = m(z.to(DEV))
        n_mis, n_out, n_tie = O.allowed_index_mismatch(z, E, i_t.cpu())
        assert n_out == 0
        diff = int((i_t != i_e).sum())
        assert diff <= n_tie, (diff, n_tie)
        if kind == "D1b":
            assert diff == 0 and torch.equal(zq_t, zq_e)
    narrow = make(D.VectorQuantizer2, torch.randn(256, 4))
    assert narrow.search_path() == "narrow-simt"
    narrow.search = "tensor"
    with pytest.raises(RuntimeError, match="not supported"):
        narrow(torch.randn(1, 4, 8, 8, device=DEV))


def test_tensor_path_outside_fp16_range_stays_exact():
    """Tokens / codebooks whose values do not fit FP16 must come out exactly as the FP32 reference
    (they are flagged by the search and scanned in FP32 by the finish kernel)."""
    z, E = vq_inputs(3, "D1b", 1, 256, 8, 16, 1024)
    z = z.clone()
    z[0, :, 0, :4] *= 3.0e5            # 4 tokens with |z| far beyond 65504
    z[0, 5, 1, 0] = 7.0e4              # one element just over the FP16 maximum
    m = make(D.VectorQuantizer2, E, sane_index_shape=True)
    assert m.search_path() == "tcgen05"
    with torch.no_grad():
        out = m(z.to(DEV))
    check_against_oracle(z, E, out, out[2][2], allow_near_ties=True)
    E2 = E.clone()
    E2[7] *= 1.0e4                     # |e|^2 / 2 no longer representable: the whole codebook is flagged unsafe
    m2 = make(D.VectorQuantizer2, E2, sane_index_shape=True)
    z2 = vq_inputs(4, "D1b", 1, 256, 8, 16, 1024)[0]
    with torch.no_grad():
        out2 = m2(z2.to(DEV))
    check_against_oracle(z2, E2, out2, out2[2][2], allow_near_ties=True)


def test_finish_ring_wraps_at_other_widths():
    """e_dim 128 / 192 (8-stage ring of the TMA finish kernel) with enough tokens that every persistent CTA goes
    round its ring more than once; tensor path and exact scan must agree, an image-sized sample goes to the oracle."""
    for Dm, K, B in ((128, 512, 48), (192, 256, 48)):
        z, E = vq_inputs(11, "D1b", B, Dm, 32, 32, K)          # 49,152 tokens = 1536 tiles > 148 x 8
        m = make(D.VectorQuantizer2, E, sane_index_shape=True)
        assert m.search_path() == "tcgen05"
        zc = z.to(DEV)
        with torch.no_grad():
            z_q, loss, (_, _, idx) = m(zc)
            m.search = "exact"
            z_q2, loss2, (_, _, idx2) = m(zc)
        differ = idx != idx2
        assert int(differ.sum()) <= 4                          # FP32 near-ties only (different summation trees)
        rows = zc.permute(0, 2, 3, 1).reshape(-1, Dm)
        picked = E.to(DEV)[idx.reshape(-1)]
        assert torch.equal(z_q.permute(0, 2, 3, 1).reshape(-1, Dm), rows + (picked - rows))
        ref = ((picked - rows) ** 2).mean()
        assert abs(float(loss) - float(ref + 0.25 * ref)) <= 1e-5 * float(ref)
        n_mis, n_out, _ = O.allowed_index_mismatch(z[B - 1:], E, idx[B - 1:].cpu())   # the last image: late tiles
        assert n_out == 0


def test_tensor_path_larger_codebooks():
    for K, Dm in ((2048, 128), (768, 64)):
        z, E = vq_inputs(5, "D1b", 1, Dm, 16, 16, K)
        m = make(D.VectorQuantizer2, E, sane_index_shape=True)
        assert m.search_path() == "tcgen05"
        with torch.no_grad():
            out = m(z.to(DEV))
        check_against_oracle(z, E, out, out[2][2], allow_near_ties=False)


@pytest.mark.parametrize("K,Dm", [(4096, 4), (2048, 8), (16384, 4)])
def test_narrow_large_codebooks(K, Dm):
    """The 8-threads-per-token variant of the narrow kernel (K >= 2048) against the oracle."""
    for kind in ("D0", "D1b"):
        z, E = vq_inputs(9, kind, 2, Dm, 12, 20, K)
        m = make(D.VectorQuantizer2, E, sane_index_shape=True)
        assert m.search_path() == "narrow-simt"
        zc = z.to(DEV).requires_grad_(True)
        out = m(zc)
        check_against_oracle(z, E, out, out[2][2], allow_near_ties=True)


@pytest.mark.parametrize("shape", [(1, 256, 64, 96), (2, 256, 9, 7), (3, 37, 8, 12)])
def test_decode_tokens_matches_reference_ops(shape):
    """argmax + accuracy + codebook gather of the decoder side (hyperprior_dc_vic_model.py:250-260)."""
    B, K, H, W = shape
    g = torch.Generator().manual_seed(21)
    logits = torch.randn(B, K, H, W, generator=g)
    logits[0, 5, 0, 0] = logits[0, :, 0, 0].max() + 1.0      # exact ties: the first maximal index must win
    logits[0, 9, 0, 0] = logits[0, 5, 0, 0]
    logits[-1, :, -1, -1] = 0.25                              # a whole row of equal values -> index 0
    E = torch.randn(K, 4, generator=g)
    gt = torch.randint(0, K, (B, H, W), generator=g)
    idx_r, lat_r, acc_r = O.decode_tokens(logits, E, gt)
    gt[idx_r % 3 == 0] = idx_r[idx_r % 3 == 0]                # make a third of them match
    idx_r, lat_r, acc_r = O.decode_tokens(logits, E, gt)
    idx, lat, acc = D.decode_tokens(logits.to(DEV), E.to(DEV), gt.to(DEV))
    assert idx.dtype == torch.int64 and torch.equal(idx.cpu(), idx_r)
    assert torch.equal(lat.cpu(), lat_r)
    assert abs(float(acc) - float(acc_r)) < 1e-7
    idx2, lat2, acc2 = D.decode_tokens(logits.to(DEV), E.to(DEV), None, want_latent=False)
    assert torch.equal(idx2.cpu(), idx_r) and lat2 is None and acc2 is None


@pytest.mark.parametrize("cb", ["default-init", "randn"])
def test_tiny_and_zero_tokens_on_the_tensor_path(cb):
    """Tokens with |z| << |e| (all-zero rows, |z| scaled by 1e-8 .. 1e-3), alone and mixed into ordinary tiles.
    Their scores z.e - |e|^2/2 are negative and decided by the folded -|e|^2/2 term: the cases the absolute
    margin term (three-way FP16 split on the 2^-24 grid) and the sign-aware chunk filter exist for."""
    kind = "D0" if cb == "default-init" else "D1b"
    z, E = vq_inputs(13, kind, 3, 256, 32, 32, 1024)
    z = z.clone()
    scales = [0.0, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3]
    for i, sc in enumerate(scales):
        z[0, :, i, :] *= sc                # 32 consecutive tokens (one finish tile) per magnitude
        z[0, :, 16 + i, 5] *= sc           # a single such token among ordinary ones
        z[1, :, 2 * i, ::2] *= sc          # every other token of a row
    z[2] *= 1e-6                           # a whole image of tiny tokens (every search tile row is tiny)
    z[2, :, :4] = 0.0                      # and 128 exact zeros
    for search in ("tensor", "exact"):
        m = make(D.VectorQuantizer2, E, sane_index_shape=True)
        m.search = search
        with torch.no_grad():
            out = m(z.to(DEV))
        n_mis, n_tie = check_against_oracle(z, E, out, out[2][2], allow_near_ties=True)
        print(f"tiny tokens, {cb} codebook, {search}: {n_mis} mismatches (all inside the clause), {n_tie} near-tie rows")


@pytest.mark.parametrize("kind", ["D0", "D1", "D1b"])
def test_c2_all_tokens_against_the_oracle(kind):
    """BASELINE config 2 at full size, every one of the 65,536 tokens against the CPU oracle (seed 0 for D0 is the
    batch bench.py times; D1 / D1b must have no mismatch at all outside the near-tie clause)."""
    z, E = vq_inputs(0 if kind == "D0" else 1, kind, 64, 256, 32, 32, 1024)
    m = make(D.VectorQuantizer2, E, sane_index_shape=True)
    m.search = "tensor"
    with torch.no_grad():
        z_q, loss, (_, _, idx) = m(z.to(DEV))
    n_mis, n_out, n_tie = O.allowed_index_mismatch(z, E, idx.cpu())
    print(f"C2 {kind}: {n_mis} index mismatches of 65536, {n_out} outside the near-tie clause, "
          f"{n_tie} rows whose reference top-2 gap is below 1e-6 relative")
    assert n_out == 0
    if kind != "D0":
        assert n_mis <= n_tie
    rows = O.token_rows(z)
    picked = E[idx.reshape(-1).cpu()]
    ste = (rows + (picked - rows)).view(64, 32, 32, 256).permute(0, 3, 1, 2)
    assert torch.equal(z_q.cpu(), ste)
    mse = ((picked - rows) ** 2).mean()
    assert abs(float(loss) - float(mse + 0.25 * mse)) <= 1e-5 * float(mse)


def test_decode_tokens_with_post_quant_conv_and_code_losses():
    """The rest of the decoder-side token path: lookup + post_quant_conv in the decode kernel
    (hyperprior_dc_vic_model.py:258-260) and the code CE / focal loss with its gradient
    (src/losses/cross_entropy_loss.py:9-52)."""
    g = torch.Generator().manual_seed(31)
    for (B, K, H, W, Dm) in ((1, 256, 64, 96, 4), (6, 256, 32, 32, 4), (2, 37, 9, 7, 8)):
        logits = 3 * torch.randn(B, K, H, W, generator=g)
        E = torch.randn(K, Dm, generator=g)
        gt = torch.randint(0, K, (B, H, W), generator=g)
        pq = torch.nn.Conv2d(Dm, Dm, 1)
        with torch.no_grad():
            pq.weight.copy_(torch.randn(Dm, Dm, 1, 1, generator=g))
            pq.bias.copy_(torch.randn(Dm, generator=g))
        idx_r = torch.argmax(logits, 1)
        lat_r = O.post_quant_latent(idx_r, E, pq.weight.detach(), pq.bias.detach())
        idx, lat, acc = D.decode_tokens(logits.to(DEV), E.to(DEV), gt.to(DEV), post_quant_conv=pq.to(DEV))
        assert torch.equal(idx.cpu(), idx_r)
        # D fused multiply-adds per output against the CPU conv of the oracle, whose summation order depends on the
        # host's thread count / oneDNN kernel choice (seen to differ by a few 1e-6 once in ~50 runs): FP32 sum of <= 8
        # O(1) products
        assert torch.allclose(lat.cpu(), lat_r, rtol=1e-5, atol=1e-5), float((lat.cpu() - lat_r).abs().max())
        assert abs(float(acc) - float((idx_r == gt).float().mean())) < 1e-7
        for gamma, red in ((0.0, "mean"), (2.0, "mean"), (1.0, "sum"), (0.5, "mean")):
            lr = logits.clone().requires_grad_(True)
            if gamma == 0.0:
                ref = O.code_cross_entropy(lr, gt, 0.7)
                ours_m = D.CrossEntropyLoss(0.7)
            else:
                ref = O.code_focal_cross_entropy(lr, gt, 0.7, gamma, red)
                ours_m = D.FocalCrossEntropyLoss(0.7, gamma, red)
            ref.backward()
            lo = logits.to(DEV).requires_grad_(True)
            ours = ours_m(lo, gt.to(DEV))
            (2.0 * ours).backward()
            assert abs(float(ours) - float(ref)) <= 2e-5 * abs(float(ref)), (gamma, red, float(ours), float(ref))
            scale = float(lr.grad.abs().max())
            assert float((lo.grad.cpu() - 2.0 * lr.grad).abs().max()) <= 2e-4 * scale + 1e-9
