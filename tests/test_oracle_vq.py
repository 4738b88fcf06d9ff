"""Pin the VQ oracle against golden vectors produced by the reference's own quantize.py.

Goldens: tests/golden/vq_*.pt (generator: tests/golden/make_golden.py, which imports
/root/reference/taming/modules/vqvae/quantize.py).  CPU only.
"""
import glob
import os

import pytest
import torch

from oracle import vq_oracle as O
from synth import sha, vq_inputs

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return torch.load(os.path.join(G, name), weights_only=False)


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)   # goldens were made single-threaded (SGEMM blocking)
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("name", ["vq_v2_narrow_default.pt", "vq_v2_narrow_randn_legacy0.pt",
                                  "vq_v2_narrow_randn_legacy1.pt", "vq_v2_d8_ragged.pt"])
def test_v2_forward_backward_matches_reference(name):
    fx = load(name)
    z = fx["z"].clone().requires_grad_(True)
    E = fx["E"].clone().requires_grad_(True)
    out = O.vq2_forward(z, E, beta=fx["beta"], legacy=fx["legacy"], sane_index_shape=fx["sane"])
    assert out.perplexity is None and out.min_encodings is None
    assert out.indices.dtype == torch.int64
    assert torch.equal(out.indices.to(torch.int32), fx["idx"])
    assert torch.equal(out.z_q, fx["z_q"])                      # bit-exact
    assert torch.equal(out.loss, fx["loss"])
    ((out.z_q * fx["g_zq"]).sum() + fx["g_loss"] * out.loss).backward()
    assert torch.allclose(z.grad, fx["dz"], rtol=0, atol=1e-7)
    assert torch.allclose(E.grad, fx["dE"], rtol=1e-6, atol=1e-8)
    assert float(fx["dE_from_zq"].abs().max()) == 0.0           # STE sends no g_zq to the codebook


def test_v1_contract_matches_reference():
    fx = load("vq_v1_small.pt")
    z = fx["z"].clone().requires_grad_(True)
    E = fx["E"].clone().requires_grad_(True)
    out = O.vq1_forward(z, E, beta=fx["beta"])
    assert out.indices.shape == (z.shape[0] * z.shape[2] * z.shape[3], 1)
    assert torch.equal(out.indices.squeeze(1).to(torch.int32), fx["idx"].squeeze(1))
    assert torch.equal(out.min_encodings.argmax(1).to(torch.int32), fx["onehot_idx"])
    assert torch.equal(out.min_encodings.sum(1), fx["onehot_rowsum"])
    assert torch.equal(out.z_q, fx["z_q"])
    assert torch.equal(out.loss, fx["loss"])
    assert torch.equal(out.perplexity, fx["perplexity"])
    ((out.z_q * fx["g_zq"]).sum() + fx["g_loss"] * out.loss).backward()
    assert torch.allclose(z.grad, fx["dz"], rtol=0, atol=1e-7)
    assert torch.allclose(E.grad, fx["dE"], rtol=1e-6, atol=1e-8)


def test_codebook_entry_matches_reference():
    fx = load("vq_codebook_entry.pt")
    got = O.codebook_entry(fx["pick"].long(), fx["E"], fx["shape"])
    assert torch.equal(got, fx["entry"])
    assert torch.equal(O.codebook_entry(fx["pick"].long(), fx["E"], None), fx["entry_flat"])
    b, h, w, _ = fx["shape"]
    assert torch.equal(O.indices_to_latent(fx["pick"].long().view(b, h, w), fx["E"]), fx["entry"])


@pytest.mark.parametrize("kind", ["D0", "D1", "D1b"])
def test_wide_codebook_matches_reference(kind):
    fx = load(f"vq_v2_wide_{kind}.pt")
    B, D, H, W, K = fx["shape"]
    z, E = vq_inputs(fx["seed"], kind, B, D, H, W, K)
    assert sha(z) == fx["z_sha"] and sha(E) == fx["E_sha"], "seeded inputs drifted: regenerate goldens"
    out = O.vq2_forward(z, E, beta=fx["beta"], legacy=fx["legacy"], sane_index_shape=fx["sane"])
    assert torch.equal(out.indices.to(torch.int32), fx["idx"])
    assert torch.equal(out.z_q[:, :, :2, :], fx["z_q"])
    assert torch.equal(out.loss, fx["loss"])
    gap = O.top2_relative_gap(O.distances(O.token_rows(z), E))
    assert torch.allclose(gap, fx["rel_gap"], rtol=1e-3, atol=2e-7)   # diagnostic only
    n_mis, n_out, _ = O.allowed_index_mismatch(z, E, out.indices)
    assert (n_mis, n_out) == (0, 0)


def test_parity_rule_flags_real_mismatch():
    z, E = vq_inputs(1, "D1b", 1, 16, 4, 4, 32)
    idx = O.vq2_forward(z, E).indices.clone()
    idx[0] = (idx[0] + 1) % 32
    n_mis, n_out, _ = O.allowed_index_mismatch(z, E, idx)
    assert n_mis == 1 and n_out == 1


def test_onehot_feature():
    idx = torch.tensor([[[0, 2], [1, 3]]])
    f = O.onehot_feature(idx, 4)
    assert f.shape == (1, 4, 2, 2) and f.dtype == torch.float32
    assert torch.equal(f.argmax(1), idx) and float(f.sum()) == 4.0


def test_goldens_do_not_need_the_reference_tree():
    assert len(glob.glob(os.path.join(G, "vq_*.pt"))) >= 9


def test_decode_tokens_oracle_is_argmax_gather():
    g = torch.Generator().manual_seed(2)
    logits = torch.randn(2, 16, 3, 5, generator=g)
    E = torch.randn(16, 4, generator=g)
    idx, lat, acc = O.decode_tokens(logits, E, torch.zeros(2, 3, 5, dtype=torch.long))
    assert torch.equal(idx, logits.argmax(1)) and lat.shape == (2, 4, 3, 5)
    assert torch.equal(lat[1, :, 2, 4], E[idx[1, 2, 4]]) and 0.0 <= float(acc) <= 1.0
