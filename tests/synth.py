"""Seeded synthetic inputs shared by tests, bench.py and tests/golden/make_golden.py."""
import hashlib

import torch


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def vq_inputs(seed: int, kind: str, B: int, D: int, H: int, W: int, K: int):
    """D0: default-init codebook U(-1/K,1/K), z~N(0,1).  D1: z = E[j] + 0.5 N(0,1) (well separated).
    D1b: E~N(0,1), z~N(0,1) (adversarial for low-precision products)."""
    g = torch.Generator().manual_seed(seed)
    if kind == "D0":
        E = torch.empty(K, D).uniform_(-1.0 / K, 1.0 / K, generator=g)
        z = torch.randn(B, D, H, W, generator=g)
    elif kind == "D1":
        E = torch.randn(K, D, generator=g)
        pick = torch.randint(0, K, (B * H * W,), generator=g)
        rows = E[pick] + 0.5 * torch.randn(B * H * W, D, generator=g)
        z = rows.view(B, H, W, D).permute(0, 3, 1, 2).contiguous()
    elif kind == "D1b":
        E = torch.randn(K, D, generator=g)
        z = torch.randn(B, D, H, W, generator=g)
    else:
        raise ValueError(kind)
    return z, E


def entropy_inputs(q: int, B: int = 64, C: int = 320, H: int = 32, W: int = 32):
    """C3 latents: mu~N(0,1), sigma_raw = exp(m_q + N(0,1)), y = mu + max(sigma,0.11) N(0,1)."""
    g = torch.Generator().manual_seed(100 + q)
    m_q = (-2.0, -1.5, -1.0, -0.5, 0.0)[q]
    mu = torch.randn(B, C, H, W, generator=g)
    sigma_raw = torch.exp(m_q + torch.randn(B, C, H, W, generator=g))
    y = mu + sigma_raw.clamp_min(0.11) * torch.randn(B, C, H, W, generator=g)
    return y, torch.cat([mu, sigma_raw], dim=1)


def entropy_inputs_init(B: int = 8, C: int = 32, H: int = 16, W: int = 16):
    """'q-init': conv-like raw scales that can be <= 0 and rely on LowerBound(0.11)."""
    g = torch.Generator().manual_seed(110)
    mu = 0.3 * torch.randn(B, C, H, W, generator=g)
    sigma_raw = 0.3 * torch.randn(B, C, H, W, generator=g)
    y = torch.randn(B, C, H, W, generator=g)
    return y, torch.cat([mu, sigma_raw], dim=1)


def noise_like(t: torch.Tensor, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(t.shape, generator=g) - 0.5
