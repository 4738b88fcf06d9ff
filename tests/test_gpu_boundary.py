"""GPU side of the drop-in boundary (no /root/reference here: the wrapper classes below restate, line for line in
behaviour, src/models/subnet/entropy_model/entropy_bottleneck.py:13-28 and gaussian_conditional.py:17-24 on top of
this package's CompressAI-surface base classes, which is what the reference's files become under the shim)."""
import pytest
import torch

import dc_vic_b200 as D
from dc_vic_b200 import entropy_models as EM
from dc_vic_b200.register import fused_subclass
from oracle import entropy_oracle as O
from synth import entropy_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class RefLikeEntropyBottleneck(EM.EntropyBottleneck):            # entropy_bottleneck.py:13-16
    def forward(self, x, is_train):
        return super().forward(x, training=is_train)


class RefLikeGaussianMeanScale(EM.GaussianConditional):          # gaussian_conditional.py:17-24
    def __init__(self, scale_bound=None):
        super().__init__(scale_table=None, scale_bound=scale_bound)

    def forward(self, y, params, is_train=True):
        mean, std = params.chunk(2, 1)
        return super().forward(y, scales=std, means=mean, training=is_train)


def test_fused_subclass_of_a_reference_wrapper_matches_it_and_the_oracle():
    torch.manual_seed(0)
    cls = fused_subclass("EntropyBottleneck", RefLikeEntropyBottleneck)
    assert issubclass(cls, RefLikeEntropyBottleneck) and cls.__name__ == "EntropyBottleneck"
    fused, plain = cls(channels=8).to(DEV), RefLikeEntropyBottleneck(channels=8).to(DEV)
    plain.load_state_dict(fused.state_dict())
    ref = O.EntropyBottleneck(channels=8)
    ref.load_state_dict({k: v.cpu() for k, v in fused.state_dict().items()})
    x = 3 * torch.randn(2, 8, 4, 6)
    with torch.no_grad():
        a, b, c = fused(x.to(DEV), False), plain(x.to(DEV), False), ref(x, training=False)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert torch.equal(a[0].cpu(), c[0]) and float(((a[1].cpu() - c[1]).abs() / c[1]).max()) < 1e-4
    assert float(fused.loss()) > 0

    gcls = fused_subclass("GaussianMeanScaleConditional", RefLikeGaussianMeanScale)
    g, gp = gcls(scale_bound=0.11).to(DEV), RefLikeGaussianMeanScale(scale_bound=0.11).to(DEV)
    y, params = entropy_inputs(1, B=2, C=16, H=8, W=8)
    with torch.no_grad():
        u, v = g(y.to(DEV), params.to(DEV), is_train=False), gp(y.to(DEV), params.to(DEV), is_train=False)
    assert torch.equal(u[0], v[0]) and torch.equal(u[1], v[1])


def test_entropy_models_stay_on_the_gpu_and_take_cpu_tensors():
    """codec_setup moves the entropy models to the CPU and feeds them CPU tensors
    (hyperprior_dc_vic_model.py:65-73,308-328): the modules stay on their GPU, compute there, answer on the CPU."""
    eb = D.SteEntropyBottleneck(channels=8).to(DEV)
    gc = D.SteGaussianMeanScaleConditional(scale_bound=0.11).to(DEV)
    assert eb.to("cpu") is eb and gc.cpu() is gc
    assert eb.quantiles.is_cuda and gc.scale_bound.is_cuda
    x = 3 * torch.randn(2, 8, 4, 4)
    y, params = entropy_inputs(1, B=2, C=16, H=8, W=8)
    with torch.no_grad():
        xh_c, l_c = eb(x, is_train=False)                      # CPU in -> CPU out
        xh_g, l_g = eb(x.to(DEV), is_train=False)
        yh_c, k_c = gc(y, params, is_train=False)
        yh_g, k_g = gc(y.to(DEV), params.to(DEV), is_train=False)
    assert not xh_c.is_cuda and not l_c.is_cuda and not yh_c.is_cuda
    assert torch.equal(xh_c, xh_g.cpu()) and torch.equal(l_c, l_g.cpu())
    assert torch.equal(yh_c, yh_g.cpu()) and torch.equal(k_c, k_g.cpu())
    gc.update_scale_table(D.get_scale_table(), force=True)
    idx = gc.build_indexes(params.chunk(2, 1)[1])             # CPU scales
    assert not idx.is_cuda and idx.dtype == torch.int32
    og = O.SteGaussianMeanScaleConditional(scale_bound=0.11)
    og.update_scale_table(O.get_scale_table(), force=True)
    assert torch.equal(idx, og.build_indexes(params.chunk(2, 1)[1]).int())
    q = gc.quantize(y, "symbols", params.chunk(2, 1)[0])
    assert not q.is_cuda and q.dtype == torch.int32
    # dtype conversions still go through, explicit opt-out restores torch's behaviour
    eb2 = D.SteEntropyBottleneck(channels=4).to(DEV)
    eb2.allow_cpu_move = True
    eb2.to("cpu")
    assert not eb2.quantiles.is_cuda


def test_swap_quantizer_freezes_a_frozen_codebook_and_reuses_the_preparation():
    class FakeRef(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.n_e, self.e_dim, self.beta, self.legacy, self.sane_index_shape, self.remap = 1024, 256, 0.25, True, True, None
            self.embedding = torch.nn.Embedding(1024, 256)
    holder = torch.nn.Module()
    holder.quantize = FakeRef().to(DEV)
    holder.requires_grad_(False)
    D.swap_quantizer(holder)
    q = holder.quantize
    assert q.codebook_frozen and q.search_path() == "tcgen05"
    z = torch.randn(2, 256, 16, 16, device=DEV)
    with torch.no_grad():
        a = q(z)
        assert q._prep_key is not None
        b = q(z)                                               # second call: REUSE_PREP
        assert torch.equal(a[0], b[0]) and torch.equal(a[2][2], b[2][2])
        q.embedding.weight.data.mul_(-1.0)                     # a .data write is invisible to the version check ...
        q.invalidate_codebook()                                # ... hence the explicit invalidation
        c = q(z)
    ref = D.VectorQuantizer2(1024, 256, 0.25, sane_index_shape=True).to(DEV)
    ref.embedding.weight.data.copy_(q.embedding.weight.data)
    with torch.no_grad():
        d = ref(z)
    assert torch.equal(c[2][2], d[2][2]) and torch.equal(c[0], d[0])
    with pytest.raises(NotImplementedError):
        D.VectorQuantizer2(16, 4, 0.25, remap="used.npy")


def test_codebook_entry_is_differentiable_like_nn_embedding():
    m = D.VectorQuantizer2(32, 4, 0.25).to(DEV)
    idx = torch.randint(0, 32, (2, 5, 7), device=DEV)
    out = m.get_codebook_entry(idx, (2, 5, 7, 4))
    g = torch.randn_like(out)
    (out * g).sum().backward()
    ref_w = m.embedding.weight.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.embedding(idx, ref_w).permute(0, 3, 1, 2)
    (ref * g).sum().backward()
    assert torch.equal(out.detach(), ref.detach())
    assert torch.allclose(m.embedding.weight.grad, ref_w.grad, rtol=1e-6, atol=1e-7)
