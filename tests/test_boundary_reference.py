"""The drop-in boundary against the reference's OWN classes (CPU, needs /root/reference: skipped elsewhere).

INTEGRATION.md section 2, verbatim: install the compressai shim, import the reference's entropy-model wrappers and
`base_model.py`, re-register the fused classes, and check what `BaseModel` relies on (base_model.py:76-104,128-146,
registry.py:37-44): class identity for the isinstance checks, a non-zero `aux_loss()`, the CDF-buffer resize path of
`load_state_dict`, `update()` after loading, the `.quantiles` parameter split and the state-dict keys."""
import sys

import pytest
import torch

import dc_vic_b200 as dcv
from ref_shims import have_reference, import_reference

pytestmark = pytest.mark.skipif(not have_reference(), reason="/root/reference is not present on this box")


@pytest.fixture(scope="module")
def ref():
    import_reference()
    # ---- INTEGRATION.md section 2 ----
    from src.utils.registry import ENTROPYMODEL_REGISTRY
    import src.models.subnet.entropy_model            # the reference's own registration
    dcv.register_entropy_models(ENTROPYMODEL_REGISTRY)
    # ----------------------------------
    import src.models.comp_model.base_model as base_model
    return ENTROPYMODEL_REGISTRY, base_model


def _build(registry, name, **kw):
    return registry.get(name)(**kw)


def test_every_compressai_import_of_the_reference_resolves(ref):
    import importlib
    for mod in ("src.models.comp_model.hyperprior_vic_model", "src.models.comp_model.hyperprior_dc_vic_model",
                "src.models.comp_model.hyperprior_charm_dc_vic_model",
                "src.models.subnet.context_model.minnen20_charm_context_model", "src.models.layer.cheng_resblock",
                "src.models.subnet.autoencoder.balle18_autoencoder"):
        importlib.import_module(mod)
    import compressai
    from compressai.models import get_scale_table
    from compressai.layers import GDN
    from compressai.ans import RansDecoder, RansEncoder, BufferedRansEncoder   # noqa: F401
    assert getattr(compressai, "__dcvic_b200_shim__", False)
    assert torch.allclose(get_scale_table()[[0, -1]], torch.tensor([0.11, 256.0]))
    y = GDN(4)(torch.randn(1, 4, 3, 3))
    assert y.shape == (1, 4, 3, 3) and bool(torch.isfinite(y).all())


def test_registry_classes_keep_the_reference_identity(ref):
    registry, base_model = ref
    from src.models.subnet.entropy_model.entropy_bottleneck import EntropyBottleneck as RefEB, \
        SteEntropyBottleneck as RefSteEB
    from src.models.subnet.entropy_model.gaussian_conditional import GaussianMeanScaleConditional as RefGMS
    from src.models.subnet.entropy_model.ste_gaussian_conditional import SteGaussianMeanScaleConditional as RefSte
    from compressai.entropy_models import GaussianConditional, EntropyBottleneck as CaiEB
    z = _build(registry, "SteEntropyBottleneck", channels=192)          # ...vq_f8_n256.yaml:53-58
    y = _build(registry, "SteGaussianMeanScaleConditional", scale_bound=0.11)
    assert isinstance(z, RefSteEB) and isinstance(z, RefEB) and isinstance(z, CaiEB)
    assert isinstance(z, base_model.EntropyBottleneck)                   # the check base_model.py:76-86 makes
    assert isinstance(y, RefSte) and isinstance(y, RefGMS) and isinstance(y, GaussianConditional)
    assert isinstance(y, base_model.GaussianConditional)
    # ... and the forward that runs is the fused one
    assert type(z).forward is dcv.entropy_models.FusedSteBottleneckForward.forward
    assert type(y).forward is dcv.entropy_models._FusedGaussianForward.forward
    assert type(z).__dcvic_b200_fused__ and type(z).__name__ == "SteEntropyBottleneck"
    for name in ("EntropyBottleneck", "GaussianScaleConditional", "GaussianMeanScaleConditional"):
        assert registry.get(name).__dcvic_b200_fused__
    # registering twice is idempotent
    dcv.register_entropy_models(registry)
    assert registry.get("SteEntropyBottleneck") is type(z)
    # the unrelated reference class is untouched
    assert not getattr(registry.get("VqCategoricalEntropyModel"), "__dcvic_b200_fused__", False)


class _Opt(dict):
    device = "cpu"

    def get(self, k, d=None):
        return d


def _tiny_model(ref):
    registry, base_model = ref

    class Tiny(base_model.BaseModel):
        def _build_subnets(self):
            self.entropy_model_z = _build(registry, "SteEntropyBottleneck", channels=6)
            self.entropy_model_y = _build(registry, "SteGaussianMeanScaleConditional", scale_bound=0.11)

    return Tiny(_Opt())


def test_base_model_paths_see_the_fused_modules(ref, tmp_path):
    m = _tiny_model(ref)
    # aux_loss (base_model.py:76-86) sums EntropyBottleneck.loss() over modules found by isinstance
    aux = m.aux_loss()
    assert isinstance(aux, torch.Tensor) and float(aux) > 0.0
    aux.backward()
    assert m.entropy_model_z.quantiles.grad is not None and float(m.entropy_model_z.quantiles.grad.abs().sum()) > 0
    assert all(p.grad is None for n, p in m.named_parameters() if not n.endswith(".quantiles"))
    # state-dict keys are CompressAI's
    keys = set(m.state_dict())
    for k in ("entropy_model_z.quantiles", "entropy_model_z._matrix0", "entropy_model_z._bias4",
              "entropy_model_z._factor3", "entropy_model_z.target", "entropy_model_z._offset",
              "entropy_model_z._quantized_cdf", "entropy_model_z._cdf_length",
              "entropy_model_z.likelihood_lower_bound.bound", "entropy_model_y.scale_table",
              "entropy_model_y.scale_bound", "entropy_model_y.lower_bound_scale.bound",
              "entropy_model_y._quantized_cdf"):
        assert k in keys, k
    # aux optimizer split (base_model.py:132-146)
    aux_names = {n for n, p in m.named_parameters() if n.endswith(".quantiles")}
    assert aux_names == {"entropy_model_z.quantiles"}
    # load_state_dict resizes the (empty) CDF buffers to the checkpoint's shapes (base_model.py:88-104)
    src = _tiny_model(ref)
    assert src.entropy_model_z.update(force=True)
    sd = src.state_dict()
    sd["entropy_model_y._quantized_cdf"] = torch.zeros(64, 7, dtype=torch.int32)
    sd["entropy_model_y._cdf_length"] = torch.full((64,), 7, dtype=torch.int32)
    sd["entropy_model_y._offset"] = torch.zeros(64, dtype=torch.int32)
    sd["entropy_model_y.scale_table"] = dcv.get_scale_table()
    assert m.entropy_model_z._quantized_cdf.numel() == 0
    m.load_state_dict(sd)
    assert m.entropy_model_z._quantized_cdf.shape == src.entropy_model_z._quantized_cdf.shape
    assert torch.equal(m.entropy_model_z._quantized_cdf, src.entropy_model_z._quantized_cdf)
    assert m.entropy_model_y._quantized_cdf.shape == (64, 7) and m.entropy_model_y.scale_table.numel() == 64
    # load_learned_weight -> update(force=False) on EntropyBottleneck children (base_model.py:106-130)
    ckpt = tmp_path / "ckpt.pth.tar"
    fresh = _tiny_model(ref)
    torch.save({"comp_model": {k: v for k, v in fresh.state_dict().items()}}, ckpt)
    other = _tiny_model(ref)
    other.load_learned_weight(str(ckpt))
    assert other.entropy_model_z._quantized_cdf.numel() > 0      # update() ran: tables were built after loading
    table = other.entropy_model_z._quantized_cdf
    lens = other.entropy_model_z._cdf_length
    for c in range(table.shape[0]):
        row = table[c, : int(lens[c])]
        assert int(row[0]) == 0 and int(row[-1]) == 1 << 16 and bool((row[1:] > row[:-1]).all())


def test_swap_quantizer_on_the_vendored_class(ref):
    from taming.modules.vqvae.quantize import VectorQuantizer2 as RefVQ
    holder = torch.nn.Module()
    holder.quantize = RefVQ(256, 4, beta=0.25, remap=None, sane_index_shape=False)
    holder.quantize.sane_index_shape = True                        # hyperprior_vic_model.py:61
    holder.requires_grad_(False)                                   # rate_distortion_vq_code_trainer.py:62
    w = holder.quantize.embedding.weight.data.clone()
    dcv.swap_quantizer(holder)
    q = holder.quantize
    assert isinstance(q, dcv.VectorQuantizer2) and q.sane_index_shape and q.legacy and q.beta == 0.25
    assert (q.n_e, q.e_dim) == (256, 4) and torch.equal(q.embedding.weight.data, w)
    assert list(holder.state_dict()) == ["quantize.embedding.weight"]
    assert not q.embedding.weight.requires_grad and q.codebook_frozen
