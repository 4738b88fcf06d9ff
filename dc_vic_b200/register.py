"""Plugging dc_vic_b200 into an unmodified DC-VIC checkout.

DC-VIC has two plugin boundaries on this path:

1. entropy models are built BY NAME from ``ENTROPYMODEL_REGISTRY`` (src/utils/registry.py:83)
   via ``build_subnet(opt, 'entropy_model')`` (src/models/subnet/__init__.py:18-32), and
   ``base_model.py:76-104,128-130`` does ``isinstance(m, EntropyBottleneck)`` against the REFERENCE'S OWN wrapper
   class (src/models/subnet/entropy_model/entropy_bottleneck.py:13) and
   ``isinstance(m, compressai.entropy_models.GaussianConditional)``;
2. the VQ quantizer is the attribute ``vq_model.quantize`` (ldm/models/autoencoder.py:39-41).

``install_compressai_shim()`` makes every ``compressai`` import of the reference's ``src/`` tree resolve to this
package when no real compressai is installed (entropy_models, ops, models[.utils|.google], layers.GDN, ans), so the
reference's wrapper files subclass the CUDA-backed base classes unchanged.
``register_entropy_models(registry)`` then replaces each registered wrapper by a SUBCLASS of it whose ``forward`` is
the fused one-kernel implementation (``type(name, (fused_forward_mixin, reference_class), {})``): class identity,
state-dict keys, ``aux_loss()``, the CDF-buffer resize in ``load_state_dict`` and ``update()`` after loading all keep
working, with a real compressai installed as well as with the shim.
``dc_vic_b200.swap_quantizer(model.vq_model)`` covers boundary 2.
"""
from __future__ import annotations

import sys
import types

from . import entropy_models as _em

ENTROPY_MODEL_CLASSES = {
    "EntropyBottleneck": _em.DcvicEntropyBottleneck,
    "SteEntropyBottleneck": _em.SteEntropyBottleneck,
    "GaussianScaleConditional": _em.GaussianScaleConditional,
    "GaussianMeanScaleConditional": _em.GaussianMeanScaleConditional,
    "SteGaussianMeanScaleConditional": _em.SteGaussianMeanScaleConditional,
}


def _update_registered_buffers(module, module_name, buffer_names, state_dict, policy="resize_if_empty",
                               dtype=None):
    """compressai.models.utils.update_registered_buffers: resize CDF buffers to the checkpoint's
    shapes before load_state_dict (base_model.py:88-104)."""
    import torch
    dtype = dtype or torch.int
    valid = [n for n, _ in module.named_buffers()]
    for name in buffer_names:
        if name not in valid:
            raise ValueError(f'Invalid buffer name "{name}"')
    for name in buffer_names:
        key = f"{module_name}.{name}"
        if key not in state_dict:
            continue
        new_size = state_dict[key].size()
        buf = getattr(module, name)
        if policy in ("resize_if_empty", "resize"):
            if policy == "resize" or buf.numel() == 0:
                buf.resize_(new_size)
        elif policy == "register":
            module.register_buffer(name, torch.empty(new_size, dtype=dtype).fill_(0))
        else:
            raise ValueError(f'Invalid policy "{policy}"')


def install_compressai_shim(force: bool = False) -> types.ModuleType:
    """Expose the subset of the ``compressai`` import surface DC-VIC uses on this path, backed by
    dc_vic_b200 (entropy_models, ops.LowerBound, models.utils.update_registered_buffers,
    models.google.get_scale_table).  No-op if a real compressai is importable, unless force."""
    if not force:
        try:
            import compressai  # noqa: F401
            if not getattr(sys.modules["compressai"], "__dcvic_b200_shim__", False):
                return sys.modules["compressai"]
        except ImportError:
            pass
    root = types.ModuleType("compressai")
    root.__dcvic_b200_shim__ = True
    root.__path__ = []
    ent = types.ModuleType("compressai.entropy_models")
    for name in ("EntropyModel", "EntropyBottleneck", "GaussianConditional"):
        setattr(ent, name, getattr(_em, name))
    ops = types.ModuleType("compressai.ops")
    ops.LowerBound = _em.LowerBound
    models = types.ModuleType("compressai.models")
    models.__path__ = []
    utils = types.ModuleType("compressai.models.utils")
    utils.update_registered_buffers = _update_registered_buffers
    google = types.ModuleType("compressai.models.google")
    google.get_scale_table = _em.get_scale_table
    models.get_scale_table = _em.get_scale_table          # `from compressai.models import get_scale_table`
    layers = types.ModuleType("compressai.layers")        # (hyperprior_vic_model.py:12, hyperprior_dc_vic_model.py:10)
    layers.__path__ = []
    layers.GDN = _gdn_class()                             # cheng_resblock.py:8, balle18_autoencoder.py:5
    from . import rans as _rans                           # minnen20_charm_context_model.py:12
    ans = types.ModuleType("compressai.ans")
    for name in ("RansEncoder", "RansDecoder", "BufferedRansEncoder"):
        setattr(ans, name, getattr(_rans, name))
    root.entropy_models, root.ops, root.models, root.layers, root.ans = ent, ops, models, layers, ans
    root.available_entropy_coders = lambda: ["ans"]
    models.utils, models.google = utils, google
    for m in (root, ent, ops, models, utils, google, layers, ans):
        sys.modules[m.__name__] = m
    return root


def _gdn_class():
    """compressai.layers.GDN (generalized divisive normalization) for the shim.  NOT on the hot path and not used by
    DC-VIC's own configs (ELIC encoder); it only has to exist because src/models/layer/cheng_resblock.py imports it at
    module level.  Plain torch, CompressAI's parameter names (beta, gamma + NonNegativeParametrizer reparametrisation)."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    class _NonNegative(nn.Module):
        def __init__(self, minimum: float = 0.0, reparam_offset: float = 2 ** -18):
            super().__init__()
            pedestal = float(reparam_offset) ** 2
            self.register_buffer("pedestal", torch.Tensor([pedestal]))
            self.lower_bound = _em.LowerBound((float(minimum) + pedestal) ** 0.5)

        def init(self, x):
            return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

        def forward(self, x):
            return self.lower_bound(x) ** 2 - self.pedestal

    class GDN(nn.Module):
        def __init__(self, in_channels: int, inverse: bool = False, beta_min: float = 1e-6, gamma_init: float = 0.1):
            super().__init__()
            self.inverse = bool(inverse)
            self.beta_reparam = _NonNegative(minimum=beta_min)
            self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
            self.gamma_reparam = _NonNegative()
            self.gamma = nn.Parameter(self.gamma_reparam.init(gamma_init * torch.eye(in_channels)))

        def forward(self, x):
            C = x.shape[1]
            beta = self.beta_reparam(self.beta)
            gamma = self.gamma_reparam(self.gamma).reshape(C, C, 1, 1)
            norm = torch.sqrt(F.conv2d(x ** 2, gamma, beta)) if self.inverse else torch.rsqrt(F.conv2d(x ** 2, gamma, beta))
            return x * norm

    return GDN


def fused_subclass(name: str, reference_class):
    """The class ``register_entropy_models`` installs for ``name``: the fused forward in front of the reference's
    own wrapper class (so it IS-A reference wrapper and IS-A compressai entropy model)."""
    mixin = _em.FUSED_FORWARDS[name]
    if issubclass(reference_class, mixin):
        return reference_class
    return type(name, (mixin, reference_class), {
        "__module__": reference_class.__module__,
        "__doc__": f"{reference_class.__module__}.{name} with dc_vic_b200's fused CUDA forward ({mixin.__name__}).",
        "__dcvic_b200_fused__": True,
    })


def register_entropy_models(registry) -> None:
    """Point the reference's ``ENTROPYMODEL_REGISTRY`` names at the fused CUDA forwards.
    ``registry`` is ``src.utils.registry.ENTROPYMODEL_REGISTRY`` (a name -> {'obj','filename'} map), AFTER
    ``import src.models.subnet.entropy_model`` has registered the reference's own wrappers: each of them is replaced
    by a subclass of itself (``fused_subclass``).  Names the registry does not hold yet get this package's
    stand-alone classes."""
    for name, cls in ENTROPY_MODEL_CLASSES.items():
        entry = registry._obj_map.get(name)
        if entry is None:
            registry._obj_map[name] = {"obj": cls, "filename": "dc_vic_b200/entropy_models.py"}
        else:
            registry._obj_map[name] = {"obj": fused_subclass(name, entry["obj"]), "filename": entry["filename"]}
