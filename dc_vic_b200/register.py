"""Plugging dc_vic_b200 into an unmodified DC-VIC checkout.

DC-VIC has two plugin boundaries on this path:

1. entropy models are built BY NAME from ``ENTROPYMODEL_REGISTRY`` (src/utils/registry.py:83)
   via ``build_subnet(opt, 'entropy_model')`` (src/models/subnet/__init__.py:18-32), and
   ``base_model.py:76-104`` does ``isinstance(m, compressai.entropy_models.EntropyBottleneck)``;
2. the VQ quantizer is the attribute ``vq_model.quantize`` (ldm/models/autoencoder.py:39-41).

``install_compressai_shim()`` makes ``import compressai.entropy_models`` resolve to this
package, so the reference's own wrapper files (src/models/subnet/entropy_model/*.py) subclass
the CUDA-backed classes unchanged and every isinstance check keeps working.
``register_entropy_models(registry)`` additionally overrides the registry names with the fused
wrappers of ``dc_vic_b200.entropy_models`` (one kernel per call instead of wrapper + torch ops).
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
    root.entropy_models, root.ops, root.models = ent, ops, models
    models.utils, models.google = utils, google
    for m in (root, ent, ops, models, utils, google):
        sys.modules[m.__name__] = m
    return root


def register_entropy_models(registry) -> None:
    """Point the reference's ``ENTROPYMODEL_REGISTRY`` names at the fused CUDA wrappers.
    ``registry`` is ``src.utils.registry.ENTROPYMODEL_REGISTRY`` (a name -> {'obj','filename'} map)."""
    for name, cls in ENTROPY_MODEL_CLASSES.items():
        registry._obj_map[name] = {"obj": cls, "filename": "dc_vic_b200/entropy_models.py"}
