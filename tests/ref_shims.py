"""TEST INFRASTRUCTURE: import stubs (no arithmetic in them) for the third-party packages the reference's own
`src/` tree imports but this image does not have (python_log_indenter, timm, pytorch_lightning,
pytorch_msssim, lpips, addict, omegaconf, skimage).  With them - and with `compressai` provided by
`dc_vic_b200.install_compressai_shim()` - `/root/reference/src` imports unmodified, so the boundary tests can build
the reference's own model classes around the CUDA-backed modules.  Only used when /root/reference exists (this
container); nothing under `-m gpu` needs it."""
import logging
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def _module(name: str, **attrs) -> types.ModuleType:
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []
        m.__dcvic_test_stub__ = True
        sys.modules[name] = m
        parent, _, leaf = name.rpartition(".")
        if parent:
            setattr(_module(parent), leaf, m)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _missing(name: str) -> bool:
    if name in sys.modules:
        return False
    import importlib.util
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def install_third_party_stubs() -> None:
    import torch
    import torch.nn as nn

    if _missing("python_log_indenter"):
        class IndentedLoggerAdapter(logging.LoggerAdapter):
            def __init__(self, logger, *a, **k):
                super().__init__(logger, {})

            def add(self, *a, **k):
                return self

            def sub(self, *a, **k):
                return self

            push = pop = add

        _module("python_log_indenter", IndentedLoggerAdapter=IndentedLoggerAdapter)

    if _missing("pytorch_msssim"):
        def _unavailable(*a, **k):
            raise NotImplementedError("pytorch_msssim is stubbed in tests/ref_shims.py")

        _module("pytorch_msssim", ms_ssim=_unavailable, ssim=_unavailable, MS_SSIM=object, SSIM=object)

    if _missing("timm"):
        class DropPath(nn.Module):
            def __init__(self, drop_prob=0.0, *a, **k):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):
                return x

        def to_2tuple(x):
            return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

        layers = dict(DropPath=DropPath, to_2tuple=to_2tuple, trunc_normal_=nn.init.trunc_normal_)
        _module("timm")
        _module("timm.models")
        _module("timm.models.layers", **layers)
        _module("timm.layers", **layers)

    if _missing("lpips"):
        class LPIPS(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()

            def forward(self, *a, **k):
                raise NotImplementedError("lpips is stubbed in tests/ref_shims.py")

        _module("lpips", LPIPS=LPIPS)

    if _missing("pytorch_lightning"):
        class LightningModule(nn.Module):
            def log(self, *a, **k):
                pass

            def log_dict(self, *a, **k):
                pass

        _module("pytorch_lightning", LightningModule=LightningModule)
        _module("pytorch_lightning.utilities")
        _module("pytorch_lightning.utilities.distributed", rank_zero_only=lambda f: f)

    if _missing("addict"):
        class Dict(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)

            def __setattr__(self, k, v):
                self[k] = v

        _module("addict", Dict=Dict)

    if _missing("omegaconf"):
        _module("omegaconf", OmegaConf=object, DictConfig=dict, ListConfig=list)

    if _missing("skimage"):
        _module("skimage")
        _module("skimage.metrics", peak_signal_noise_ratio=None, structural_similarity=None)

    if _missing("wandb"):
        _module("wandb")


def import_reference():
    """Put /root/reference on sys.path with every missing dependency stubbed and compressai shimmed."""
    import dc_vic_b200 as dcv
    install_third_party_stubs()
    dcv.install_compressai_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
