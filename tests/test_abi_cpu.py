"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol
include/dcvic_b200.h declares; the product package never touches oracle/ and refuses CPU tensors."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "dcvic_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dcvic_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from dc_vic_b200 import build, _lib
    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from dc_vic_b200 import _lib
    names = header_functions()
    assert len(names) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signature table out of sync with the header"


def test_host_only_entry_points(lib):
    assert b"sm_100a" in lib.dcvic_version()
    assert lib.dcvic_error_string(-2).decode().startswith("shape not supported")
    assert lib.dcvic_vq_path(4, 256, 0) == 0
    assert lib.dcvic_vq_path(256, 1024, 2) == 1          # FORCE_EXACT
    assert lib.dcvic_vq_path(4096, 16, 0) == -2
    assert lib.dcvic_vq_workspace_bytes(64, 256, 32, 32, 1024) > 0
    assert lib.dcvic_vq_workspace_bytes(0, 256, 32, 32, 1024) == 0
    assert lib.dcvic_gc_workspace_bytes(64, 320 * 32 * 32) > 0
    # argument checks fire before anything is enqueued (safe without a GPU)
    assert lib.dcvic_vq_forward(None, None, 1, 4, 2, 2, 8, 0.25, 1, None, None, None, None, None, 0, None, 0, None) == -1
    assert lib.dcvic_gc_forward(None, None, None, None, 1, 4, 4, 4, 4, 0.11, 1e-9, 0, None, None, None, None, 0, None) == -1
    assert lib.dcvic_rate_bits(None, 1, 4, None, None, 0, None) == -1


def test_pmf_to_quantized_cdf_matches_oracle():
    import numpy as np
    from dc_vic_b200 import pmf_to_quantized_cdf
    from oracle import entropy_oracle as O
    rng = np.random.default_rng(1)
    for n in (1, 2, 7, 64, 513):
        p = rng.random(n).astype(np.float32) ** 6
        p /= p.sum()
        assert pmf_to_quantized_cdf(p).tolist() == O.pmf_to_quantized_cdf(p).tolist()
    assert pmf_to_quantized_cdf([0.5, 0.0, 0.5]).tolist() == [0, 32767, 32768, 65536]
    with pytest.raises(RuntimeError):
        pmf_to_quantized_cdf([float("nan"), 0.5])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dc_vic_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "/root/reference" not in text, f


def test_cpu_tensors_are_refused():
    import dc_vic_b200 as d
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d.VectorQuantizer2(8, 4, 0.25)(torch.zeros(1, 4, 2, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d.SteGaussianMeanScaleConditional(scale_bound=0.11)(torch.zeros(1, 2, 2, 2), torch.ones(1, 4, 2, 2), False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d.SteEntropyBottleneck(channels=2)(torch.zeros(1, 2, 2, 2), False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d.likelihood_to_bit(torch.ones(4), 1)


def test_state_dict_names_match_reference_layout():
    import dc_vic_b200 as d
    from oracle import entropy_oracle as O
    eb, ob = d.SteEntropyBottleneck(channels=3), O.SteEntropyBottleneck(channels=3)
    assert {k: tuple(v.shape) for k, v in eb.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in ob.state_dict().items()}
    gc, og = d.SteGaussianMeanScaleConditional(scale_bound=0.11), O.SteGaussianMeanScaleConditional(scale_bound=0.11)
    assert sorted(gc.state_dict()) == sorted(og.state_dict())
    assert list(d.VectorQuantizer2(16, 4, 0.25).state_dict()) == ["embedding.weight"]
    with pytest.raises(TypeError):
        d.GaussianMeanScaleConditional(scale_bound=None)


def test_compressai_shim_and_registry():
    import sys
    import dc_vic_b200 as d
    saved = {k: v for k, v in sys.modules.items() if k == "compressai" or k.startswith("compressai.")}
    try:
        d.install_compressai_shim(force=True)
        from compressai.entropy_models import EntropyBottleneck, GaussianConditional
        from compressai.models.utils import update_registered_buffers
        from compressai.ops import LowerBound  # noqa: F401
        assert EntropyBottleneck is d.EntropyBottleneck and GaussianConditional is d.GaussianConditional

        class Reg:
            _obj_map = {}
        d.register_entropy_models(Reg)
        assert Reg._obj_map["SteEntropyBottleneck"]["obj"] is d.SteEntropyBottleneck
        m = d.SteEntropyBottleneck(channels=2)
        sd = {"entropy_model_z._quantized_cdf": torch.zeros(2, 9, dtype=torch.int32)}
        update_registered_buffers(m, "entropy_model_z", ["_quantized_cdf"], sd)
        assert tuple(m._quantized_cdf.shape) == (2, 9)
    finally:
        for k in [k for k in sys.modules if k == "compressai" or k.startswith("compressai.")]:
            del sys.modules[k]
        sys.modules.update(saved)
