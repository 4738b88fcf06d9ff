"""In-tree build of libdcvic_b200.so (hand-written CUDA for sm_100a, C ABI in include/dcvic_b200.h).

    python -m dc_vic_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
repo snapshot; nothing is installed into site-packages.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libdcvic_b200.so")
SOURCES = ["abi.cu", "vq_simt.cu", "vq_tcgen05.cu", "gaussian_conditional.cu", "entropy_bottleneck.cu", "codec_tables.cu", "token_decode.cu", "vq_finish_tma.cu", "vq_fused.cu", "rans.cu", "tiling.cu"]
HEADERS = ["common.cuh", "vq_common.cuh", "sm100_ptx.cuh", "../../include/dcvic_b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--use_fast_math=false"]


def _nvcc() -> str:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > built for d in deps)


def build(force: bool = False, verbose: bool = False, trace: bool = False, variant: str = "", defs=()) -> str:
    """trace=True builds lib/libdcvic_b200_trace.so with -DDCVIC_TRACE (role timing in the tcgen05 kernel;
    a measurement aid for tools/trace_run.py, never loaded by the package itself).  variant="x" with defs=("-DA=1",)
    builds lib/libdcvic_b200_x.so for A/B experiments (selected with DCVIC_B200_LIB, never loaded by default)."""
    out = LIB_PATH.replace(".so", "_trace.so") if trace else LIB_PATH
    if variant:
        out = LIB_PATH.replace(".so", f"_{variant}.so")
    if not trace and not variant and not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + (["-DDCVIC_TRACE"] if trace else []) + list(defs)
    cmd = [_nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + srcs
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building libdcvic_b200.so")
    if verbose:
        sys.stderr.write(proc.stderr)
    return out


if __name__ == "__main__":
    _variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv, variant=_variant,
                defs=[a for a in sys.argv if a.startswith("-D")]))
