"""Build libhmvit_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["csrc/api.cu"]
HEADERS = ["csrc/common.cuh", "csrc/rowgemm.cuh", "csrc/attn.cuh", "csrc/attn_split.cuh", "csrc/attn_fused.cuh", "csrc/attn_fa2.cuh", "csrc/attn_bwd.cuh", "csrc/bwd.cuh", "csrc/wgrad_tc.cuh", "csrc/dgrad_cat.cuh", "csrc/dropout.cuh", "csrc/wmma_shared.cuh", "csrc/chain.cuh", "csrc/qkv.cuh", "csrc/decoder.cuh", "csrc/postproc.cuh", "csrc/pillar.cuh", "../include/hmvit_b200.h"]
OUTPUT = os.path.join(_HERE, "libhmvit_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC"]


def _stale() -> bool:
    if not os.path.exists(OUTPUT):
        return True
    t = os.path.getmtime(OUTPUT)
    return any(os.path.getmtime(os.path.join(_HERE, f)) > t for f in SOURCES + HEADERS)


def build_extension(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUTPUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUTPUT] + SOURCES
    res = subprocess.run(cmd, cwd=_HERE, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUTPUT
