"""Build libldit_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libldit_b200.so")
SOURCES = ["ldit_api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ldit.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str | None = None) -> str:
    """Product build by default.  ``extra_flags``: e.g. ``["-DLDIT_EXPERIMENTAL"]`` (superseded attention variants and the
    fused fc1+fc2 kernel), ``["-DLDIT_DEBUG_HOOKS"]`` (GEMM timeline / LDIT_GEMM_DBG), ``["-DLDIT_A3_TIMELINE"]``;
    also read from the environment variable LDIT_BUILD_FLAGS.  ``out``: write a variant somewhere else (A/B builds)."""
    extra_flags = list(extra_flags) + os.environ.get("LDIT_BUILD_FLAGS", "").split()
    target = out or LIB
    if not force and not extra_flags and out is None and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, "-o", target] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return target


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
