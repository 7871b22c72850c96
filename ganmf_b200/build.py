"""Builds libganmf_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

    python -m ganmf_b200.build          # or ganmf_b200.build.build()

The shared object lands next to this file so it travels with the repository snapshot to the
GPU box; it is git-ignored (build artefact)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libganmf_b200.so")
SOURCES = ["capi.cu"]
HEADERS = ["ptx.cuh", "tc_gemm.cuh", "kernels.cuh", "csr_kernels.cuh", "eval_kernels.cuh", "score_select.cuh",
           "gen_gemm.cuh"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "ganmf_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v" if verbose else "-O3",
           "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libganmf_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
