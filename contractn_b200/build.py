"""Build the CUDA shared library in-tree (sm_100a only).

``python -m contractn_b200.build`` or ``__graft_entry__.build()``.  nvcc
cross-compiles without a GPU; the resulting ``libctn_b200.so`` is git-ignored
but travels to the GPU box with the repository snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libctn_b200.so")
SOURCES = ["plan.cu", "kernels.cu", "gemm.cu", "gemm_tf32.cu", "skinny.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _source_digest():
    """Content hash of everything the library is built from (mtimes do not survive a
    snapshot copy to the GPU box; contents do)."""
    import hashlib

    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    deps.append(os.path.join(ROOT, "include", "ctn_b200.h"))
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + SOURCES).encode())
    return h.hexdigest()


def _stale():
    if not os.path.exists(LIB) or not os.path.exists(LIB + ".digest"):
        return True
    with open(LIB + ".digest") as f:
        return f.read().strip() != _source_digest()


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu into contractn_b200/libctn_b200.so.  Returns its path."""
    if not force and not _stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB + ".tmp"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    os.replace(LIB + ".tmp", LIB)
    with open(LIB + ".digest", "w") as f:
        f.write(_source_digest())
    if verbose:
        print(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
