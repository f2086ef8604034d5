"""Build the CUDA shared library in-tree (sm_100a only).

``python -m contractn_b200.build`` or ``__graft_entry__.build()``.  nvcc
cross-compiles without a GPU; the resulting ``libctn_b200.so`` is git-ignored
but travels to the GPU box with the repository snapshot.

Every ``csrc/*.cu`` is compiled to its own object (in parallel, cached by a digest
of the file, the headers and the flags), then linked: touching one kernel file
recompiles that file only.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libctn_b200.so")
SOURCES = ["plan.cu", "kernels.cu", "gemm.cu", "gemm_tf32.cu", "gemm_tf32_v1.cu", "skinny.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _headers():
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    deps.append(os.path.join(ROOT, "include", "ctn_b200.h"))
    return deps


def _digest(paths, extra=""):
    """Content hash (mtimes do not survive a snapshot copy to the GPU box; contents do)."""
    h = hashlib.sha256()
    for d in paths:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(extra.encode())
    return h.hexdigest()


def _source_digest():
    return _digest([os.path.join(CSRC, s) for s in SOURCES] + _headers(), " ".join(NVCC_FLAGS + SOURCES))


def _stale():
    if not os.path.exists(LIB) or not os.path.exists(LIB + ".digest"):
        return True
    with open(LIB + ".digest") as f:
        return f.read().strip() != _source_digest()


def _compile_one(src, verbose):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, src[:-3] + ".o")
    dig = _digest([path] + _headers(), " ".join(NVCC_FLAGS))
    if os.path.exists(obj) and os.path.exists(obj + ".digest"):
        with open(obj + ".digest") as f:
            if f.read().strip() == dig:
                return obj, ""
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", path, "-o", obj + ".tmp.o"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n" + proc.stdout + proc.stderr)
    os.replace(obj + ".tmp.o", obj)
    with open(obj + ".digest", "w") as f:
        f.write(dig)
    return obj, proc.stderr


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu into contractn_b200/libctn_b200.so.  Returns its path."""
    if not force and not _stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            if f.endswith(".digest"):
                os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(lambda s: _compile_one(s, verbose), SOURCES))
    objs = [r[0] for r in results]
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB + ".tmp"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stdout + proc.stderr)
    os.replace(LIB + ".tmp", LIB)
    with open(LIB + ".digest", "w") as f:
        f.write(_source_digest())
    if verbose:
        for _, log in results:
            print(log)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
