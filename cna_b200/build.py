"""Build ``libcna_b200.so`` (the C-ABI of include/cna_b200.h) in-tree with nvcc for sm_100a.

    python -m cna_b200.build [--force] [--verbose]

The library has no Python or torch dependency: it is plain CUDA C++ behind ``extern "C"``.  The
object files and the shared library land in ``cna_b200/_build/`` (git-ignored, but shipped to the
GPU box by gpurun).  A content hash of the sources is stored next to the library so that rebuilds
happen only when something changed.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libcna_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (need CUDA 12.9 for sm_100a)")


# host-only translation units (perm_host.cpp restates numpy's RNG arithmetic literally: baseline
# x86-64 code generation, no FMA contraction, so that every double matches numpy's bit for bit)
GXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-ffp-contract=off", "-pthread"]


def sources():
    return sorted(f for f in os.listdir(SRC) if f.endswith((".cu", ".cpp")))


def _digest():
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS + GXX_FLAGS).encode())
    files = [os.path.join(SRC, f) for f in sorted(os.listdir(SRC))] + \
            [os.path.join(INCLUDE, f) for f in sorted(os.listdir(INCLUDE))]
    for path in files:
        h.update(path.encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ and link the shared library.  Returns its path."""
    os.makedirs(OUT, exist_ok=True)
    stamp = os.path.join(OUT, "sources.sha256")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    nvcc = _nvcc()

    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc))), "include")

    def compile_one(name):
        obj = os.path.join(OUT, os.path.splitext(name)[0] + ".o")
        if name.endswith(".cpp"):
            cmd = [os.environ.get("CXX", "g++"), *GXX_FLAGS, "-I", INCLUDE, "-I", cuda_inc, "-c",
                   os.path.join(SRC, name), "-o", obj]
        else:
            cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(SRC, name), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj[:-2] + ".ptxas.log", "w") as fh:
            fh.write(res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {name}:\n{res.stderr}\n{res.stdout}")
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-lcudart_static", "-lcuda"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stderr}\n{res.stdout}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
