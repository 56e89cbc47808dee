"""In-tree build of the C-ABI CUDA library (``madrl_b200/libmadrl_b200.so``) for sm_100a.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmadrl_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def source_hash():
    """sha256 over every source / header the library is built from + the nvcc flags.  A content
    hash, not mtimes: the tree is copied to the GPU box and git checkouts do not keep mtimes."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    deps = sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(
        glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def kernel_hash(family):
    """sha256 over the files one env kernel is compiled from (its .cu + the shared headers) + the nvcc
    flags: what a committed ncu capture of that kernel is valid for (profiles/r2_traffic.json)."""
    import hashlib
    src = {"ww": "waterworld.cu", "pe": "pursuit.cu", "hw": "hostage.cu"}[family]
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for name in (src, "common.cuh", "philox.cuh", "host_pipeline.cuh"):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    return h.hexdigest()


def _stale():
    """True if the library is missing or was built from other sources than the ones in the tree."""
    if not os.path.exists(LIB) or not os.path.exists(LIB + ".srchash"):
        return True
    with open(LIB + ".srchash") as f:
        return f.read().strip() != source_hash()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ (in parallel) and link one shared library.  Returns its path."""
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + cflags + ["-c", "-o", obj, src]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return obj, cmd, res

    with ThreadPoolExecutor(max_workers=8) as pool:
        results = list(pool.map(compile_one, sources()))
    for obj, cmd, res in results:
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed (%d): %s" % (res.returncode, " ".join(cmd)))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [r[0] for r in results]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("nvcc link failed (%d): %s" % (res.returncode, " ".join(cmd)))
    with open(LIB + ".srchash", "w") as f:
        f.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
