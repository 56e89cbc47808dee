"""TEST INFRASTRUCTURE: compile the env kernels of madrl_b200/csrc with g++ against the fake
<cuda_runtime.h> of tests/emu/include into a CPU library with the product's C ABI
(tests/emu/_build/libmadrl_b200_emu[_<tag>].so).  Used only by tests/test_emulated_kernels.py."""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "madrl_b200", "csrc")
OUT = os.path.join(HERE, "_build")
SOURCES = ["common.cu", "waterworld.cu", "pursuit.cu", "hostage.cu", "postproc.cu", "heuristics.cu"]
# postproc.cu: its shared arrays are block-static, not `extern`; the fixed reduction grid (592 blocks of
# 256 threads on the GPU) is narrowed so that a launch does not create 150 000 fibers
PER_FILE_FLAGS = {"postproc.cu": ["-D__shared__=static", "-DMADRL_MOM_BLOCKS=6"]}
CXXFLAGS = ["-std=c++17", "-O1", "-g", "-fPIC", "-Wno-unknown-pragmas", "-Wno-attributes", "-x", "c++",
            "-I", os.path.join(HERE, "include")]


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out += [os.path.join(HERE, "include", "cuda_runtime.h"), os.path.join(HERE, "emu_runtime.cpp"),
            os.path.join(ROOT, "include", "madrl_b200.h"), os.path.abspath(__file__)]
    return out


def build(defines=(), force=False):
    """Returns the path of the emulator library built with the given -D flags."""
    tag = hashlib.sha1(" ".join(sorted(defines)).encode()).hexdigest()[:10] if defines else "default"
    lib = os.path.join(OUT, "libmadrl_b200_emu_%s.so" % tag)
    if not force and os.path.exists(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in _deps()):
        return lib
    objdir = os.path.join(OUT, "obj_" + tag)
    os.makedirs(objdir, exist_ok=True)
    cxx = os.environ.get("CXX", "g++")

    def one(src):
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        cmd = [cxx] + CXXFLAGS + list(defines) + PER_FILE_FLAGS.get(os.path.basename(src), []) + ["-c", "-o", obj, src]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode:
            sys.stderr.write(res.stdout)
            raise RuntimeError("g++ failed: " + " ".join(cmd))
        return obj

    srcs = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(HERE, "emu_runtime.cpp")]
    with ThreadPoolExecutor(max_workers=6) as pool:
        objs = list(pool.map(one, srcs))
    subprocess.check_call([cxx, "-shared", "-o", lib] + objs)
    return lib


if __name__ == "__main__":
    print(build([a for a in sys.argv[1:] if a.startswith("-D")], force="--force" in sys.argv))
