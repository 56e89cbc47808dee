"""Build an experiment variant of the library: scripts/build_variant.py NAME -DFLAG=1 ...
-> madrl_b200/variants/libmadrl_b200_NAME.so (git-ignored, ships to the GPU box).  Select it at run
time with MADRL_B200_LIB=<path>."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madrl_b200 import build as B  # noqa: E402


def main():
    name, defs = sys.argv[1], sys.argv[2:]
    outdir = os.path.join(B.HERE, "variants")
    objdir = os.path.join(outdir, "obj_" + name)
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in B.NVCC_FLAGS if f != "-shared"] + defs

    def one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        res = subprocess.run(["nvcc"] + cflags + ["-c", "-o", obj, src], stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True)
        if res.returncode:
            sys.stderr.write(res.stdout)
            raise SystemExit("nvcc failed on " + src)
        return obj

    with ThreadPoolExecutor(max_workers=8) as pool:
        objs = list(pool.map(one, B.sources()))
    lib = os.path.join(outdir, "libmadrl_b200_%s.so" % name)
    subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs)
    print(lib)


if __name__ == "__main__":
    main()
