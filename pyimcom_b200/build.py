"""Build recipe of libpyimcom_b200.so (the product's CUDA library, sm_100a only).

``python -m pyimcom_b200.build`` compiles pyimcom_b200/csrc/*.cu with nvcc (cross-compiles without a GPU)
into pyimcom_b200/lib/libpyimcom_b200.so.  The library is built IN-TREE so that it travels to the GPU box;
it is git-ignored.  There is no CPU fallback: pyimcom_b200._lib raises if the library is missing.
"""

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libpyimcom_b200.so")
SOURCES = ["capi.cu", "interp.cu", "linalg.cu", "eigen.cu", "kappa.cu", "iter.cu", "coadd.cu", "partition.cu", "ozaki.cu", "trieig.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, tag: str = "", extra_flags=()) -> str:
    """tag / extra_flags: an A/B build beside the product library (lib/libpyimcom_b200_<tag>.so, loaded with
    B200_LIB=...), e.g. ``build(tag="ozcap", extra_flags=["-DB200_OZ_LB=512"])``."""
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj" + ("_" + tag if tag else ""))
    LIB = os.path.join(LIBDIR, f"libpyimcom_b200{'_' + tag if tag else ''}.so")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "pyimcom_b200.h"))
    nvcc = os.environ.get("NVCC", "nvcc")

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src[:-3] + ".o")
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError(f"nvcc failed on {src}")
        return o

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    _tag = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--tag=")), "")
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, tag=_tag,
                extra_flags=[a for a in sys.argv[1:] if a.startswith("-D")]))
