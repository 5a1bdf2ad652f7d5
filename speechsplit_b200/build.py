"""In-tree build of libssfe.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m speechsplit_b200.build [--force]

The shared library lands next to this file's package (speechsplit_b200/libssfe.so) so that it
travels to the GPU box with the repo snapshot; it is git-ignored.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "libssfe.so")
OBJ = os.path.join(PKG, "_obj")
CU = ["api.cu", "stft_mel.cu", "filtfilt.cu", "mt19937.cu", "f0_post.cu", "rapt.cu", "interp.cu"]
CPP = ["filt_consts.cpp", "mt_jump.cpp"]
# RAPT reproduces the original's float evaluation order; fused multiply-adds would change it
EXTRA = {"rapt.cu": ["--fmad=false"]}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(s) > t for s in (src,) + tuple(extra))


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    headers = tuple(os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".cuh", ".h")))
    headers += (os.path.join(PKG, "..", "include", "ssfe.h"),)
    objs = []
    logs = []
    for f in CU:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ, f + ".o")
        objs.append(obj)
        if force or _newer(src, obj, headers):
            r = subprocess.run([nvcc] + NVCC_FLAGS + EXTRA.get(f, []) + ["-c", src, "-o", obj],
                               capture_output=True, text=True)
            logs.append((f, r.stderr))
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed on %s" % f)
            with open(os.path.join(OBJ, f + ".ptxas.log"), "w") as fh:
                fh.write(r.stderr)
            if verbose:
                sys.stderr.write(r.stderr)
    for f in CPP:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ, f + ".o")
        objs.append(obj)
        if force or _newer(src, obj):
            subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-c", src, "-o", obj])
    if force or any(_newer(o, OUT) for o in objs):
        subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs
                              + ["-lcudart", "-lquadmath"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
