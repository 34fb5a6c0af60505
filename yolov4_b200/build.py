"""Compiles csrc/*.cu into yolov4_b200/libyolohead.so for sm_100a (nvcc cross-compiles without a GPU).

The .so is built in-tree (git-ignored, but it travels to the GPU box with the repo snapshot).
"""
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libyolohead.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                      # no FMA contraction: the reference rounds a*b and +c separately
    "-Xcompiler", "-fPIC",
]


def _nvcc():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; libyolohead.so cannot be built")
    return p


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False, extra=(), out=None):
    """extra: additional nvcc flags (e.g. -DYL_WT_STAGES=3 for tuning builds); out: alternative output path."""
    if out is None and not force and not is_stale():
        return LIB
    objs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        if out is not None:
            obj = obj[:-2] + "." + os.path.basename(out) + ".o"
        cmd = [_nvcc()] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        log, _ = p.communicate()
        if verbose and log:
            print(log)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), log))
    cmd = [_nvcc(), "-shared", "-o", out or LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    subprocess.check_call(cmd)
    return out or LIB


if __name__ == "__main__":
    import sys
    print(build_lib(force=True, verbose="-v" in sys.argv))
