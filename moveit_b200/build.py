"""Builds libb200df.so (the C-ABI library, include/b200df.h) in-tree with nvcc for sm_100a.

The kernels that must round like the reference's non-FMA double arithmetic are compiled with --fmad=false.
cudart is linked statically so the library loads next to torch's own runtime without path games.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb200df.so")
# the same library with -DB200DF_BOUNDS_CHECK: every computed voxel / list index is tested on the device before use
# (b200df_internal.cuh); loaded through B200DF_LIB by the bounds-check test, never by the product
OUT_CHECKED = os.path.join(HERE, "libb200df_checked.so")
SOURCES = ["api.cu", "field.cu", "edt.cu", "edt_tma.cu", "validity.cu", "gradients.cu", "validity_fast.cu", "validity_small.cu", "motion.cu", "multi.cu"]
FMA_OK = {"validity_fast.cu"}  # fp32 screening kernel: only has to stay inside its error margins
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def _flags(src):
    f = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]
    if src not in FMA_OK:
        f.append("--fmad=false")
    return f


def needs_build(out=None):
    out = out or OUT
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b200df.h"),
                                                                 os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, checked=False):
    out = OUT_CHECKED if checked else OUT
    if not force and not needs_build(out):
        return out
    nvcc = nvcc_path()
    objdir = os.path.join(HERE, "build", "checked") if checked else os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    extra = ["-DB200DF_BOUNDS_CHECK"] if checked else []
    for s in SOURCES:
        o = os.path.join(objdir, s.replace(".cu", ".o"))
        cmd = [nvcc] + ARCH + _flags(s) + extra + (["-Xptxas", "-v"] if verbose else []) + \
            ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for cmd, p in procs:
        log, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(log)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [nvcc] + ARCH + ["-shared", "-cudart", "static", "-o", out] + objs
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--checked" in sys.argv:
        print(build(force="--force" in sys.argv, checked=True))
