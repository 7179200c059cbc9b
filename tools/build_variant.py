#!/usr/bin/env python
"""Builds a copy of libb200df.so with extra -D flags on selected sources, for kernel tuning sweeps on the GPU box:
    python tools/build_variant.py NAME edt_tma.cu -DB200DF_EDT_MINB=16 -DB200DF_EDT_SB=3
writes variants/libb200df_NAME.so (loaded with B200DF_LIB=variants/libb200df_NAME.so).  The other objects are the ones
moveit_b200/build.py left in moveit_b200/build/."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moveit_b200 import build as b  # noqa: E402


def main():
    name = sys.argv[1]
    srcs = [a for a in sys.argv[2:] if a.endswith(".cu")]
    defs = [a for a in sys.argv[2:] if not a.endswith(".cu")]
    b.build()
    out_dir = os.path.join(ROOT, "variants")
    os.makedirs(out_dir, exist_ok=True)
    objs = []
    for s in b.SOURCES:
        o = os.path.join(b.HERE, "build", s.replace(".cu", ".o"))
        if s in srcs:
            o = os.path.join(out_dir, "%s_%s.o" % (name, s.replace(".cu", "")))
            subprocess.check_call([b.nvcc_path()] + b.ARCH + b._flags(s) + defs + ["-c", os.path.join(b.CSRC, s), "-o", o])
        objs.append(o)
    out = os.path.join(out_dir, "libb200df_%s.so" % name)
    subprocess.check_call([b.nvcc_path()] + b.ARCH + ["-shared", "-cudart", "static", "-o", out] + objs)
    print(out)


if __name__ == "__main__":
    main()
