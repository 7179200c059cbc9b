"""Compiles the C++ host mirror's driver (g++, C++17) against libb200df.so.  The mirror itself is header-only."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(HERE, "host_driver")


def build(force=False):
    src = os.path.join(HERE, "host_driver.cpp")
    hdr = os.path.join(HERE, "collision_env_b200.hpp")
    hdr2 = os.path.join(HERE, "distance_field_b200.hpp")
    hdr3 = os.path.join(HERE, "b200_config.hpp")
    lib = os.path.join(PKG, "libb200df.so")
    if not os.path.exists(lib):
        raise RuntimeError("build libb200df.so first (python -m moveit_b200.build)")
    deps = [src, hdr, hdr2, hdr3, lib, os.path.join(PKG, "..", "include", "b200df.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-o", OUT, src, "-L" + PKG, "-lb200df",
           "-lz", "-Wl,-rpath,$ORIGIN/..", "-pthread"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
