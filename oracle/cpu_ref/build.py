"""Builds the CPU baseline (restated reference path, see linmpc_admm.cpp) with g++ -O3 -fopenmp."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "linmpc_admm.cpp")
OUT = os.path.join(HERE, "libcpuref.so")


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) > os.path.getmtime(SRC):
        return OUT
    # -march=x86-64-v3 (AVX2/FMA) rather than -march=native: the .so is built in the CPU container and
    # travels to the GPU box, whose host CPU may differ
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-std=c++17", "-o", OUT, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building libcpuref.so")
    return OUT


if __name__ == "__main__":
    print(build(force=True))
