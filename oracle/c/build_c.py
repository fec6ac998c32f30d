"""Compile the oracle's C port (gcc, OpenMP).  Oracle/CPU-baseline infrastructure, not product."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "kalman_em.c")
LIB = os.path.join(HERE, "liboracle_kem.so")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    subprocess.run(["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", SRC, "-o", LIB, "-lm"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
