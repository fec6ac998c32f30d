"""TEST HARNESS ONLY: compile the kernel source (dfm_api.cu + kernels) with g++ -DDFM_EMU so the
kernels' index/algebra logic can be exercised in the GPU-less build container.  The resulting
tests/emu/libdfm_emu.so is never loaded by the package, bench.py or smoke()."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "dynamic_factor_models_b200", "csrc", "dfm_api.cu")
LIB = os.path.join(HERE, "libdfm_emu.so")


def build(force=False):
    csrc = os.path.dirname(SRC)
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(ROOT, "include", "dfm_b200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DDFM_EMU", "-x", "c++", SRC, "-o", LIB,
                    "-Wno-unused-function"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
