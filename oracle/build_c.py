"""Builds oracle/_ref-free C restatement: oracle/libtzk_oracle.so (gcc -O3 -fopenmp).  Test infrastructure."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "tzk_oracle.c")
LIB = os.path.join(HERE, "libtzk_oracle.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"],
                       check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
