"""Compile oracle/maze_oracle.c into oracle/libmaze_oracle.so with gcc (no reference sources involved)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "maze_oracle.c")
LIB = os.path.join(HERE, "libmaze_oracle.so")


def build_oracle(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-std=c11", "-Wall", "-o", tmp, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build_oracle(force=True))
