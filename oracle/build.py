"""Compile the plain-C oracle (``oracle/csrc/oracle.c``) with gcc.  TEST INFRASTRUCTURE ONLY.

Output: ``oracle/_build/liboracle.so`` (git-ignored; it travels to the GPU box with the snapshot).
The reference itself is pure Python, so there is no ``oracle/_ref`` build of reference sources
(nothing to compile); the literal reference is executed only by ``oracle/make_golden.py`` in the
build container.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "liboracle.so")

CFLAGS = ["-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-std=c11", "-Wall"]


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", *CFLAGS, SRC, "-o", OUT, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
