"""Build recipe for the CPU oracle (test infrastructure; see oracle/__init__.py).

``python -m oracle.build`` compiles oracle/croutines.c -> oracle/liboracle_croutines.so with gcc.
-ffp-contract=off keeps one rounding per operation, as in the Numba/C reference statement order.
"""

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "croutines.c")
LIB = os.path.join(HERE, "liboracle_croutines.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
