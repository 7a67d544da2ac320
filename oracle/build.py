"""Build oracle/liboracle.so (the CPU restatement).  TEST INFRASTRUCTURE ONLY."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pointnet2_oracle.c")
SRCS = [SRC, os.path.join(HERE, "mc_oracle.c")]
DEPS = SRCS + [os.path.join(HERE, "mc_tables_oracle.h")]
LIB = os.path.join(HERE, "liboracle.so")


def build(force=False):
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= max(os.path.getmtime(f) for f in DEPS)):
        return LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
           "-o", LIB] + SRCS + ["-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
