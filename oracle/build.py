"""Build the oracle's C restatement (oracle/csrc/oracle.c -> oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY.  Called by __graft_entry__.build(), tests/conftest.py and bench.py's
cpu_baseline / --impl reference legs.  The reference (/root/reference) is pure Python, so there
is no compiled `oracle/_ref`; the Python reference is instead imported by oracle/make_golden.py in
the build container to produce the committed fixtures under tests/golden/.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "oracle.c")
LIB = os.path.join(HERE, "liboracle.so")


def build(force=False):
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
        return LIB
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-mfma",
           "-fno-fast-math", "-o", LIB, SRC, "-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
