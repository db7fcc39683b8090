"""Build recipe for the CPU oracle library (TEST INFRASTRUCTURE ONLY -- see oracle/csrc/nf_oracle.c).

    python oracle/build_oracle.py            # -> oracle/libnf_oracle.so

The reference (syguan96/NeuroFluid) is pure Python and ships no native sources, so there is no
`oracle/_ref` build: the compiled part of the oracle is our own C restatement of the third-party
operators the reference calls (PyTorch3D ball_query, Open3D ContinuousConv).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "nf_oracle.c")
OUT = os.path.join(HERE, "libnf_oracle.so")


def build(force: bool = False) -> str:
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
           "-fvisibility=hidden", "-o", OUT, SRC, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
