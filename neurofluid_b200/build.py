"""In-tree build of libnf_b200.so (hand-written sm_100a kernels + C ABI, include/nf_b200.h).

    python -m neurofluid_b200.build [--force] [--verbose] [--tuning]

`--tuning` builds libnf_b200_tune.so with -DNF_TUNING: the same kernels plus the NF_* environment overrides the
profiling / tuning scripts use (tests/gpu_tune.py, tests/gpu_mlp_trace.py; load it with NF_B200_LIB=<path>).  The
release library never reads the environment.

nvcc cross-compiles for sm_100a without a GPU; the resulting .so is git-ignored but travels to the
GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libnf_b200.so")
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _deps():
    return glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(os.path.dirname(HERE), "include", "nf_b200.h")]


def build(force: bool = False, verbose: bool = False, tuning: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    newest = max(os.path.getmtime(p) for p in _deps())
    out = OUT[:-3] + "_tune.so" if tuning else OUT
    objdir = OBJ + ("_tune" if tuning else "")
    flags = FLAGS + (["-DNF_TUNING"] if tuning else [])
    if not force and os.path.exists(out) and os.path.getmtime(out) >= newest:
        return out
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        r = subprocess.run([NVCC, *flags, "-c", src, "-o", obj], capture_output=True, text=True)
        return src, obj, r

    objs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for src, obj, r in ex.map(compile_one, srcs):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"--- {os.path.basename(src)}\n{r.stderr}\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
            with open(obj + ".ptxas.log", "w") as f:
                f.write(r.stderr)
            objs.append(obj)
    subprocess.check_call([NVCC, "-shared", "-o", out, *objs, "-lcudart", "-ldl"])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, tuning="--tuning" in sys.argv))
