"""Multi-GPU parity under the driver's `pytest -m gpu`: spawns tests/mgpu_check.py under torchrun on 2 GPUs of the box
(sharded renderer == single GPU, sharded transition == single GPU, sharded eval rollout == single process, all bit for
bit).  Skips when the box has one GPU; the log is kept under gpurun_out/ (copied to profiles/ for the record)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_paths_equal_single_gpu_bitwise():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs, box has {n}")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "mgpu_check.log"), "w") as f:
        f.write(r.stdout + "\n--- stderr ---\n" + r.stderr[-4000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "renderer sharded==single True" in r.stdout and "transition sharded==single True" in r.stdout
