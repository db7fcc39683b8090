"""Tuning driver (torchrun, >= 2 GPUs): the sharded transition step at config[2]'s size with the peer-memory exchange and with
ncclAllGather, in one process (the workspace is re-registered between the two), CUDA-event timed, max over ranks."""
import ctypes as C
import os
import sys

import torch
torch.set_grad_enabled(False)
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neurofluid_b200 as nb  # noqa: E402
from neurofluid_b200 import _lib, scenes, distributed  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 31
half = (n - 1) / 2 * 0.05
pos = torch.from_numpy(scenes.lattice_particles(n, 0, center=(0.0, 0.0, -1 + 0.03 + half))).to(dev)
vel = torch.zeros_like(pos)
bp, bn = scenes.box_points(0.032)
box, box_n = torch.from_numpy(bp).to(dev), torch.from_numpy(bn).to(dev)
tn = nb.ParticleNet(gravity=(0.0, 0.0, -9.81)); tn.load_state_dict(scenes.init_particle_state(0)); tn = tn.to(dev)


def run(no_peer, no_graph=False, steps=100):
    os.environ["NF_B200_NO_PEER"] = "1" if no_peer else "0"
    os.environ["NF_B200_NO_GRAPH"] = "1" if no_graph else "0"
    distributed._comm_ready.pop("registered", None)
    for _ in range(5):
        p, v, _n = distributed.transition_step_sharded(tn, pos, vel, box, box_n)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):          # the same state every step (a long free-running rollout of random weights clumps)
        p, v, _n = distributed.transition_step_sharded(tn, pos, vel, box, box_n)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    cnt = C.c_uint(0)
    _lib.check(_lib.lib().nf_comm_exchange_timeouts(C.byref(cnt)), "timeouts")
    return float(ms), distributed._comm_ready.get("peer_memory"), cnt.value, p


ms_ne, _, _, p_ne = run(True, True)
ms_pe, _, _, p_pe = run(False, True)
ms_ng, _, _, p_ng = run(True, False)
ms_pg, pm, to, p_pg = run(False, False)
ms_ng2, _, _, _ = run(True, False)
same = torch.equal(p_ne, p_pe) and torch.equal(p_ne, p_ng) and torch.equal(p_ne, p_pg)
if rank == 0:
    print(f"world {world}, ms per step: launched from the host: ncclAllGather {ms_ne:.4f}, peer-memory exchange {ms_pe:.4f}; "
          f"CUDA graph replay: ncclAllGather {ms_ng:.4f} (again {ms_ng2:.4f}), peer-memory exchange {ms_pg:.4f} "
          f"(active {pm}, wait time-outs {to}); identical results {same}")
dist.destroy_process_group()
