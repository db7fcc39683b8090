"""Multi-GPU correctness check (run under torchrun on >= 2 GPUs): the sharded renderer and the sharded
transition step must reproduce the single-GPU results.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
"""
import os
import sys

import torch
torch.set_grad_enabled(False)
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neurofluid_b200 as nb  # noqa: E402
from neurofluid_b200 import scenes  # noqa: E402
from neurofluid_b200.distributed import render_image_sharded, transition_step_sharded  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)

H = 101
rays, focal, cw = scenes.camera_rays(H, H)
particles = torch.from_numpy(scenes.lattice_particles(12, 0)).to(dev)
net = nb.RenderNet(scenes.render_cfg(), scenes.NEAR, scenes.FAR)
net.load_state_dict(scenes.init_render_state(0, 5.0))
net = net.to(dev)
ro = cw[:, 3].to(dev)
single = net(particles, ro, rays.to(dev), focal, cw)["rgb1"].view(H, H, 3)
sharded = render_image_sharded(net, particles, ro, rays.view(H, H, 6).to(dev), focal, cw)
ok_r = torch.equal(single, sharded)

tn = nb.ParticleNet(gravity=(0.0, 0.0, -9.81))
tn.load_state_dict(scenes.init_particle_state(0))
tn = tn.to(dev)
pos = torch.from_numpy(scenes.lattice_particles(13, 1, center=(0.0, 0.0, -0.65))).to(dev)
vel = torch.zeros_like(pos)
bp, bn = scenes.box_points(0.06)
box, box_n = torch.from_numpy(bp).to(dev), torch.from_numpy(bn).to(dev)
p1, v1, n1 = tn(pos, vel, box, box_n)
p1, v1, n1 = p1.clone(), v1.clone(), n1.clone()
p2, v2, n2 = transition_step_sharded(tn, pos, vel, box, box_n)
ok_t = torch.equal(p1, p2) and torch.equal(v1, v2) and torch.equal(n1, n2)
# the same step through the other exchange / launch variants: ncclAllGather instead of the peer-memory exchange, CUDA graph replay
from neurofluid_b200 import distributed as nbd  # noqa: E402
for no_peer, graph in (("1", False), ("1", True), ("0", True), ("0", False)):
    os.environ["NF_B200_NO_PEER"] = no_peer
    nbd._comm_ready.pop("registered", None)          # re-register (collective: every rank does this here)
    for _ in range(2):                                # the second call of a graph variant is a pure replay
        p3, v3, n3 = transition_step_sharded(tn, pos, vel, box, box_n, graph=graph)
    ok_t = ok_t and torch.equal(p1, p3) and torch.equal(v1, v3) and torch.equal(n1, n3)
os.environ["NF_B200_NO_PEER"] = "0"
ok_t = ok_t and nbd.exchange_timeouts() == 0
# eval_e2e-shaped rollout: sharded rays (+ sharded or replicated transition) == one process doing everything
from neurofluid_b200 import ops, pipeline  # noqa: E402
cams = [(cw, focal)]
ok_e = True
p0, v0 = pos, vel
ref_img, ref_pos = [], []
for f in range(2):
    p0, v0, _ = tn(p0, v0, box, box_n)
    p0, v0 = p0.clone(), v0.clone()
    rr = ops.generate_rays(H, H, focal, cw.to(dev))
    ref_img.append(net(p0, ro, rr, focal, cw)["rgb1"].view(H, H, 3).clone())
    ref_pos.append(p0)
for mode in ("replicated", "sharded"):
    out = pipeline.rollout_and_render(tn, net, pos, vel, box, box_n, cams, H, H, 2, keep_images=True, transition=mode)
    for f in range(2):
        ok_e = ok_e and torch.equal(out["positions"][f], ref_pos[f]) and torch.equal(out["images"][f][0], ref_img[f])
res = torch.tensor([int(ok_r), int(ok_t), int(ok_e)], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"mgpu_check world={world}: renderer sharded==single {bool(res[0])}, transition sharded==single {bool(res[1])}, "
          f"rollout (sharded rays, replicated / sharded transition)==single {bool(res[2])}")
dist.destroy_process_group()
sys.exit(0 if bool(res.min()) else 1)
