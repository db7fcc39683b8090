"""Profile driver: three transition-model training steps (forward + backward) at config[2]'s size (29,791 particles)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import neurofluid_b200 as nb
from neurofluid_b200 import scenes
dev = torch.device("cuda:0")
n = 31
half = (n - 1) / 2 * 0.05
pos = torch.from_numpy(scenes.lattice_particles(n, 0, center=(0.0, 0.0, -1 + 0.03 + half))).to(dev)
vel = torch.zeros_like(pos)
bp, bn = scenes.box_points(0.032)
box, box_n = torch.from_numpy(bp).to(dev), torch.from_numpy(bn).to(dev)
net = nb.ParticleNet(gravity=(0.0, 0.0, -9.81)); net.load_state_dict(scenes.init_particle_state(0)); net = net.to(dev)
gt = pos + 0.001
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(3):
    ev[0].record()
    p1, v1, n1 = net(pos, vel, box, box_n)
    loss = (torch.exp(-n1 / 40.0) * ((p1 - gt) ** 2).sum(-1)).mean()
    ev[1].record()
    for p_ in net.parameters():
        p_.grad = None
    loss.backward()
    ev[2].record()
    torch.cuda.synchronize()
    print(f"step {it}: forward {ev[0].elapsed_time(ev[1]):.3f} ms, backward {ev[1].elapsed_time(ev[2]):.3f} ms")
