"""Host-side profile of the transition-model training step (cProfile over 30 steps, device kept busy)."""
import os, sys, time, cProfile, pstats, torch
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

def step():
    p1, v1, n1 = net(pos, vel, box, box_n)
    loss = (torch.exp(-n1 / 40.0) * ((p1 - gt) ** 2).sum(-1)).mean()
    for p_ in net.parameters():
        p_.grad = None
    loss.backward()

for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(30):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / 30:.3f} ms/step, with drain {1e3 * (t2 - t0) / 30:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(30):
    step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
