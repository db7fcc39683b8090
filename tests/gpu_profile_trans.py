"""Profile driver: a few whole transition steps at BASELINE config[2] size (ncu -k regex:k_cconv_tc ...)."""
import os, sys, torch
torch.set_grad_enabled(False)
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
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    pos, vel, _ = net(pos, vel, box, box_n)
torch.cuda.synchronize()
print("ok")
