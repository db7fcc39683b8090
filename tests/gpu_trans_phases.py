import os, sys, time, torch, ctypes as C
torch.set_grad_enabled(False)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import neurofluid_b200 as nb
from neurofluid_b200 import scenes, _lib
dev = torch.device("cuda:0")
n = 31
half = (n - 1) / 2 * 0.05
pos = torch.from_numpy(scenes.lattice_particles(n, 0, center=(0.0, 0.0, -1 + 0.03 + half))).to(dev)
vel = torch.zeros_like(pos)
bp, bn = scenes.box_points(0.032)
box, box_n = torch.from_numpy(bp).to(dev), torch.from_numpy(bn).to(dev)
net = nb.ParticleNet(gravity=(0.0, 0.0, -9.81)); net.load_state_dict(scenes.init_particle_state(0)); net = net.to(dev)
net(pos, vel, box, box_n)
p, v, b, bf, outs, ws = net._prepare(pos, vel, box, box_n, None)
N = pos.shape[0]
for rep in range(2):
    ts = []
    for ph in range(5):
        a = net._args(p, v, b, bf, outs, ws, phase=ph, shard=(0, N))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _lib.check(_lib.lib().nf_transition_step(C.byref(a), _lib.stream_ptr()), "step"); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print("phase ms:", [round(t, 3) for t in ts], "sum", round(sum(ts), 3))
# whole step (phase -1: cell-ordered kernels), CUDA events over 50 steps
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pp, vv = pos, vel
    e0.record()
    for _ in range(50):
        pp, vv, _ = net(pp, vv, box, box_n)
    e1.record(); torch.cuda.synchronize()
    print("whole step ms:", round(e0.elapsed_time(e1) / 50, 4))
