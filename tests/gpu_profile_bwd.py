"""Profile driver: three renderer training steps (forward + backward) on 1024 random rays of config[1]'s scene."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import neurofluid_b200 as nb
from neurofluid_b200 import scenes
dev = torch.device("cuda:0")
H = 800
rays, focal, cw = scenes.camera_rays(H, H)
g = torch.Generator().manual_seed(0)
crop = scenes.center_crop_rays(rays, H, H, 200)
sel = torch.randperm(crop.shape[0], generator=g)[:1024]
r = crop[sel].contiguous().to(dev)
particles = torch.from_numpy(scenes.lattice_particles(27, 0)).to(dev).requires_grad_(True)
net = nb.RenderNet(scenes.render_cfg(), scenes.NEAR, scenes.FAR); net.load_state_dict(scenes.init_render_state(0, 5.0)); net = net.to(dev)
target = torch.rand(1024, 3, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for it in range(3):
    ev[0].record()
    out = net(particles, cw[:, 3].to(dev), r, focal, cw)
    loss = ((out["rgb0"] - target) ** 2).mean() + ((out["rgb1"] - target) ** 2).mean()
    ev[1].record()
    loss.backward()
    ev[2].record()
    torch.cuda.synchronize()
    rows = net.last_stats.sum(0).tolist()
    print(f"step {it}: forward {ev[0].elapsed_time(ev[1]):.3f} ms, backward {ev[1].elapsed_time(ev[2]):.3f} ms, rows coarse/fine {rows[0]}/{rows[1]}")
