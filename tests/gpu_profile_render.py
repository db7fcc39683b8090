"""Profiling workload (not a pytest): N forwards of BASELINE config[1] (800x800, 64+128, 19,683 particles) and,
optionally, a few transition steps; prints per-chunk row counts so an ncu capture of launch i can be matched to
its rows.   python tests/gpu_profile_render.py [n_forwards] [n_transition_steps]     (NF_PROFILE_LATTICE=37: config[4]'s 50,653 particles)"""
import json, os, sys, torch
torch.set_grad_enabled(False)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import neurofluid_b200 as nb
from neurofluid_b200 import scenes
dev = torch.device("cuda:0")
nfwd = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ntr = int(sys.argv[2]) if len(sys.argv) > 2 else 0
H = 800
rays, focal, cw = scenes.camera_rays(H, H)
particles = torch.from_numpy(scenes.lattice_particles(int(os.environ.get("NF_PROFILE_LATTICE", "27")), 0))
net = nb.RenderNet(scenes.render_cfg(), scenes.NEAR, scenes.FAR); net.load_state_dict(scenes.init_render_state(0, 5.0)); net = net.to(dev)
rays_d, p_d, ro = rays.to(dev), particles.to(dev), cw[:, 3].to(dev)
for it in range(nfwd):
    out = net(p_d, ro, rays_d, focal, cw)
torch.cuda.synchronize()
st = net.last_stats.cpu().tolist()
info = {"chunks": [{"rays": min(net.max_rays_per_launch, H * H - i * net.max_rays_per_launch), "rows_coarse": s[0], "rows_fine": s[1]}
                   for i, s in enumerate(st)], "launches_per_forward": {"k_stage_q0": len(st), "k_nerf_mlp": 2 * len(st),
                                                                         "k_stage_mid": len(st), "k_stage_fin": len(st)}}
if ntr:
    n = 31
    half = (n - 1) / 2 * 0.05
    pos = torch.from_numpy(scenes.lattice_particles(n, 0, center=(0.0, 0.0, -1 + 0.03 + half))).to(dev)
    vel = torch.zeros_like(pos)
    bp, bn = scenes.box_points(0.032)
    box, box_n = torch.from_numpy(bp).to(dev), torch.from_numpy(bn).to(dev)
    tn = nb.ParticleNet(gravity=(0.0, 0.0, -9.81)); tn.load_state_dict(scenes.init_particle_state(0)); tn = tn.to(dev)
    for it in range(ntr):
        pos, vel, _ = tn(pos, vel, box, box_n)
    torch.cuda.synchronize()
    info["transition"] = {"n_particles": n ** 3, "n_box": int(box.shape[0]), "steps": ntr}
print(json.dumps(info))
