import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import neurofluid_b200 as nb
from neurofluid_b200 import scenes, _lib
import ctypes
dev = torch.device("cuda:0")
H = 800
rays, focal, cw = scenes.camera_rays(H, H)
particles = torch.from_numpy(scenes.lattice_particles(27, 0))
net = nb.RenderNet(scenes.render_cfg(), scenes.NEAR, scenes.FAR); net.load_state_dict(scenes.init_render_state(0, 5.0)); net = net.to(dev)
rays_d, p_d, ro = rays.to(dev), particles.to(dev), cw[:, 3].to(dev)
L = _lib.lib()
for occ in sys.argv[1:]:
    os.environ["NF_LOCKSTEP_MIN_OCC"] = occ
    for it in range(2):
        net(p_d, ro, rays_d, focal, cw)
    L.nf_profile_enable(1)
    torch.cuda.synchronize(); t0 = time.time()
    out = net(p_d, ro, rays_d, focal, cw)
    torch.cuda.synchronize(); dt = time.time() - t0
    ms = (ctypes.c_double * 5)(); n = ctypes.c_int(0); L.nf_profile_read(ms, ctypes.byref(n)); L.nf_profile_enable(0)
    st = net.last_stats.sum(0).tolist()
    print(f"min_occ={occ}: {dt*1e3:.1f} ms stages={[round(x,1) for x in ms]} lock_q={st[4]} rows_q={st[5]} "
          f"steps/group={64*st[6]/max(st[4],1):.1f} rebuilds/group={st[5]/max(st[4],1):.2f} cands/group={64*st[7]/max(st[4],1):.1f}", flush=True)
