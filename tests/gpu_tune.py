import os, sys, time, torch
torch.set_grad_enabled(False)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import neurofluid_b200 as nb
from neurofluid_b200 import scenes, _lib
import ctypes
os.environ["NF_TUNE_LIVE"] = "1"
dev = torch.device("cuda:0")
H = 800
rays, focal, cw = scenes.camera_rays(H, H)
particles = torch.from_numpy(scenes.lattice_particles(int(os.environ.get('NF_TUNE_N', '27')), 0))
net = nb.RenderNet(scenes.render_cfg(), scenes.NEAR, scenes.FAR); net.load_state_dict(scenes.init_render_state(0, 5.0)); net = net.to(dev)
rays_d, p_d, ro = rays.to(dev), particles.to(dev), cw[:, 3].to(dev)
L = _lib.lib()
for occ in sys.argv[1:]:        # "stream:solo_max_occ,peel_lanes,peel_from" or "scs:sub_span"
    mode, arg = occ.split(":")
    net.search = {"stream": _lib.NF_SEARCH_STREAM, "scs": _lib.NF_SEARCH_SWEEP}[mode]
    if mode == "stream":
        a, b, c = arg.split(",")
        os.environ["NF_SOLO_MAX_OCC"], os.environ["NF_PEEL_LANES"], os.environ["NF_PEEL_FROM"] = a, b, c
    else:
        os.environ["NF_SUB_SPAN"], os.environ["NF_SUB_LOOK"] = arg.split(",")
    for it in range(2):
        net(p_d, ro, rays_d, focal, cw)
    L.nf_profile_enable(1)
    torch.cuda.synchronize(); t0 = time.time()
    out = net(p_d, ro, rays_d, focal, cw)
    torch.cuda.synchronize(); dt = time.time() - t0
    ms = (ctypes.c_double * 5)(); n = ctypes.c_int(0); L.nf_profile_read(ms, ctypes.byref(n)); L.nf_profile_enable(0)
    st = net.last_stats.sum(0).tolist()
    print(f"tune={occ}: {dt*1e3:.1f} ms stages={[round(x,1) for x in ms]} fine: groups={st[4]} solo={st[5]} "
          f"steps/group={64*st[6]/max(st[4],1):.1f} tests/group={64*st[7]/max(st[4],1):.1f} | coarse: groups={st[8]} solo={st[9]} "
          f"steps/group={64*st[10]/max(st[8],1):.1f} tests/group={64*st[11]/max(st[8],1):.1f} rgb1={out['rgb1'].double().sum().item():.6f}", flush=True)
