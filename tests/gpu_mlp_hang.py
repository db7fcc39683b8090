"""Tuning driver: one MLP launch of N tiles, with a 5 s watchdog (a new kernel variant that deadlocks must not hold the GPU box
until gpurun's limit), compared with the tuning build's one-tile kernel (NF_MLP_IMPL=1).  Usage: gpu_mlp_hang.py TILES"""
import os, sys, time, torch
torch.set_grad_enabled(False)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import neurofluid_b200 as nb
from neurofluid_b200 import scenes, ops
dev = torch.device("cuda:0")
net = nb.RenderNet(scenes.render_cfg(), 9., 13.); net.load_state_dict(scenes.init_render_state(0)); net = net.to(dev)
packed = ops.pack_nerf_weights(net.nerf_fine.ordered_params())
tiles = int(sys.argv[1])
n = 128 * tiles - 5
rec = torch.randn(n, 16, device=dev)


def run(env):
    os.environ.pop("NF_MLP_IMPL", None)
    os.environ.update(env)
    out = ops.nerf_mlp(packed, rec)
    ev = torch.cuda.Event(); ev.record()
    t0 = time.time()
    while not ev.query() and time.time() - t0 < 5:
        time.sleep(0.05)
    if not ev.query():
        print(f"tiles {tiles} {env}: no completion after 5 s"); sys.stdout.flush()
        os._exit(3)
    return out.clone()


out = run({})
if "tune" in os.path.basename(os.environ.get("NF_B200_LIB", "")):
    ref = run({"NF_MLP_IMPL": "1"})
    print(f"tiles {tiles}: max abs diff two-tile vs one-tile kernel {float((out - ref).abs().max()):.3e}")
else:
    print(f"tiles {tiles}: completed")
