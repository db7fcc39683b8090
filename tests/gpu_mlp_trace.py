import os, sys, torch, numpy as np
torch.set_grad_enabled(False)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import neurofluid_b200 as nb
from neurofluid_b200 import scenes, _lib, ops
dev = torch.device("cuda:0")
net = nb.RenderNet(scenes.render_cfg(), 9., 13.); net.load_state_dict(scenes.init_render_state(0)); net = net.to(dev)
packed = ops.pack_nerf_weights(net.nerf_fine.ordered_params())
n = 128 * 148 * 8
rec = torch.randn(n, 16, device=dev)
trace = torch.zeros(512, dtype=torch.int64, device=dev)
for _ in range(2): ops.nerf_mlp(packed, rec)
os.environ["NF_MLP_TRACE_PTR"] = str(trace.data_ptr())
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); ops.nerf_mlp(packed, rec); t1.record(); torch.cuda.synchronize()
print("ms", t0.elapsed_time(t1), "tiles/CTA 8 -> us/tile", t0.elapsed_time(t1) * 1e3 / 8)
tr = trace.cpu().numpy()
base = tr[100]  # MMA starts layer 0
f = lambda v: int(v - base) if v else None
for l in range(10):
    print(f"L{l}: mma start {f(tr[100+l*8])} act_ready {[f(tr[100+l*8+1+c]) for c in range(4)]} commit-issued {f(tr[100+l*8+5])} | "
          f"epi accfull {f(tr[l*8])} chunks {[f(tr[l*8+1+c]) for c in range(4)]}")
print("PE: start", f(tr[300]), "xyz ready", f(tr[301]), "dir ready", f(tr[302]))
