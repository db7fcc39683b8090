import os, sys, torch, numpy as np
torch.set_grad_enabled(False)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import neurofluid_b200 as nb
from neurofluid_b200 import scenes, _lib, ops
dev = torch.device("cuda:0")
net = nb.RenderNet(scenes.render_cfg(), 9., 13.); net.load_state_dict(scenes.init_render_state(0)); net = net.to(dev)
packed = ops.pack_nerf_weights(net.nerf_fine.ordered_params())
n = 128 * 148 * 16
rec = torch.randn(n, 16, device=dev)
trace = torch.zeros(512, dtype=torch.int64, device=dev)
for _ in range(2): ops.nerf_mlp(packed, rec)
os.environ["NF_MLP_TRACE_PTR"] = str(trace.data_ptr())
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); ops.nerf_mlp(packed, rec); t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1)
print("ms", ms, "TFLOP/s", 2 * 665984 * n / ms / 1e9)
tr = trace.cpu().numpy()
base = tr[100]
f = lambda v: int(v - base) if v else None
for l in range(10):
    print(f"L{l}: issue start {f(tr[100+l*12])} units(after T1 issue) {[f(tr[100+l*12+1+u]) for u in range(4)]} | "
          f"epi (h,T) full/done {[(f(tr[l*8+k*2]), f(tr[l*8+k*2+1])) for k in range(4)]}")
