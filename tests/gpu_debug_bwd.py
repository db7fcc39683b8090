import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurofluid_b200 as nb
from neurofluid_b200 import scenes
from oracle import renderer as orender
from helpers import load_render_case, rel_l2
from test_gpu_backward import with_margin
dev = torch.device("cuda:0")
name, mode = sys.argv[1], sys.argv[2]
c = load_render_case(name); c["sd"] = with_margin(c["sd"])
rng = np.random.RandomState(1)
target = torch.from_numpy(rng.uniform(0, 1, (c["rays"].shape[0], 3)).astype(np.float32))
net = nb.RenderNet(c["cfg"], scenes.NEAR, scenes.FAR); net.load_state_dict(c["sd"]); net = net.to(dev)
part = c["particles"].to(dev).requires_grad_(True)
out = net(part, c["ro"].to(dev), c["rays"].to(dev), 1.0, c["cw"].to(dev))
loss = ((out["rgb0"] - target.to(dev)) ** 2).mean() + ((out["rgb1"] - target.to(dev)) ** 2).mean()
loss.backward()
dbg = net.debug_view()
sdg = {k: v.clone().requires_grad_(True) if v.is_floating_point() else v for k, v in c["sd"].items()}
pg = c["particles"].clone().requires_grad_(True)
z1 = dbg["z1"].cpu()
if bool(c["g"]["use_mask"]):
    hit = (out["num_nn_1"].sum((1, 2)) + out["num_nn_0"].sum((1, 2)) > 0).cpu()
    ref0 = orender.render_forward(c["sd"], c["cfg"], scenes.NEAR, scenes.FAR, c["particles"], c["ro"], c["rays"], debug=True)
    z1 = torch.where(hit[:, None], z1, ref0["dbg_z1"])
ref = orender.render_forward_grad(sdg, c["cfg"], scenes.NEAR, scenes.FAR, pg, c["ro"], c["rays"], z1_override=z1, quant=orender.operand_rounding(torch.float16))
lref = ((ref["rgb0"] - target) ** 2).mean() + ((ref["rgb1"] - target) ** 2).mean()
lref.backward()
print("loss", float(loss), float(lref), "particles", rel_l2(part.grad.cpu(), pg.grad))
for k, p in net.named_parameters():
    g = sdg[k].grad
    print(f"{k:45s} rel {rel_l2(p.grad.cpu(), g):.3e}  |ref| {float(g.norm()):.3e} |got| {float(p.grad.norm()):.3e}")
