"""GPU tests of the backward pass: the CUDA gradients against torch autograd through the CPU
oracle's differentiable restatement of the same functions.  Tolerance: 1e-2 relative L2 per parameter group (bf16
gradient operands on the tensor cores, fp32 accumulation; the reference computes fp32)."""
import numpy as np
import pytest
import torch

import neurofluid_b200 as nb
from neurofluid_b200 import _lib, ops, scenes
from oracle import renderer as orender
from helpers import load_render_case, rel_l2

pytestmark = pytest.mark.gpu
GRAD_TOL = 1e-2      # networks whose pre-activations keep a margin from the ReLU kinks: the arithmetic itself
KINK_TOL = 5e-2      # default-init networks: pre-activations are dense around 0, and a relative difference of 1e-4 between two
                     # forward evaluations (fp16 rounding boundaries of the encoded features) flips ~1e-4 of the ReLU masks per
                     # layer -- each flipped unit is an O(1) error of that unit's gradient, i.e. ~sqrt(9e-4) = 1-3 % relative
                     # L2 over nine layers.  Inherent to comparing two evaluations of a ReLU network's gradient; measured 1.2-1.8e-2.


def with_margin(sd):
    """Same architecture, weights x0.25 and hidden biases +-1 alternating: every hidden unit is either clearly on or clearly
    off (|pre-activation| > ~10 sigma of its input-dependent part), so both evaluations use identical ReLU masks."""
    out = {}
    for k, v in sd.items():
        v = v.clone()
        hidden = ("xyz_encoding_" in k and "final" not in k) or "dir_encoding" in k
        if k.endswith(".weight") and (hidden or "final" in k):
            v *= 0.25
        if k.endswith(".bias") and hidden:
            v = torch.where(torch.arange(v.numel()) % 2 == 0, torch.ones_like(v), -torch.ones_like(v))
        out[k] = v
    return out


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _param_views(flat, net):
    out, o = [], 0
    for p in net.ordered_params():
        out.append(flat[o:o + p.numel()].view(p.shape))
        o += p.numel()
    assert o == flat.numel()
    return out


@pytest.mark.parametrize("n,gain,margin", [(3000, 1.0, True), (128 * 7, 1.0, True), (3000, 1.0, False), (128 * 7, 2.0, False)])
def test_mlp_backward_vs_autograd(dev, n, gain, margin):
    rng = np.random.RandomState(7)
    sd = scenes.init_render_state(5, weight_gain=gain)
    if margin:
        sd = with_margin(sd)
    TOL = GRAD_TOL if margin else KINK_TOL
    net = nb.RenderNet(scenes.render_cfg(), scenes.NEAR, scenes.FAR)
    net.load_state_dict(sd)
    net = net.to(dev)
    rec = np.concatenate([rng.uniform(-1.5, 1.5, (n, 3)), rng.uniform(0, 12, (n, 1)), rng.uniform(-1.5, 1.5, (n, 3)),
                          rng.uniform(0, 0.02, (n, 3)), rng.randn(n, 6)], 1).astype(np.float32)
    rec[:, 10:13] /= np.linalg.norm(rec[:, 10:13], axis=1, keepdims=True)
    rec[:, 13:16] /= np.linalg.norm(rec[:, 13:16], axis=1, keepdims=True)
    rec = torch.from_numpy(rec)
    G = torch.from_numpy(rng.normal(0, 1e-3, (n, 4)).astype(np.float32))
    G[::7] = 0                                               # rows without gradient
    # ---- reference: autograd through the oracle MLP evaluated in the CUDA path's operand precision (fp16 operands of the
    #      ten tensor-core layers, straight-through; same ReLU kinks), on the fp32 encoded features
    pe = orender.positional_encoding
    with torch.enable_grad():
        sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("nerf_fine.")}
        feats = torch.cat([pe(rec[:, 0:3], 10), pe(rec[:, 3:4], 4), pe(rec[:, 4:7], 10), pe(rec[:, 7:10], 10),
                           pe(rec[:, 10:13], 4), pe(rec[:, 13:16], 4)], 1).requires_grad_(True)
        out = orender.nerf_mlp(sdg, "nerf_fine", feats, 198, 54, quant=orender.operand_rounding(torch.float16))
        (out * G).sum().backward()
    rgb = out[:, :3].detach()
    dout = torch.cat([G[:, :3] * rgb * (1 - rgb), G[:, 3:4]], 1)      # gradient w.r.t. (pre-sigmoid rgb, sigma)
    # ---- CUDA
    params = net.nerf_fine.ordered_params()
    pf = ops.pack_nerf_weights(params, _lib.NF_DTYPE_F16)
    pb = ops.pack_nerf_weights_bwd(params)
    dfeat, dpar = ops.nerf_mlp_backward(pf, pb, rec.to(dev), dout.to(dev))
    dfeat, dpar = dfeat.cpu(), dpar.cpu()
    assert rel_l2(dfeat[:, :198], feats.grad[:, :198]) < TOL, rel_l2(dfeat[:, :198], feats.grad[:, :198])
    assert rel_l2(dfeat[:, 208:262], feats.grad[:, 198:252]) < TOL
    assert (dfeat[:, 198:208] == 0).all() and (dfeat[:, 262:] == 0).all()
    names = [f"xyz_encoding_{i}.0" for i in range(1, 9)] + ["xyz_encoding_final", "dir_encoding.0", "sigma", "rgb.0"]
    views = _param_views(dpar, net.nerf_fine)
    for i, nm in enumerate(names):
        gw, gb = sdg[f"nerf_fine.{nm}.weight"].grad, sdg[f"nerf_fine.{nm}.bias"].grad
        assert rel_l2(views[2 * i], gw) < TOL, (nm, "weight", rel_l2(views[2 * i], gw))
        assert rel_l2(views[2 * i + 1], gb) < TOL, (nm, "bias", rel_l2(views[2 * i + 1], gb))


@pytest.mark.parametrize("name,mode,margin", [("small_boost", "forward", True), ("small_nomask", "forward", True), ("cfg0_sub", "forward", True),
                                              ("small_boost", "coarse", True), ("small_boost", "forward", False), ("cfg0_sub", "forward", False),
                                              ("small_wo_sdir", "forward", True), ("small_min_enc", "forward", True),
                                              ("small_incl_ray", "forward", True)])
def test_render_backward_vs_oracle_autograd(dev, name, mode, margin):
    """loss = mse(rgb0) + mse(rgb1) as in trainer/trainer_e2e.py:236-244; gradients w.r.t. every parameter tensor of both
    MLPs and w.r.t. the particle positions, against autograd through the oracle evaluated on the same merged depths."""
    c = load_render_case(name)
    if margin:
        c["sd"] = with_margin(c["sd"])
    TOL = GRAD_TOL if margin else KINK_TOL
    rng = np.random.RandomState(1)
    target = torch.from_numpy(rng.uniform(0, 1, (c["rays"].shape[0], 3)).astype(np.float32))
    net = nb.RenderNet(c["cfg"], scenes.NEAR, scenes.FAR)
    net.load_state_dict(c["sd"])
    net = net.to(dev)
    part = c["particles"].to(dev).requires_grad_(True)
    with torch.enable_grad():
        fn = net.forward if mode == "forward" else net.coarse_rendering
        out = fn(part, c["ro"].to(dev), c["rays"].to(dev), 1.0, c["cw"].to(dev))
        assert out["rgb0"].requires_grad and not out["mask_0"].requires_grad and not out["num_nn_0"].requires_grad
        loss = ((out["rgb0"] - target.to(dev)) ** 2).mean()
        if mode == "forward":
            loss = loss + ((out["rgb1"] - target.to(dev)) ** 2).mean()
        loss.backward()
    dbg = net.debug_view()
    # ---- oracle
    with torch.enable_grad():
        sdg = {k: v.clone().requires_grad_(True) if v.is_floating_point() else v for k, v in c["sd"].items()}
        pg = c["particles"].clone().requires_grad_(True)
        z1 = None
        if mode == "forward":
            hit = (out["num_nn_1"].sum((1, 2)) + out["num_nn_0"].sum((1, 2)) > 0).cpu()
            ref0 = orender.render_forward(c["sd"], c["cfg"], scenes.NEAR, scenes.FAR, c["particles"], c["ro"], c["rays"], debug=True)
            z1 = torch.where(hit[:, None], dbg["z1"].cpu(), ref0["dbg_z1"]) if bool(c["g"]["use_mask"]) else dbg["z1"].cpu()
        ref = orender.render_forward_grad(sdg, c["cfg"], scenes.NEAR, scenes.FAR, pg, c["ro"], c["rays"], mode=mode, z1_override=z1,
                                          quant=orender.operand_rounding(torch.float16))
        lref = ((ref["rgb0"] - target) ** 2).mean()
        if mode == "forward":
            lref = lref + ((ref["rgb1"] - target) ** 2).mean()
        lref.backward()
    assert abs(float(loss) - float(lref)) < 1e-4 * abs(float(lref))
    if pg.grad is None:        # every particle-dependent encoding switched off: the image does not depend on particle positions
        assert float(part.grad.abs().max()) == 0.0
    else:
        assert rel_l2(part.grad.cpu(), pg.grad) < TOL, ("particles", rel_l2(part.grad.cpu(), pg.grad))
    scale = max(float(sdg[k].grad.norm()) for k, _ in net.named_parameters() if sdg[k].grad is not None)
    for k, p in net.named_parameters():
        gref = sdg[k].grad
        if mode == "coarse" and k.startswith("nerf_fine"):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
            continue
        assert gref is not None and p.grad is not None, k
        if float(gref.norm()) < 1e-5 * scale:
            # a gradient that is pure cancellation noise next to the others (the sigma head of a saturated ray: d alpha / d sigma
            # = delta (1 - alpha) ~ 0): bounded, not compared digit for digit
            assert float(p.grad.norm()) < 1e-4 * scale, k
        else:
            assert rel_l2(p.grad.cpu(), gref) < TOL, (k, rel_l2(p.grad.cpu(), gref))


def test_render_backward_with_jitter_vs_oracle_autograd(dev):
    """perturb / noise_std are live in training: gradients with the reference's stored random draws."""
    c = load_render_case("small_jitter")
    c["sd"] = with_margin(c["sd"])
    g = c["g"]
    jit = {k: torch.from_numpy(g["draw." + k]) for k in ("z_rand", "noise0", "u", "noise1")}
    rng = np.random.RandomState(2)
    target = torch.from_numpy(rng.uniform(0, 1, (c["rays"].shape[0], 3)).astype(np.float32))
    net = nb.RenderNet(c["cfg"], scenes.NEAR, scenes.FAR)
    net.load_state_dict(c["sd"])
    net = net.to(dev)
    net._given_jitter = jit
    part = c["particles"].to(dev).requires_grad_(True)
    with torch.enable_grad():
        out = net(part, c["ro"].to(dev), c["rays"].to(dev), 1.0, c["cw"].to(dev), perturb=1.0, noise_std=0.5)
        loss = ((out["rgb0"] - target.to(dev)) ** 2).mean() + ((out["rgb1"] - target.to(dev)) ** 2).mean()
        loss.backward()
    z1 = net.debug_view()["z1"].cpu()
    with torch.enable_grad():
        sdg = {k: v.clone().requires_grad_(True) if v.is_floating_point() else v for k, v in c["sd"].items()}
        pg = c["particles"].clone().requires_grad_(True)
        ref = orender.render_forward_grad(sdg, c["cfg"], scenes.NEAR, scenes.FAR, pg, c["ro"], c["rays"], z1_override=z1,
                                          quant=orender.operand_rounding(torch.float16), perturb=1.0, noise_std=0.5, jitter=jit)
        lref = ((ref["rgb0"] - target) ** 2).mean() + ((ref["rgb1"] - target) ** 2).mean()
        lref.backward()
    assert abs(float(loss) - float(lref)) < 1e-4 * abs(float(lref))
    assert rel_l2(part.grad.cpu(), pg.grad) < GRAD_TOL, rel_l2(part.grad.cpu(), pg.grad)
    scale = max(float(sdg[k].grad.norm()) for k, _ in net.named_parameters())
    for k, p in net.named_parameters():
        gref = sdg[k].grad
        if float(gref.norm()) < 1e-5 * scale:
            assert float(p.grad.norm()) < 1e-4 * scale, k
        else:
            assert rel_l2(p.grad.cpu(), gref) < GRAD_TOL, (k, rel_l2(p.grad.cpu(), gref))


def test_render_training_step_changes_the_loss(dev):
    """Several forwards before one backward (one per view, trainer/trainer_e2e.py:219-244), Adam on the renderer."""
    c = load_render_case("small_boost")
    net = nb.RenderNet(c["cfg"], scenes.NEAR, scenes.FAR)
    net.load_state_dict(c["sd"])
    net = net.to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=5e-4)
    target = torch.full((c["rays"].shape[0], 3), 0.3, device=dev)
    p, ro, rays, cw = c["particles"].to(dev), c["ro"].to(dev), c["rays"].to(dev), c["cw"].to(dev)
    losses = []
    with torch.enable_grad():
        for it in range(6):
            total = 0.
            for view in range(2):
                r = net(p, ro, rays if view == 0 else rays.flip(0), 1.0, cw)
                t = target if view == 0 else target.flip(0)
                total = total + ((r["rgb0"] - t) ** 2).mean() + ((r["rgb1"] - t) ** 2).mean()
            opt.zero_grad()
            total.backward()
            opt.step()
            losses.append(float(total))
    assert losses[-1] < losses[0], losses


def _transition_case(n_lat=8, seed=3):
    rng = np.random.RandomState(seed)
    sd = scenes.init_particle_state(seed, last_layer_scale=1.0)
    pos = torch.from_numpy(scenes.lattice_particles(n_lat, seed, jitter=0.02, center=(0.1, -0.2, -0.7)))
    vel = torch.from_numpy(rng.normal(0, 0.5, pos.shape).astype(np.float32))
    bp, bn = scenes.box_points(0.06)
    return sd, pos, vel, torch.from_numpy(bp), torch.from_numpy(bn), rng


def test_transition_backward_two_step_unroll_vs_oracle_autograd(dev):
    """trainer/trainer_transmodel.py:179-197: two unrolled steps, loss on both predictions; gradients w.r.t. all 18 parameter
    tensors and w.r.t. the initial positions / velocities against autograd through the pure-torch ContinuousConv oracle."""
    from oracle import transition as otrans
    sd, pos, vel, box, box_n, rng = _transition_case()
    gt1 = pos + torch.from_numpy(rng.normal(0, 0.01, pos.shape).astype(np.float32))
    gt2 = pos + torch.from_numpy(rng.normal(0, 0.02, pos.shape).astype(np.float32))
    net = nb.ParticleNet(gravity=(0.0, 0.0, -9.81))
    net.load_state_dict(sd)
    net = net.to(dev)

    def loss_fn(p1, v1, n1, p2, v2, n2, g1, g2):
        w1, w2 = torch.exp(-n1 / 40.0), torch.exp(-n2 / 40.0)          # neighbour-count weighting, trainer_transmodel.py:73-89
        return 0.5 * (w1 * ((p1 - g1) ** 2).sum(-1)).mean() + 0.5 * (w2 * ((p2 - g2) ** 2).sum(-1)).mean() + 1e-3 * (v2 ** 2).mean()

    p0 = pos.to(dev).requires_grad_(True)
    v0 = vel.to(dev).requires_grad_(True)
    with torch.enable_grad():
        p1, v1, n1 = net(p0, v0, box.to(dev), box_n.to(dev))
        assert p1.requires_grad and v1.requires_grad and not n1.requires_grad
        p2, v2, n2 = net(p1, v1, box.to(dev), box_n.to(dev))
        loss = loss_fn(p1, v1, n1, p2, v2, n2, gt1.to(dev), gt2.to(dev))
        loss.backward()
    net.check_neighbor_overflow()
    with torch.enable_grad():
        sdg = {k: (v.clone().requires_grad_(True) if k != "gravity" and "offset" not in k else v.clone()) for k, v in sd.items()}
        rp0, rv0 = pos.clone().requires_grad_(True), vel.clone().requires_grad_(True)
        qf = orender.operand_rounding(torch.float16)
        r1 = otrans.particle_step_grad(sdg, rp0, rv0, box, box_n, quant=qf)
        r2 = otrans.particle_step_grad(sdg, r1[0], r1[1], box, box_n, quant=qf)
        lref = loss_fn(*r1, *r2, gt1, gt2)
        lref.backward()
    assert torch.equal(n1.cpu(), r1[2]) and torch.equal(n2.cpu(), r2[2])
    assert abs(float(loss) - float(lref)) < 1e-5 * abs(float(lref))
    assert rel_l2(p0.grad.cpu(), rp0.grad) < GRAD_TOL, ("pos", rel_l2(p0.grad.cpu(), rp0.grad))
    assert rel_l2(v0.grad.cpu(), rv0.grad) < GRAD_TOL, ("vel", rel_l2(v0.grad.cpu(), rv0.grad))
    for k, p in net.named_parameters():
        gref = sdg[k].grad
        assert gref is not None and p.grad is not None, k
        assert rel_l2(p.grad.cpu(), gref) < GRAD_TOL, (k, rel_l2(p.grad.cpu(), gref))


def test_end2end_train_step_like_trainer_e2e(dev):
    """trainer/trainer_e2e.py:189-302 through the drop-in modules: transition step -> render random rays of two views ->
    mse(rgb0) + mse(rgb1) + boundary L1 -> backward into BOTH networks -> two Adam steps.  Gradients of the transition
    model (which receive everything through d rgb / d particle position) against the oracle's autograd."""
    from oracle import transition as otrans
    sd_t, pos, vel, box, box_n, rng = _transition_case(n_lat=9, seed=4)
    pos = pos - torch.tensor([0.1, -0.2, -0.7]) + torch.tensor([0.0, 0.0, 0.0])      # centred: the camera looks at the origin
    cfg = scenes.render_cfg()
    sd_r = with_margin(scenes.init_render_state(0, 5.0))
    tn = nb.ParticleNet(gravity=(0.0, 0.0, -9.81)); tn.load_state_dict(sd_t); tn = tn.to(dev)
    rn = nb.RenderNet(cfg, scenes.NEAR, scenes.FAR); rn.load_state_dict(sd_r); rn = rn.to(dev)
    H = 40
    rays, focal, cw = scenes.camera_rays(H, H)
    rays = scenes.center_crop_rays(rays, H, H, 12)
    target = torch.from_numpy(rng.uniform(0, 1, (rays.shape[0], 3)).astype(np.float32))
    opt_r = torch.optim.Adam(rn.parameters(), lr=1e-4)
    opt_t = torch.optim.Adam(tn.parameters(), lr=1e-5)
    with torch.enable_grad():
        pred_pos, pred_vel, _ = tn(pos.to(dev), vel.to(dev), box.to(dev), box_n.to(dev))
        total = 0.
        for view in range(2):
            rr = rays if view == 0 else rays.flip(0)
            tt = target if view == 0 else target.flip(0)
            r = rn(pred_pos, cw[:, 3].to(dev), rr.to(dev), focal, cw.to(dev))
            total = total + ((r["rgb0"] - tt.to(dev)) ** 2).mean() + ((r["rgb1"] - tt.to(dev)) ** 2).mean()
        bd = torch.relu(pred_pos.abs() - 0.9).mean()                    # cal_boundary_loss-style L1 term
        total = total + 0.1 * bd
        opt_r.zero_grad(); opt_t.zero_grad()
        total.backward()
    g_before = {k: p.grad.clone() for k, p in tn.named_parameters()}
    assert all(torch.isfinite(g).all() for g in g_before.values())
    assert sum(float(g.abs().sum()) for g in g_before.values()) > 0      # the image loss reached the transition model
    w_before = tn.conv1.kernel.detach().clone()
    opt_r.step(); opt_t.step()
    assert not torch.equal(w_before, tn.conv1.kernel.detach())
    # ---- oracle: same graph on the CPU (merged depths of the fine pass taken from the CUDA forward's neighbour counts is not
    #      possible for two views cheaply; the coarse-only loss is compared instead)
    with torch.enable_grad():
        tn2 = nb.ParticleNet(gravity=(0.0, 0.0, -9.81)); tn2.load_state_dict(sd_t); tn2 = tn2.to(dev)
        pp, _, _ = tn2(pos.to(dev), vel.to(dev), box.to(dev), box_n.to(dev))
        r = rn.coarse_rendering(pp, cw[:, 3].to(dev), rays.to(dev), focal, cw.to(dev))
        lc = ((r["rgb0"] - target.to(dev)) ** 2).mean()
        for p in rn.parameters():
            p.grad = None
        lc.backward()
        sdg = {k: (v.clone().requires_grad_(True) if k != "gravity" and "offset" not in k else v.clone()) for k, v in sd_t.items()}
        sdr = {k: v.detach().cpu().clone() for k, v in rn.state_dict().items()}
        qf = orender.operand_rounding(torch.float16)
        op, _, _ = otrans.particle_step_grad(sdg, pos, vel, box, box_n, quant=qf)
        ref = orender.render_forward_grad(sdr, cfg, scenes.NEAR, scenes.FAR, op, cw[:, 3], rays, mode="coarse", quant=qf)
        lr = ((ref["rgb0"] - target) ** 2).mean()
        lr.backward()
    assert abs(float(lc) - float(lr)) < 1e-4 * abs(float(lr))
    for k, p in tn2.named_parameters():
        assert rel_l2(p.grad.cpu(), sdg[k].grad) < 2 * GRAD_TOL, (k, rel_l2(p.grad.cpu(), sdg[k].grad))


def test_training_steps_release_their_workspaces(dev):
    """Every training forward owns a workspace (hundreds of MB at scene size) that must die with its graph, by reference
    counting: with the garbage collector switched off the allocated memory may not grow from step to step."""
    import gc
    sd, pos, vel, box, box_n, rng = _transition_case()
    tn = nb.ParticleNet(gravity=(0.0, 0.0, -9.81)); tn.load_state_dict(sd); tn = tn.to(dev)
    rn = nb.RenderNet(scenes.render_cfg(), scenes.NEAR, scenes.FAR); rn.load_state_dict(scenes.init_render_state(0, 5.0)); rn = rn.to(dev)
    rays, focal, cw = scenes.camera_rays(40, 40)
    rays = scenes.center_crop_rays(rays, 40, 40, 12).to(dev)
    args = [t.to(dev) for t in (pos - torch.tensor([0.1, -0.2, -0.7]), vel, box, box_n)]
    gc.collect()
    gc.disable()
    try:
        seen = []
        for it in range(6):
            with torch.enable_grad():
                pp, vv, _ = tn(*args)
                r = rn(pp, cw[:, 3].to(dev), rays, focal, cw.to(dev))
                loss = (r["rgb1"] ** 2).mean() + (r["rgb0"] ** 2).mean() + (vv ** 2).mean()
                for p in list(tn.parameters()) + list(rn.parameters()):
                    p.grad = None
                loss.backward()
            del pp, vv, r, loss
            torch.cuda.synchronize()
            seen.append(torch.cuda.memory_allocated(dev))
    finally:
        gc.enable()
    assert max(seen[3:]) <= max(seen[1:3]), seen      # (caches alternate between two sizes; growth is the failure)
