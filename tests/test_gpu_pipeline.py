"""GPU parity tests for the caller-side kernels (csrc/nf_aux.cu) and host mirrors (neurofluid_b200/pipeline.py):
camera rays, nearest-neighbour metrics, PSNR, the chunked render_image loop and the eval_e2e-shaped rollout."""
import numpy as np
import pytest
import torch

import neurofluid_b200 as nb
from neurofluid_b200 import ops, pipeline, scenes
from oracle import metrics as ometrics
from helpers import load_render_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("H,W", [(64, 64), (37, 53), (400, 400)])
def test_generate_rays_matches_reference_host_rays(dev, H, W):
    ref, focal, cw = scenes.camera_rays(H, W)          # bit-identical to utils/ray_utils.py get_rays (oracle/make_golden.py)
    out = ops.generate_rays(H, W, focal, cw.to(dev)).cpu()
    assert out.shape == (H * W, 6)
    assert torch.equal(out[:, :3], ref[:, :3])                        # origins are copies
    assert (out[:, 3:] - ref[:, 3:]).abs().max() < 2e-7               # unit directions: <= 1-2 ulp (matmul rounding)
    assert (out[:, 3:].norm(dim=1) - 1).abs().max() < 2e-7
    assert ops.generate_rays(0, 5, focal, cw.to(dev)).shape == (0, 6)


def test_nearest_distance_matches_ckdtree(dev):
    rng = np.random.RandomState(1)
    pred = rng.uniform(-1, 1, (5000, 3)).astype(np.float32)
    gt = (pred + rng.normal(0, 0.03, pred.shape)).astype(np.float32)
    gt[:20] = rng.uniform(-6, 6, (20, 3))                             # far outside the grid
    gt[20:40] = pred[20:40]                                           # zero distance
    ref = ometrics.gt_to_pred_distance(pred, gt)
    for cell in (0.03, 0.1, 0.7):
        d, idx = ops.nearest_distance(torch.from_numpy(gt).to(dev), torch.from_numpy(pred).to(dev), cell=cell,
                                      return_index=True)
        assert np.allclose(d.cpu().numpy(), ref, rtol=2e-6, atol=1e-7), cell
        chosen = np.linalg.norm(gt.astype(np.float64) - pred[idx.cpu().numpy()].astype(np.float64), axis=1)
        assert np.allclose(chosen, ref, rtol=2e-6, atol=1e-7)
    one = ops.nearest_distance(torch.zeros(3, 3, device=dev), torch.ones(1, 3, device=dev))
    assert torch.allclose(one.cpu(), torch.full((3,), 3 ** 0.5))
    assert ops.nearest_distance(torch.zeros(0, 3, device=dev), torch.ones(4, 3, device=dev)).shape == (0,)
    assert (ops.nearest_distance(torch.zeros(2, 3, device=dev), torch.ones(0, 3, device=dev)) > 1e30).all()   # empty set


def test_fluid_errors_and_psnr_match_reference_formulas(dev):
    rng = np.random.RandomState(2)
    pred = rng.uniform(-0.5, 0.5, (3001, 3)).astype(np.float32)       # odd and even counts exercise np.median
    gt = (pred + rng.normal(0, 0.01, pred.shape)).astype(np.float32)
    fe = pipeline.FluidErrors()
    for t, n in enumerate((3001, 3000)):
        got = fe.cal_errors(torch.from_numpy(pred[:n]).to(dev), torch.from_numpy(gt[:n]).to(dev), t + 1)
        ref = ometrics.fluid_errors(pred[:n], gt[:n])
        assert abs(float(got) - ref["gt2pred_mean"]) < 1e-4 * ref["gt2pred_mean"]
        e = fe.errors[t + 1]
        for k, v in ref.items():
            assert abs(e[k] - v) <= 2e-5 * abs(v) + 1e-9, (k, e[k], v)
    x, y = torch.rand(400, 400, 3), torch.rand(400, 400, 3)
    mse = ops.img2mse(x.to(dev), y.to(dev))
    assert abs(float(mse) - ometrics.img2mse(x.numpy(), y.numpy())) < 1e-7
    assert abs(float(ops.mse2psnr(mse)) - ometrics.mse2psnr(ometrics.img2mse(x.numpy(), y.numpy()))) < 1e-4


def test_render_image_chunk_loop_is_bitwise_the_forward(dev):
    c = load_render_case("cfg0_sub")
    net = nb.RenderNet(c["cfg"], scenes.NEAR, scenes.FAR)
    net.load_state_dict(c["sd"])
    net = net.to(dev)
    p, ro, rays, cw = c["particles"].to(dev), c["ro"].to(dev), c["rays"].to(dev), c["cw"].to(dev)
    whole = net(p, ro, rays, 1.0, cw)
    ret = pipeline.render_image(net, p, rays.shape[0], ro, rays, 1.0, cw, ray_chunk=net.cfg.ray.ray_chunk, iseval=True)      # the reference's loop: 1024-ray chunks
    assert set(ret) == {"pred_rgbs_0", "num_nn_0", "mask_0", "pred_rgbs_1", "num_nn_1", "mask_1"}
    assert torch.equal(ret["pred_rgbs_1"], whole["rgb1"]) and torch.equal(ret["pred_rgbs_0"], whole["rgb0"])
    assert torch.equal(ret["num_nn_1"], whole["num_nn_1"].view(-1)) and torch.equal(ret["mask_0"], whole["mask_0"])
    lean = pipeline.render_image(net, p, rays.shape[0], ro, rays, 1.0, cw, ray_chunk=300)
    assert set(lean) == {"pred_rgbs_0", "num_nn_0", "pred_rgbs_1", "num_nn_1"}
    assert torch.equal(lean["pred_rgbs_1"], whole["rgb1"])
    one = pipeline.render_image(net, p, rays.shape[0], ro, rays, 1.0, cw)        # default: the whole image in one call
    assert torch.equal(one["pred_rgbs_1"], whole["rgb1"]) and torch.equal(one["num_nn_0"], whole["num_nn_0"].view(-1))


def test_rollout_and_render_is_the_eval_loop(dev):
    H = W = 48
    tn = nb.ParticleNet(gravity=(0.0, 0.0, -9.81))
    tn.load_state_dict(scenes.init_particle_state(0))
    tn = tn.to(dev)
    rn = nb.RenderNet(scenes.render_cfg(), scenes.NEAR, scenes.FAR)
    rn.load_state_dict(scenes.init_render_state(0, 5.0))
    rn = rn.to(dev)
    half = 9 / 2 * 0.05
    pos = torch.from_numpy(scenes.lattice_particles(10, 0, center=(0.0, 0.0, -1 + 0.03 + half))).to(dev)
    vel = torch.zeros_like(pos)
    bp, bn = scenes.box_points(0.08)
    box, box_n = torch.from_numpy(bp).to(dev), torch.from_numpy(bn).to(dev)
    _, focal, cw = scenes.camera_rays(H, W)
    cams = [(cw, focal), (cw, 0.9 * focal)]
    # "ground truth": the same rollout, so position errors are exactly 0 and PSNR is infinite
    p, v = pos, vel
    gt_pos, gt_img = [], []
    for f in range(3):
        p, v, _ = tn(p, v, box, box_n)
        p, v = p.clone(), v.clone()
        gt_pos.append(p)
        imgs = []
        for c2w, fo in cams:
            rays = ops.generate_rays(H, W, fo, c2w.to(dev))
            imgs.append(rn(p, c2w[:, 3].to(dev), rays, fo, c2w.to(dev))["rgb1"].clone())
        gt_img.append(imgs)
    out = pipeline.rollout_and_render(tn, rn, pos, vel, box, box_n, cams, H, W, 3, gt_positions=gt_pos, gt_images=gt_img)
    for f in range(3):
        assert torch.equal(out["positions"][f], gt_pos[f])
        for vi in range(2):
            assert torch.equal(out["images"][f][vi], gt_img[f][vi])
    assert torch.isinf(out["psnr"]).all() and out["psnr"].shape == (3, 2)
    errs = out["fluid_errors"].errors
    assert sorted(errs) == [1, 2, 3] and all(e["gt2pred_max"] == 0.0 and e["mean"] == 0.0 for e in errs.values())
    # against a shifted ground truth the distances are the shift
    shifted = [g + torch.tensor([0.0, 0.0, 0.004], device=dev) for g in gt_pos]
    out2 = pipeline.rollout_and_render(tn, rn, pos, vel, box, box_n, cams[:1], H, W, 1, gt_positions=shifted)
    e = out2["fluid_errors"].errors[1]
    assert abs(e["mean"] - 4.0) < 1e-3 and e["gt2pred_mean"] <= e["mean"] + 1e-6


def test_config3_rollout_and_render_against_the_oracle(dev):
    """BASELINE config[3] shape (eval_e2e.py:58-120: per frame one transition step, then a rendered view) against the
    CPU oracle running the same loop on its OWN trajectory: 23^3 = 12,167 particles, 3 frames, a strided sample of the
    400x400 image per frame.  Positions <= 1e-5, rgb <= 1e-3 relative L2 (north_star)."""
    from oracle import renderer as orender
    from oracle import transition as otrans
    from helpers import rel_l2
    n, H = 23, 400
    half = (n - 1) / 2 * 0.05
    sd_t, sd_r, cfg = scenes.init_particle_state(0), scenes.init_render_state(0, 5.0), scenes.render_cfg()
    tn = nb.ParticleNet(gravity=(0.0, 0.0, -9.81)); tn.load_state_dict(sd_t); tn = tn.to(dev)
    rn = nb.RenderNet(cfg, scenes.NEAR, scenes.FAR); rn.load_state_dict(sd_r); rn = rn.to(dev)
    pos = torch.from_numpy(scenes.lattice_particles(n, 0, center=(0.0, 0.0, -1 + 0.03 + half)))
    vel = torch.zeros_like(pos)
    bp, bn = scenes.box_points(0.032)
    box, box_n = torch.from_numpy(bp), torch.from_numpy(bn)
    rays, focal, cw = scenes.camera_rays(H, H)
    cw = cw.clone(); cw[2, 3] += -1 + 0.03 + half                  # aim the camera at the block on the box floor
    out = pipeline.rollout_and_render(tn, rn, pos.to(dev), vel.to(dev), box.to(dev), box_n.to(dev), [(cw, focal)], H, H, 3)
    rays = ops.generate_rays(H, H, focal, cw.to(dev)).cpu()
    sel = torch.arange(0, H * H, (H * H) // 300)[:300]
    op, ov = pos, vel
    for f in range(3):
        op, ov, _ = otrans.particle_step(sd_t, op, ov, box, box_n)
        assert rel_l2(out["positions"][f].cpu(), op) < 1e-5, f
        ref = orender.render_forward(sd_r, cfg, scenes.NEAR, scenes.FAR, op, cw[:, 3], rays[sel])
        got = out["images"][f][0].cpu()[sel]
        assert (ref["rgb1"] < 0.999).any()                         # the sample does see the fluid
        assert rel_l2(got, ref["rgb1"]) < 1e-3, (f, rel_l2(got, ref["rgb1"]))


def test_config4_50k_particles_800x800_spot_check(dev):
    """BASELINE config[4] size on one GPU: 37^3 = 50,653 particles (the sweep search's bitmap grows with P: fewer
    resident blocks), whole 800x800 image in one call, 512 strided rays against the CPU oracle."""
    from oracle import renderer as orender
    from helpers import rel_l2
    H = 800
    rays, focal, cw = scenes.camera_rays(H, H)
    particles = torch.from_numpy(scenes.lattice_particles(37, 0))
    cfg, sd = scenes.render_cfg(), scenes.init_render_state(0, 5.0)
    net = nb.RenderNet(cfg, scenes.NEAR, scenes.FAR); net.load_state_dict(sd); net = net.to(dev)
    out = net(particles.to(dev), cw[:, 3].to(dev), rays.to(dev), focal, cw.to(dev))
    o = {k: v.cpu() for k, v in out.items()}
    assert torch.isfinite(o["rgb1"]).all()
    assert torch.equal(o["mask_1"].view(-1), (o["num_nn_1"] == 20).sum(1).view(-1).float())
    sel = torch.arange(0, H * H, (H * H) // 512)[:512]
    ref = orender.render_forward(sd, cfg, scenes.NEAR, scenes.FAR, particles, cw[:, 3], rays[sel])
    assert torch.equal(o["num_nn_0"][sel], ref["num_nn_0"])
    assert (ref["num_nn_0"] == 20).any()
    assert rel_l2(o["rgb0"][sel], ref["rgb0"]) < 1e-3 and rel_l2(o["rgb1"][sel], ref["rgb1"]) < 1e-3
    assert (o["num_nn_1"][sel] != ref["num_nn_1"]).float().mean() < 2e-4
    # both search flavours at this size, on the rays through the fluid
    sub = rays.view(H, H, 6)[300:500:4, 300:500:4].reshape(-1, 6).contiguous()
    a = nb.RenderNet(cfg, scenes.NEAR, scenes.FAR, search="stream"); a.load_state_dict(sd); a = a.to(dev)
    ra = a(particles.to(dev), cw[:, 3].to(dev), sub.to(dev), focal, cw.to(dev))
    rb = net(particles.to(dev), cw[:, 3].to(dev), sub.to(dev), focal, cw.to(dev))
    for k in ra:
        assert torch.equal(ra[k], rb[k]), k
