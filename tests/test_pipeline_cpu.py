"""CPU tests of the caller-side mirrors: metric oracle vs brute force, checkpoint loaders
(trainer/basetrainer.py:87-122, eval_e2e.py:50-55).  No GPU compute."""
import numpy as np
import pytest
import torch

import neurofluid_b200 as nb
from neurofluid_b200 import pipeline, scenes
from neurofluid_b200._lib import NFError
from oracle import metrics as ometrics


def test_metric_oracle_against_brute_force():
    rng = np.random.RandomState(0)
    pred = rng.uniform(-1, 1, (700, 3)).astype(np.float32)
    gt = (pred + rng.normal(0, 0.02, pred.shape)).astype(np.float32)
    d = ometrics.gt_to_pred_distance(pred, gt)
    brute = np.sqrt(((gt[:, None, :].astype(np.float64) - pred[None].astype(np.float64)) ** 2).sum(-1)).min(1)
    assert np.allclose(d, brute, rtol=0, atol=1e-12)
    e = ometrics.fluid_errors(pred, gt)
    assert set(e) == {"mean", "mse", "var", "min", "max", "median", "num_particles"} | \
        {"gt2pred_" + k for k in ("mean", "mse", "var", "min", "max", "median", "num_particles")}
    assert e["gt2pred_mean"] <= e["mean"] + 1e-9 and e["num_particles"] == 700
    assert abs(ometrics.mse2psnr(ometrics.img2mse(np.zeros(10), np.full(10, 0.1))) - 20.0) < 1e-9


def test_checkpoint_loaders_follow_the_reference(tmp_path):
    r_src = nb.RenderNet(scenes.render_cfg(), scenes.NEAR, scenes.FAR)
    r_src.load_state_dict(scenes.init_render_state(3))
    t_src = nb.ParticleNet(gravity=(0.0, 0.0, -9.81))
    t_src.load_state_dict(scenes.init_particle_state(3))
    path = tmp_path / "ckpt.pt"
    torch.save({"renderer_state_dict": r_src.state_dict(), "transition_model_state_dict": t_src.state_dict()}, path)

    r = nb.RenderNet(scenes.render_cfg(), scenes.NEAR, scenes.FAR)
    t = nb.ParticleNet(gravity=(0.0, -9.81, 0.0))
    pipeline.resume(r, t, str(path))
    assert all(torch.equal(a, b) for a, b in zip(r.state_dict().values(), r_src.state_dict().values()))
    assert torch.equal(t.gravity, t_src.gravity)

    # load_pretained_transition_model: gravity is NOT taken from the checkpoint (basetrainer.py:98)
    t2 = nb.ParticleNet(gravity=(0.0, -9.81, 0.0))
    for wrap in ("transition_model_state_dict", "model_state_dict", None):
        ck = t_src.state_dict() if wrap is None else {wrap: t_src.state_dict()}
        pipeline.load_pretrained_transition_model(t2, ck)
        assert torch.equal(t2.conv1.kernel, t_src.conv1.kernel)
        assert torch.equal(t2.gravity, torch.tensor([0.0, -9.81, 0.0]))

    # partial_load keeps only sigma / xyz_encoding layers (basetrainer.py:113-115)
    r2 = nb.RenderNet(scenes.render_cfg(), scenes.NEAR, scenes.FAR)
    r2.load_state_dict(scenes.init_render_state(9))
    before = {k: v.clone() for k, v in r2.state_dict().items()}
    pipeline.load_pretrained_renderer_model(r2, str(path), partial_load=True)
    for k, v in r2.state_dict().items():
        if "sigma" in k or "xyz_encoding" in k:
            assert torch.equal(v, r_src.state_dict()[k]), k
        else:
            assert torch.equal(v, before[k]), k
    pipeline.load_pretrained_renderer_model(r2, str(path))
    assert all(torch.equal(a, b) for a, b in zip(r2.state_dict().values(), r_src.state_dict().values()))
    with pytest.raises(KeyError):
        pipeline.load_pretrained_renderer_model(r2, {"model_state_dict": {}})


def test_render_image_needs_the_device():
    r = nb.RenderNet(scenes.render_cfg(), scenes.NEAR, scenes.FAR)
    rays, focal, cw = scenes.camera_rays(8, 8)
    with pytest.raises(NFError):          # CPU tensors: no fallback
        pipeline.render_image(r, torch.zeros(10, 3), rays.shape[0], cw[:, 3], rays, focal, cw)


def test_render_image_chunk_loop_with_a_stand_in_renderer():
    """trainer/basetrainer.py:264-309: the chunk loop itself (host logic): one call for the whole image by default, the
    reference's ray_chunk loop on request, results concatenated in ray order."""
    from types import SimpleNamespace

    class Fake:
        cfg = SimpleNamespace(ray=SimpleNamespace(ray_chunk=1024, N_importance=128))

        def __init__(self):
            self.calls = []

        def __call__(self, particles, ro, rays, focal, cw):
            self.calls.append(rays.shape[0])
            idx = rays[:, 0:1]
            return {"rgb0": idx.repeat(1, 3), "rgb1": 2 * idx.repeat(1, 3), "num_nn_0": idx.long().view(-1, 1, 1),
                    "num_nn_1": idx.long().view(-1, 1, 1), "mask_0": idx, "mask_1": idx}

    rays = torch.arange(2500, dtype=torch.float32).view(-1, 1).repeat(1, 6)
    f = Fake()
    one = pipeline.render_image(f, None, rays.shape[0], None, rays, 1.0, None, iseval=True)
    assert f.calls == [2500]
    f = Fake()
    ref = pipeline.render_image(f, None, rays.shape[0], None, rays, 1.0, None, ray_chunk=f.cfg.ray.ray_chunk, iseval=True)
    assert f.calls == [1024, 1024, 452]
    assert set(one) == set(ref) == {"pred_rgbs_0", "num_nn_0", "mask_0", "pred_rgbs_1", "num_nn_1", "mask_1"}
    for k in one:
        assert torch.equal(one[k], ref[k]), k
    assert torch.equal(one["pred_rgbs_1"][:, 0], 2 * torch.arange(2500, dtype=torch.float32))
