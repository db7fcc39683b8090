"""CPU tests: pin the oracle restatements (oracle/renderer.py, oracle/transition.py) to the golden
vectors that oracle/make_golden.py produced from the reference's own unmodified Python, and
cross-check the compiled third-party-op restatements against their pure-torch twins."""
import numpy as np
import pytest
import torch

from oracle import renderer as orender
from oracle import third_party_ops as tpo
from oracle import transition as otrans
from neurofluid_b200 import scenes
from helpers import ABLATION_CASES, RENDER_CASES, TRANSITION_CASES, load_render_case, load_transition_case, rel_l2


@pytest.mark.parametrize("name", RENDER_CASES + ABLATION_CASES)
def test_render_oracle_matches_reference_golden(name):
    c = load_render_case(name)
    g = c["g"]
    out = orender.render_forward(c["sd"], c["cfg"], scenes.NEAR, scenes.FAR, c["particles"], c["ro"], c["rays"])
    for k in ("num_nn_0", "num_nn_1"):
        assert np.array_equal(out[k].numpy().astype(np.int8), g[f"forward.{k}"]), k
    for k in ("mask_0", "mask_1"):
        assert np.array_equal(out[k].numpy(), g[f"forward.{k}"]), k
    for k in ("rgb0", "rgb1", "depth0", "depth1", "opacity0", "opacity1"):
        assert rel_l2(out[k], g[f"forward.{k}"]) < 1e-6, k
    co = orender.render_forward(c["sd"], c["cfg"], scenes.NEAR, scenes.FAR, c["particles"], c["ro"], c["rays"],
                                mode="coarse")
    assert rel_l2(co["rgb0"], g["coarse.rgb0"]) < 1e-6
    assert "rgb1" not in co


def test_render_oracle_jitter_matches_reference_golden():
    """perturb > 0 / noise_std > 0 (utils/ray_utils.py:245-253, 186-190; models/renderer.py:192-196): the golden is the
    reference's own forward under a fixed seed; its four random draws, replayed in the reference's call order, are stored
    with it -- the oracle fed those numbers must reproduce the reference's outputs."""
    c = load_render_case("small_jitter")
    g = c["g"]
    jit = {k: torch.from_numpy(g["draw." + k]) for k in ("z_rand", "noise0", "u", "noise1")}
    out = orender.render_forward(c["sd"], c["cfg"], scenes.NEAR, scenes.FAR, c["particles"], c["ro"], c["rays"],
                                 perturb=float(g["perturb"]), noise_std=float(g["noise_std"]), jitter=jit)
    for k in ("num_nn_0", "num_nn_1"):
        assert np.array_equal(out[k].numpy().astype(np.int8), g[f"forward.{k}"]), k
    for k in ("rgb0", "rgb1", "depth0", "depth1", "opacity0", "opacity1"):
        assert rel_l2(out[k], g[f"forward.{k}"]) < 1e-6, k
    # and the jitter matters: the deterministic forward is something else
    det = orender.render_forward(c["sd"], c["cfg"], scenes.NEAR, scenes.FAR, c["particles"], c["ro"], c["rays"])
    assert rel_l2(det["rgb1"], g["forward.rgb1"]) > 1e-2


def test_render_oracle_fine_mode_self_consistent():
    # reference fine_rendering is broken on shipped configs (see make_golden.py); the restated intent
    # must equal the full forward when the coarse pass only contributes sigma (it always does).
    c = load_render_case("small_boost")
    full = orender.render_forward(c["sd"], c["cfg"], scenes.NEAR, scenes.FAR, c["particles"], c["ro"], c["rays"])
    fine = orender.render_forward(c["sd"], c["cfg"], scenes.NEAR, scenes.FAR, c["particles"], c["ro"], c["rays"],
                                  mode="fine")
    assert "rgb0" not in fine
    assert rel_l2(fine["rgb1"], full["rgb1"]) < 1e-6


@pytest.mark.parametrize("name", TRANSITION_CASES)
def test_transition_oracle_matches_reference_golden(name):
    c = load_transition_case(name)
    g = c["g"]
    pos, vel = c["pos"], c["vel"]
    for s in range(int(g["steps"])):
        pos, vel, nn, dbg = otrans.particle_step(c["sd"], pos, vel, c["box"], c["box_n"], debug=True)
        assert np.array_equal(nn.numpy().astype(np.int16), g[f"nnbr_{s}"])
        if s == 0:
            assert rel_l2(dbg["feats"][0], g["feats0"]) < 1e-5
            assert rel_l2(dbg["feats"][-1] / 128, g["delta0"]) < 1e-4
        assert rel_l2(pos, g[f"pos_{s}"]) < 1e-6
        assert rel_l2(vel, g[f"vel_{s}"]) < 1e-4


def test_ball_query_c_vs_torch_and_edge_cases():
    rng = np.random.RandomState(0)
    p = torch.from_numpy(rng.uniform(-0.5, 0.5, (700, 3)).astype(np.float32))
    q = torch.from_numpy(rng.uniform(-0.7, 0.7, (300, 3)).astype(np.float32))
    q[0] = p[5]                                   # zero-distance neighbour (counts as padding downstream)
    for K, r in ((20, 0.225), (1, 0.1), (32, 0.4)):
        d_c, i_c = tpo.ball_query_shared(q, p, K, r)
        d_t, i_t = tpo.ball_query_shared_torch(q, p, K, r)
        assert torch.equal(i_c, i_t)
        assert torch.equal(d_c, d_t)
    # empty particle set and empty query set
    d, i = tpo.ball_query_shared(q, torch.zeros(0, 3), 20, 0.2)
    assert (i == -1).all() and (d == 0).all()
    d, i = tpo.ball_query_shared(torch.zeros(0, 3), p, 20, 0.2)
    assert i.shape == (0, 20)
    # first-K-by-index, not K-nearest: with everything in range the answer is 0..K-1
    d, i = tpo.ball_query_shared(torch.zeros(1, 3), p * 0.01, 20, 1.0)
    assert i[0].tolist() == list(range(20))


def test_radius_search_and_cconv_c_vs_torch():
    rng = np.random.RandomState(1)
    ip = torch.from_numpy(rng.uniform(-0.3, 0.3, (400, 3)).astype(np.float32))
    op = torch.cat([ip[:100], torch.from_numpy(rng.uniform(-0.3, 0.3, (50, 3)).astype(np.float32))])
    nb_c, rs_c, d_c = tpo.radius_search(ip, op, 0.1125, True)
    nb_t, rs_t, d_t = tpo.radius_search_torch(ip, op, 0.1125, True)
    assert torch.equal(rs_c, rs_t) and torch.equal(nb_c, nb_t) and torch.equal(d_c, d_t)
    feat = torch.from_numpy(rng.randn(400, 5).astype(np.float32))
    kern = torch.from_numpy(rng.uniform(-0.05, 0.05, (4, 4, 4, 5, 7)).astype(np.float32))
    bias = torch.from_numpy(rng.randn(7).astype(np.float32))
    win = lambda x: torch.clamp((1 - x) ** 3, 0, 1)
    o_c, cnt = tpo.cconv_forward(feat, ip, op, 0.225, kern, bias, torch.zeros(3))
    o_t, _ = tpo.cconv_forward_torch(feat, ip, op, 0.225, kern, bias, torch.zeros(3), True, win)
    assert torch.equal(cnt, rs_c[1:] - rs_c[:-1])
    assert rel_l2(o_c, o_t) < 1e-5
    assert tpo.reduce_subarrays_sum(torch.ones(int(rs_c[-1])), rs_c).tolist() == cnt.float().tolist()


def test_filter_geometry_properties():
    # the volume-preserving map sends the unit ball into the cube [-1,1]^3 and the sphere to its surface
    rng = np.random.RandomState(2)
    v = rng.randn(2000, 3)
    v = v / np.linalg.norm(v, axis=1, keepdims=True) * rng.uniform(0, 1, (2000, 1)) ** (1 / 3)
    c = tpo.ball_to_cube_volume_preserving(torch.from_numpy(v.astype(np.float32)))
    assert float(c.abs().max()) <= 1 + 1e-5
    s = tpo.ball_to_cube_volume_preserving(torch.from_numpy((v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)))
    assert torch.allclose(s.abs().max(dim=1).values, torch.ones(2000), atol=1e-4)
    cells, w = tpo.filter_corners_torch(torch.from_numpy(v.astype(np.float32)) * 0.1125, 1 / 0.1125, 4, torch.zeros(3))
    assert torch.allclose(w.sum(1), torch.ones(2000), atol=1e-5)
    assert int(cells.min()) >= 0 and int(cells.max()) < 64
    assert tpo.ball_to_cube_volume_preserving(torch.zeros(1, 3)).abs().sum() == 0
