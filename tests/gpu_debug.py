"""Stage-by-stage diagnostics on a GPU box (not a pytest): prints error metrics for every kernel so
one gpurun call localises a failure.   python tests/gpu_debug.py [stage ...]"""
import os
import sys
import time

import numpy as np
import torch
torch.set_grad_enabled(False)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_render_case, rel_l2  # noqa: E402
from neurofluid_b200 import ops, scenes, _lib  # noqa: E402
import neurofluid_b200 as nb  # noqa: E402
from oracle import renderer as orender  # noqa: E402
from oracle import third_party_ops as tpo  # noqa: E402

dev = torch.device("cuda:0")


def stage_grid():
    rng = np.random.RandomState(0)
    p = torch.from_numpy(scenes.lattice_particles(12, 0))
    q = torch.from_numpy(rng.uniform(-0.6, 0.6, (5000, 3)).astype(np.float32))
    q[:50] = p[:50]
    for K, r in ((20, 0.225), (32, 0.1), (5, 0.3)):
        d2o, io = tpo.ball_query_shared(q, p, K, r)
        d2, idx, nn = ops.ball_query(q.to(dev), p.to(dev), K, r)
        torch.cuda.synchronize()
        same = torch.equal(idx.cpu(), io)
        print(f"[grid] K={K} r={r}: idx equal={same} mism rows={(idx.cpu() != io).any(1).sum().item()} "
              f"d2 equal={torch.equal(d2.cpu(), d2o)}")


def mlp_reference(sd, net, rec):
    x, dens, sm, var, d, sd_ = rec[:, 0:3], rec[:, 3:4], rec[:, 4:7], rec[:, 7:10], rec[:, 10:13], rec[:, 13:16]
    pe = orender.positional_encoding
    feats = torch.cat([pe(x, 10), pe(dens, 4), pe(sm, 10), pe(var, 10), pe(d, 4), pe(sd_, 4)], 1)
    return orender.nerf_mlp(sd, net, feats, 198, 54), orender.nerf_mlp(sd, net, feats, 198, 54, sigma_only=True)


def stage_mlp():
    rng = np.random.RandomState(1)
    for gain, n in ((2.45, 1000), (1.0, 128 * 148 * 2 + 77)):
        sd = scenes.init_render_state(7, weight_gain=gain)
        net = nb.RenderNet(scenes.render_cfg(), 9.0, 13.0)
        net.load_state_dict(sd)
        net = net.to(dev)
        rec = np.concatenate([rng.uniform(-1.5, 1.5, (n, 3)), rng.uniform(0, 12, (n, 1)), rng.uniform(-1.5, 1.5, (n, 3)),
                              rng.uniform(0, 0.02, (n, 3)), rng.randn(n, 6)], 1).astype(np.float32)
        rec[:, 10:13] /= np.linalg.norm(rec[:, 10:13], axis=1, keepdims=True)
        rec[:, 13:16] /= np.linalg.norm(rec[:, 13:16], axis=1, keepdims=True)
        rec = torch.from_numpy(rec)
        ref, ref_sig = mlp_reference(sd, "nerf_coarse", rec)
        for swap in (0,):
            os.environ["NF_MLP_DESC_SWAP"] = str(swap)
            for dt, name in ((_lib.NF_DTYPE_F16, "fp16"), (_lib.NF_DTYPE_BF16, "bf16")):
                packed = ops.pack_nerf_weights([p.to(dev) for p in net.nerf_coarse.ordered_params()], dt)
                t0 = time.time()
                out = ops.nerf_mlp(packed, rec.to(dev), dt)
                torch.cuda.synchronize()
                o = out.cpu()
                print(f"[mlp] gain={gain} n={n} swap={swap} {name}: rgb rel={rel_l2(o[:, :3], ref[:, :3]):.3e} "
                      f"sigma rel={rel_l2(o[:, 3], ref[:, 3]):.3e} max|sig|={ref[:,3].abs().max():.3f} "
                      f"nan={torch.isnan(o).sum().item()} ({time.time() - t0:.3f}s)")
                if swap == 0 and name == "fp16":
                    so = ops.nerf_mlp(packed, rec.to(dev), dt, sigma_only=True).cpu()
                    print(f"[mlp]   sigma_only: rel={rel_l2(so[:, 3], ref_sig[:, 0]):.3e} rgb zero={bool((so[:, :3] == 0).all())}")
        os.environ["NF_MLP_DESC_SWAP"] = "0"


def stage_render():
    for name in ("small_boost", "small_he", "small_nomask", "cfg0_sub", "small_default"):
        c = load_render_case(name)
        g = c["g"]
        net = nb.RenderNet(c["cfg"], scenes.NEAR, scenes.FAR)
        net.load_state_dict(c["sd"])
        net = net.to(dev)
        out = net(c["particles"].to(dev), c["ro"].to(dev), c["rays"].to(dev), 0.0, c["cw"].to(dev))
        torch.cuda.synchronize()
        msg = [f"[render] {name}:"]
        for k in ("num_nn_0", "num_nn_1"):
            msg.append(f"{k} mism={(out[k].cpu().numpy().astype(np.int8) != g['forward.' + k]).sum()}")
        for k in ("mask_0", "mask_1"):
            msg.append(f"{k} mism={(out[k].cpu().numpy() != g['forward.' + k]).sum()}")
        for k in ("rgb0", "depth0", "opacity0", "rgb1", "depth1", "opacity1"):
            msg.append(f"{k}={rel_l2(out[k].cpu(), g['forward.' + k]):.2e}")
        msg.append(f"stats={net.last_stats.sum(0).tolist()}")
        print(" ".join(msg))
        co = net.coarse_rendering(c["particles"].to(dev), c["ro"].to(dev), c["rays"].to(dev), 0.0, c["cw"].to(dev))
        fi = net.fine_rendering(c["particles"].to(dev), c["ro"].to(dev), c["rays"].to(dev), 0.0, c["cw"].to(dev))
        print(f"[render] {name}: coarse rgb0={rel_l2(co['rgb0'].cpu(), g['coarse.rgb0']):.2e} "
              f"fine-mode rgb1 vs forward rgb1={rel_l2(fi['rgb1'].cpu(), g['forward.rgb1']):.2e}")


def stage_bench():
    H = 800
    rays, focal, cw = scenes.camera_rays(H, H)
    particles = torch.from_numpy(scenes.lattice_particles(27, 0))
    net = nb.RenderNet(scenes.render_cfg(), scenes.NEAR, scenes.FAR)
    net.load_state_dict(scenes.init_render_state(0, 5.0))
    net = net.to(dev)
    rays_d, p_d, ro = rays.to(dev), particles.to(dev), cw[:, 3].to(dev)
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        out = net(p_d, ro, rays_d, focal, cw)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"[bench] 800x800 P={particles.shape[0]}: {dt * 1e3:.1f} ms -> {rays.shape[0] / dt / 1e6:.3f} Mrays/s "
              f"stats={net.last_stats.sum(0).tolist()} rgb1 mean={out['rgb1'].mean().item():.4f}")




def stage_transition():
    from helpers import load_transition_case
    from oracle import transition as otrans
    for name in ("small", "medium"):
        c = load_transition_case(name)
        g = c["g"]
        net = nb.ParticleNet(gravity=(0.0, 0.0, -9.81))
        net.load_state_dict(c["sd"])
        net = net.to(dev)
        pos, vel = c["pos"].to(dev), c["vel"].to(dev)
        box, box_n = c["box"].to(dev), c["box_n"].to(dev)
        for s in range(int(g["steps"])):
            dbg = {}
            pos, vel, nn = net(pos, vel, box, box_n, debug=dbg)
            torch.cuda.synchronize()
            msg = f"[trans] {name} step {s}: nnbr mism={(nn.cpu().numpy().astype(np.int16) != g[f'nnbr_{s}']).sum()} " \
                  f"pos={rel_l2(pos.cpu(), g[f'pos_{s}']):.2e} vel={rel_l2(vel.cpu(), g[f'vel_{s}']):.2e}"
            if s == 0:
                msg += f" feats0={rel_l2(dbg['feats0'].cpu(), g['feats0']):.2e} delta={rel_l2(net.pos_correction.cpu(), g['delta0']):.2e}"
            print(msg)
    # timing at the BASELINE config[2] size
    n = 31
    half = (n - 1) / 2 * 0.05
    pos = torch.from_numpy(scenes.lattice_particles(n, 0, center=(0.0, 0.0, -1 + 0.03 + half))).to(dev)
    vel = torch.zeros_like(pos)
    bp, bn = scenes.box_points(0.032)
    box, box_n = torch.from_numpy(bp).to(dev), torch.from_numpy(bn).to(dev)
    net = nb.ParticleNet(gravity=(0.0, 0.0, -9.81))
    net.load_state_dict(scenes.init_particle_state(0))
    net = net.to(dev)
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        p, v = pos, vel
        for s in range(10):
            p, v, nn = net(p, v, box, box_n)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / 10
        print(f"[trans] N={pos.shape[0]} M={box.shape[0]}: {dt * 1e3:.3f} ms/step -> {pos.shape[0] / dt / 1e6:.2f} M particle-steps/s "
              f"mean nbrs={nn.mean().item():.1f}")


if __name__ == "__main__":
    stages = sys.argv[1:] or ["grid", "mlp", "render", "bench"]
    for s in stages:
        try:
            globals()["stage_" + s]()
        except Exception as e:  # keep going: one call should tell us as much as possible
            import traceback
            traceback.print_exc()
            print(f"[{s}] FAILED: {e}")
            try:
                torch.cuda.synchronize()
            except Exception as e2:
                print("device error persists:", e2)
                break
