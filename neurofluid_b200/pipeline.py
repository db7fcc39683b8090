"""Caller-side pieces around the two hot paths, kept on the device.

Host mirrors of (reference file:line):
  BaseTrainer.render_image                     trainer/basetrainer.py:264-309   -> render_image
  BaseTrainer.load_pretained_transition_model  trainer/basetrainer.py:87-103    -> load_pretrained_transition_model
  BaseTrainer.load_pretained_renderer_model    trainer/basetrainer.py:106-122   -> load_pretrained_renderer_model
  Evaluator.resume                             eval_e2e.py:50-55                -> resume
  FluidErrors                                  utils/point_eval.py:31-60        -> FluidErrors (device statistics)
  Evaluator.eval loop                          eval_e2e.py:58-120               -> rollout_and_render
The arithmetic (ray generation, nearest-neighbour distances, MSE) runs in csrc/nf_aux.cu through the C ABI.
"""
from __future__ import annotations

import json

import torch

from . import ops
from ._lib import NFError
from .distributed import gather_image, shard_rows


# ------------------------------------------------------------------------------------------------ render_image
def render_image(renderer, particle_pos, N_ray, ro, rays, focal_length, cw, ray_chunk=None, iseval=False):
    """trainer/basetrainer.py:264-309.  The reference walks the image in cfg.ray.ray_chunk = 1024-ray pieces because its
    O(rays x particles) `repeat` does not fit otherwise; here any chunk size gives bit-identical results (rays are
    independent; tests/test_gpu_pipeline.py), so the default is ONE call for the whole image (the renderer cuts it into
    `max_rays_per_launch` launches itself: 5 launches per 800 x 800 image instead of 625 calls).  Pass `ray_chunk` to get the
    reference's loop."""
    chunk = int(ray_chunk) if ray_chunk else max(int(N_ray), 1)
    fine = renderer.cfg.ray.N_importance > 0
    acc = {k: [] for k in ("pred_rgbs_0", "num_nn_0", "mask_0", "pred_rgbs_1", "num_nn_1", "mask_1")}
    for ray_idx in range(0, N_ray, chunk):
        r = renderer(particle_pos, ro, rays[ray_idx:ray_idx + chunk], focal_length, cw)
        acc["pred_rgbs_0"].append(r["rgb0"])
        if "num_nn_0" in r:                      # absent when the renderer was told to skip them (return_num_nn=False)
            acc["num_nn_0"].append(r["num_nn_0"].view(-1))
        if iseval:
            acc["mask_0"].append(r["mask_0"])
        if fine:
            acc["pred_rgbs_1"].append(r["rgb1"])
            if "num_nn_1" in r:
                acc["num_nn_1"].append(r["num_nn_1"].view(-1))
            if iseval:
                acc["mask_1"].append(r["mask_1"])
    ret = {"pred_rgbs_0": torch.cat(acc["pred_rgbs_0"], 0)}
    if acc["num_nn_0"]:
        ret["num_nn_0"] = torch.cat(acc["num_nn_0"], 0)
    if iseval:
        ret["mask_0"] = torch.cat(acc["mask_0"], 0)
    if fine:
        ret["pred_rgbs_1"] = torch.cat(acc["pred_rgbs_1"], 0)
        if acc["num_nn_1"]:
            ret["num_nn_1"] = torch.cat(acc["num_nn_1"], 0)
        if iseval:
            ret["mask_1"] = torch.cat(acc["mask_1"], 0)
    return ret


# ------------------------------------------------------------------------------------------------ checkpoints
def _load(path_or_dict, map_location=None):
    if isinstance(path_or_dict, dict):
        return path_or_dict
    return torch.load(path_or_dict, map_location=map_location, weights_only=True)


def load_pretrained_transition_model(transition_model, ckpt, map_location=None):
    """trainer/basetrainer.py:87-103: accepts 'transition_model_state_dict' / 'model_state_dict' / a bare state
    dict, drops `gravity`, loads strictly."""
    ckpt = _load(ckpt, map_location)
    if "transition_model_state_dict" in ckpt:
        ckpt = ckpt["transition_model_state_dict"]
    elif "model_state_dict" in ckpt:
        ckpt = ckpt["model_state_dict"]
    ckpt = {k: v for k, v in ckpt.items() if "gravity" not in k}
    sd = transition_model.state_dict()
    sd.update(ckpt)
    transition_model.load_state_dict(sd, strict=True)
    return transition_model


def load_pretrained_renderer_model(renderer, ckpt, partial_load=False, map_location=None):
    """trainer/basetrainer.py:106-122: 'renderer_state_dict'; partial_load keeps only sigma / xyz_encoding layers."""
    ckpt = _load(ckpt, map_location)["renderer_state_dict"]
    if partial_load:
        ckpt = {k: v for k, v in ckpt.items() if "sigma" in k or "xyz_encoding" in k}
    sd = renderer.state_dict()
    sd.update(ckpt)
    renderer.load_state_dict(sd, strict=True)
    return renderer


def resume(renderer, transition_model, ckpt, map_location=None):
    """eval_e2e.py:50-55."""
    ckpt = _load(ckpt, map_location)
    renderer.load_state_dict(ckpt["renderer_state_dict"], strict=True)
    transition_model.load_state_dict(ckpt["transition_model_state_dict"], strict=True)


# ------------------------------------------------------------------------------------------------ metrics
def _stats(x: torch.Tensor) -> torch.Tensor:
    """utils/point_eval.py:17-28 as one (6,) device tensor [mean, mse, var, min, max, median] * 1000."""
    s, _ = torch.sort(x)
    n = s.shape[0]
    median = 0.5 * (s[(n - 1) // 2] + s[n // 2])                 # np.median: mean of the two middle values
    xd = x.double()
    return torch.stack([xd.mean(), (xd * xd).mean(), xd.var(unbiased=False), xd.min(), xd.max(), median.double()]) * 1000.0


_STAT_KEYS = ("mean", "mse", "var", "min", "max", "median")


class FluidErrors:
    """utils/point_eval.py:31-83 with the statistics left on the device: `cal_errors` launches the nearest-neighbour
    kernel and returns the gt->pred mean distance (x1000) as a 0-dim tensor without synchronising; `errors` /
    `save` materialise floats (one sync for the whole rollout instead of a `.cpu().numpy()` per frame,
    trainer/trainer_e2e.py:334)."""

    def __init__(self, cell: float = 0.1):
        self._dev = {}
        self.cell = float(cell)

    def cal_errors(self, pred_pos, gt_pos, time_idx):
        ops.require_cuda(pred_pos, gt_pos)
        pred = pred_pos.detach().to(torch.float32).reshape(-1, 3)
        gt = gt_pos.detach().to(torch.float32).reshape(-1, 3)
        # the reference prints and returns None on non-finite input (utils/point_eval.py:37-42); here the check is
        # folded into the result (NaN statistics) to keep the call asynchronous
        a = _stats(ops.pair_distance(pred, gt))
        b = _stats(ops.nearest_distance(gt, pred, cell=self.cell))
        self._dev[time_idx] = (a, b, pred.shape[0], gt.shape[0])
        return b[0]

    @property
    def errors(self):
        out = {}
        for t, (a, b, n_a, n_b) in self._dev.items():
            e = {k: float(v) for k, v in zip(_STAT_KEYS, a.tolist())}
            e["num_particles"] = n_a
            for k, v in zip(_STAT_KEYS, b.tolist()):
                e["gt2pred_" + k] = float(v)
            e["gt2pred_num_particles"] = n_b
            out[t] = e
        return out

    def save(self, path):
        with open(path, "w") as f:
            json.dump(list(self.errors.items()), f, indent=4)


# ------------------------------------------------------------------------------------------------ rollout
@torch.no_grad()
def rollout_and_render(transition_model, renderer, pos, vel, box, box_normals, cameras, H, W, n_frames,
                       gt_positions=None, gt_images=None, group=None, keep_images=False, transition="replicated"):
    """The loop of eval_e2e.py:58-120 without the dataset: for every frame one transition step, the position
    metrics, and for every camera `(c2w (3,4), focal)` one rendered image (+ PSNR against `gt_images[f][v]`).

    With torch.distributed initialised the rays of every image are sharded block-cyclically by row over the ranks
    of `group`; the transition model runs replicated (`transition="replicated"`, the throughput-optimal layout of
    SURVEY.md section 8e) or particle-block sharded with an NCCL all-gather per layer and of the positions per step
    (`transition="sharded"`, BASELINE.json's north_star layout; bit-identical results).  The returned images are
    the local rows unless `keep_images`, which all-gathers them.  The loop issues no host synchronisation (the camera
    position is read from device memory, the particle grid and workspaces are cached); the neighbour-overflow check
    after the last frame is the one sync of the rollout.  The per-sample neighbour counts (`num_nn_*`, which the
    reference's eval loop never reads) are not materialised here."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if len(cameras) == 0:
        raise NFError("rollout_and_render: no cameras")
    if transition not in ("replicated", "sharded"):
        raise NFError(f"rollout_and_render: transition={transition!r}")
    from .distributed import transition_step_sharded
    fe = FluidErrors()
    cams = [(torch.as_tensor(c2w, dtype=torch.float32, device=pos.device), float(f)) for c2w, f in cameras]
    # rays of a fixed camera do not change between frames: generate once on the device, keep this rank's rows
    rays = [shard_rows(ops.generate_rays(H, W, f, c2w).view(H, W, 6), rank, world).reshape(-1, 6).contiguous()
            for c2w, f in cams]
    frames, psnr, pos_hist = [], [], []
    want_nn, renderer.return_num_nn = getattr(renderer, "return_num_nn", True), False
    for fidx in range(n_frames):
        if transition == "sharded" and world > 1:
            pos, vel, _ = transition_step_sharded(transition_model, pos, vel, box, box_normals, group)
        else:
            pos, vel, _ = transition_model(pos, vel, box, box_normals)
        pos, vel = pos.clone(), vel.clone()                      # eval_e2e.py:83
        if gt_positions is not None:
            fe.cal_errors(pos, gt_positions[fidx], fidx + 1)
        views = []
        for v, (c2w, f) in enumerate(cams):
            ro = renderer.set_ro(c2w)
            r = render_image(renderer, pos, rays[v].shape[0], ro, rays[v], f, c2w, ray_chunk=max(rays[v].shape[0], 1),
                             iseval=True)
            img = r["pred_rgbs_1"] if "pred_rgbs_1" in r else r["pred_rgbs_0"]
            if gt_images is not None:
                gt = shard_rows(gt_images[fidx][v].view(H, W, 3), rank, world).reshape(-1, 3)
                psnr.append(ops.img2mse(img, gt))                # per-rank partial MSE; reduced below
            if keep_images:
                img = gather_image(img.view(-1, W, 3), H, group)
            views.append(img)
        frames.append(views)
        pos_hist.append(pos)
    renderer.return_num_nn = want_nn
    if hasattr(transition_model, "check_neighbor_overflow"):
        transition_model.check_neighbor_overflow(group)          # the one host sync of the rollout (all ranks decide together)
    out = {"positions": pos_hist, "images": frames, "fluid_errors": fe}
    if gt_images is not None:
        mse = torch.stack(psnr).view(n_frames, len(cams))
        if world > 1:                                            # mean over equally weighted pixels of all ranks
            n_loc = torch.tensor([float(rays[0].shape[0])], device=pos.device)
            tot = mse * n_loc
            dist.all_reduce(tot, group=group)
            dist.all_reduce(n_loc, group=group)
            mse = tot / n_loc
        out["psnr"] = ops.mse2psnr(mse)
    return out
