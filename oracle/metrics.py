"""CPU ORACLE (test infrastructure only -- never imported by neurofluid_b200/): the reference's evaluation
metrics restated line by line.

  _distance, _ground_truth_to_prediction_distance, _compute_stats, FluidErrors.cal_errors
                                                           utils/point_eval.py:7-8, 11-14, 17-28, 36-60
  img2mse, mse2psnr                                        trainer/trainer_e2e.py:24-25

scipy (cKDTree) is the same dependency the reference uses.  Pinned by tests/test_oracle.py against a brute-force
O(n^2) nearest-neighbour search; the reference ships no golden vectors for these functions.
"""
import numpy as np
from scipy.spatial import cKDTree


def pair_distance(x, y):                                    # utils/point_eval.py:7-8
    return np.linalg.norm(x - y, axis=-1)


def gt_to_pred_distance(pred, gt):                          # utils/point_eval.py:11-14
    tree = cKDTree(pred)
    dist, _ = tree.query(gt)
    return dist


def compute_stats(x):                                       # utils/point_eval.py:17-28
    tmp = {"mean": np.mean(x), "mse": np.mean(x ** 2), "var": np.var(x), "min": np.min(x), "max": np.max(x),
           "median": np.median(x)}
    tmp = {k: float(v) * 1000 for k, v in tmp.items()}
    tmp["num_particles"] = x.shape[0]
    return tmp


def fluid_errors(pred_pos, gt_pos):                         # utils/point_eval.py:36-60 (one time index)
    errs = compute_stats(pair_distance(pred_pos, gt_pos))
    for k, v in compute_stats(gt_to_pred_distance(pred_pos, gt_pos)).items():
        errs["gt2pred_" + k] = v
    return errs


def img2mse(x, y):                                          # trainer/trainer_e2e.py:24
    return float(np.mean((np.asarray(x, np.float64) - np.asarray(y, np.float64)) ** 2))


def mse2psnr(mse):                                          # trainer/trainer_e2e.py:25
    return float(-10.0 * np.log(mse) / np.log(10.0))
