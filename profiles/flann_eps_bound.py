#!/usr/bin/env python
"""How far can a real PCL/FLANN build (kd-tree epsilon = map_grid_min / 16 = 0.00625, src/mcl_3dl.cpp:1328) move the
likelihood score away from the exact nearest-neighbour search this repo (and its oracle) implements?

FLANN's KDTreeSingleIndex prunes a branch when mindist^2 * (1 + eps) > worst, so with eps > 0 it may
  (a) return a neighbour whose squared distance is up to (1 + eps) x the true nearest one's, and
  (b) miss neighbours whose squared distance lies in (r^2 / (1 + eps), r^2].
Worst case for the score  sum_i (R - max(d_i, F)) * W  (src/lidar_measurement_model_likelihood.cpp:124-135): every
matched eval reports d_i * sqrt(1 + eps) and every eval in band (b) is dropped.  This script evaluates that bound per
particle on the BASELINE configs c1 and c2 (CPU only: scipy cKDTree over the rescaled map)."""
import json
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcl_3dl_b200 import synth  # noqa: E402

EPS = 0.1 / 16.0
R, F, W = 0.2, 0.05, 5.0
DW = np.array([1.0, 1.0, 5.0])


def bound(n_map, P, n_lik, seed):
    s = synth.scene(n_map, P, n_lik, 0, seed=seed)
    mp = np.stack([s["map"]["x"], s["map"]["y"], s["map"]["z"]], axis=1).astype(np.float64) * DW
    tree = cKDTree(mp)
    lik = np.stack([s["lik"]["x"], s["lik"]["y"], s["lik"]["z"]], axis=1).astype(np.float64)
    rel, dq = [], []
    for p in s["particles"]:
        Rm = synth.rot_matrix([p["qx"], p["qy"], p["qz"], p["qw"]])
        q = (lik @ Rm.T + np.array([p["px"], p["py"], p["pz"]])) * DW
        d, _ = tree.query(q, distance_upper_bound=R)
        hit = np.isfinite(d)
        exact = ((R - np.maximum(d[hit], F)) * W).sum()
        dw = d[hit] * np.sqrt(1.0 + EPS)
        keep = dw < R
        worst = ((R - np.maximum(dw[keep], F)) * W).sum()
        if exact > 0:
            rel.append((exact - worst) / exact)
        dq.append((hit.sum() - keep.sum()) / max(len(lik), 1))
    return {"particles": P, "points": n_lik, "score_rel_deviation_max": float(np.max(rel)), "score_rel_deviation_mean": float(np.mean(rel)),
            "quality_abs_deviation_max": float(np.max(dq))}


if __name__ == "__main__":
    out = {"eps": EPS, "c1": bound(50_000, 64, 96, 1000), "c2": bound(1_000_000, 256, 512, 1000)}
    print(json.dumps(out, indent=1))
