"""Deterministic synthetic {map, scan, particle set} generators.

Used by the parity tests and bench.py.  Everything is numpy with `default_rng(seed)`; nothing
here touches the GPU.  Shapes follow the reference's data: map = voxel-filtered cloud with at
most one point per `map_downsample` voxel in pcl::VoxelGrid order (src/mcl_3dl.cpp:1155-1158),
scan = points in the base frame inside the model's clip window
(src/lidar_measurement_model_likelihood.cpp:79-103), particles = State6DOF pos_/rot_
(include/mcl_3dl/state_6dof.h:55-56).
"""
import math

import numpy as np

POINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("label", "<u4")])
POSE = np.dtype([("px", "<f4"), ("py", "<f4"), ("pz", "<f4"), ("_pad", "<f4"),
                 ("qx", "<f4"), ("qy", "<f4"), ("qz", "<f4"), ("qw", "<f4")])
# mcl3dl_state: the State6DOF fields the resident particle set keeps (include/mcl3dl_b200.h), 68 bytes
STATE = np.dtype([("pos", "<f4", 3), ("rot", "<f4", 4), ("noise_ll", "<f4"), ("noise_la", "<f4"), ("noise_al", "<f4"),
                  ("noise_aa", "<f4"), ("odom_err_integ_lin", "<f4", 3), ("odom_err_integ_ang", "<f4", 3)])
RESULT = np.dtype([("score_like", "<f4"), ("match_cnt", "<u4"), ("score_beam", "<f4"),
                   ("n_short", "<u4"), ("n_hit", "<u4"), ("n_long", "<u4")])


def make_points(xyz, label=None):
    xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
    out = np.zeros(len(xyz), dtype=POINT)
    out["x"], out["y"], out["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    if label is not None:
        out["label"] = np.asarray(label, dtype=np.uint32)
    return out


def make_poses(pos, quat):
    pos = np.asarray(pos, dtype=np.float32).reshape(-1, 3)
    quat = np.asarray(quat, dtype=np.float32).reshape(-1, 4)
    out = np.zeros(len(pos), dtype=POSE)
    out["px"], out["py"], out["pz"] = pos[:, 0], pos[:, 1], pos[:, 2]
    out["qx"], out["qy"], out["qz"], out["qw"] = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    return out


def quat_from_rpy(rpy):
    """Quat::setRPY (include/mcl_3dl/quat.h:194-209), vectorised; returns (n,4) xyzw."""
    rpy = np.asarray(rpy, dtype=np.float64).reshape(-1, 3)
    t2, t3 = np.cos(rpy[:, 0] / 2), np.sin(rpy[:, 0] / 2)
    t4, t5 = np.cos(rpy[:, 1] / 2), np.sin(rpy[:, 1] / 2)
    t0, t1 = np.cos(rpy[:, 2] / 2), np.sin(rpy[:, 2] / 2)
    x = t0 * t3 * t4 - t1 * t2 * t5
    y = t0 * t2 * t5 + t1 * t3 * t4
    z = t1 * t2 * t4 - t0 * t3 * t5
    w = t0 * t2 * t4 + t1 * t3 * t5
    return np.stack([x, y, z, w], axis=1).astype(np.float32)


def rot_matrix(q):
    x, y, z, w = [float(v) for v in q]
    n = math.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _plane(rng, u0, u1, v0, v1, h, fixed_axis, fixed_val, sigma):
    """One jittered sample per h-cell of a rectangle; the normal coordinate gets N(0, sigma)."""
    nu = max(1, int(round((u1 - u0) / h)))
    nv = max(1, int(round((v1 - v0) / h)))
    uu, vv = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    u = u0 + (uu.ravel() + 0.5 + rng.uniform(-0.3, 0.3, nu * nv)) * h
    v = v0 + (vv.ravel() + 0.5 + rng.uniform(-0.3, 0.3, nu * nv)) * h
    w = fixed_val + rng.normal(0.0, sigma, nu * nv)
    cols = [None, None, None]
    others = [a for a in range(3) if a != fixed_axis]
    cols[fixed_axis] = w
    cols[others[0]] = u
    cols[others[1]] = v
    return np.stack(cols, axis=1)


def warehouse_map(n_target, height=None, seed=1, h=0.1, sigma=0.01, box_pitch=10.0, box_size=2.0,
                  labelled_wall=True):
    """Floor + ceiling + perimeter walls + a grid of box obstacles, one point per h-voxel.

    The footprint L x L x H is solved from n_target (points ~ surface area / h^2).  Returns
    (points[POINT], info dict).  One perimeter wall carries label 2 (the reference's label-filter
    path, test/src/test_beam_label.cpp:49-105).
    """
    rng = np.random.default_rng(seed)
    H = height if height is not None else (3.0 if n_target < 200_000 else 10.0)
    per_box = 5.0 * box_size * box_size
    # area(L) = 2 L^2 + 4 L H + (L/box_pitch)^2 * per_box  ==  n_target * h^2
    a = 2.0 + per_box / (box_pitch * box_pitch)
    b = 4.0 * H
    c = -n_target * h * h
    L = (-b + math.sqrt(b * b - 4 * a * c)) / (2 * a)
    L = max(4.0, round(L / h) * h)
    off = 0.03  # keep planes off voxel faces so the z-noise does not straddle two voxels
    parts, labels = [], []

    def add(p, lab=0):
        parts.append(p)
        labels.append(np.full(len(p), lab, dtype=np.uint32))

    add(_plane(rng, 0, L, 0, L, h, 2, off, sigma))          # floor
    add(_plane(rng, 0, L, 0, L, h, 2, H - off, sigma))      # ceiling
    add(_plane(rng, 0, L, 0, H, h, 0, off, sigma))          # wall x=0   (y,z)
    add(_plane(rng, 0, L, 0, H, h, 0, L - off, sigma), 2 if labelled_wall else 0)
    add(_plane(rng, 0, L, 0, H, h, 1, off, sigma))          # wall y=0   (x,z)
    add(_plane(rng, 0, L, 0, H, h, 1, L - off, sigma))
    nb = int(L // box_pitch)
    for i in range(nb):
        for j in range(nb):
            cx, cy = (i + 0.5) * box_pitch + 0.04, (j + 0.5) * box_pitch + 0.04
            x0, x1, y0, y1 = cx - box_size / 2, cx + box_size / 2, cy - box_size / 2, cy + box_size / 2
            add(_plane(rng, y0, y1, 0, box_size, h, 0, x0, sigma))
            add(_plane(rng, y0, y1, 0, box_size, h, 0, x1, sigma))
            add(_plane(rng, x0, x1, 0, box_size, h, 1, y0, sigma))
            add(_plane(rng, x0, x1, 0, box_size, h, 1, y1, sigma))
            add(_plane(rng, x0, x1, y0, y1, h, 2, box_size + off, sigma))
    xyz = np.concatenate(parts).astype(np.float32)
    lab = np.concatenate(labels)
    # pcl::VoxelGrid: at most one point per voxel, output sorted by voxel index (x fastest, z slowest)
    ijk = np.floor(xyz.astype(np.float64) / h).astype(np.int64)
    ijk -= ijk.min(axis=0)
    dims = ijk.max(axis=0) + 1
    key = ijk[:, 0] + dims[0] * (ijk[:, 1] + dims[1] * ijk[:, 2])
    _, first = np.unique(key, return_index=True)
    xyz, lab = xyz[first], lab[first]
    info = {"L": float(L), "H": float(H), "n": int(len(xyz)), "boxes": nb * nb, "box_pitch": box_pitch,
            "box_size": box_size, "h": h}
    return make_points(xyz, lab), info


def free_space_position(info, rng, z=0.6):
    """A position in the open aisle between obstacles."""
    L, pitch = info["L"], info["box_pitch"]
    nb = max(1, int(L // pitch))
    i, j = rng.integers(0, nb), rng.integers(0, nb)
    # aisle crossings sit on multiples of the pitch (boxes are centred at (i+0.5)*pitch)
    x = min(max(i * pitch + rng.uniform(-1.0, 1.0), 1.5), L - 1.5)
    y = min(max(j * pitch + rng.uniform(-1.0, 1.0), 1.5), L - 1.5)
    return np.array([x, y, z])


def make_scan(map_pts, truth_pos, truth_quat, n, clip_near, clip_far, clip_z=(-2.0, 2.0), sigma=0.02,
              n_origins=1, seed=2):
    """n map-surface points inside the clip window around the truth pose, + noise, in the base frame.

    Same window as the models' filter(): planar range in [clip_near, clip_far], z in clip_z
    (src/lidar_measurement_model_likelihood.cpp:83-93).  label = sensor index (< n_origins).
    """
    rng = np.random.default_rng(seed)
    xyz = np.stack([map_pts["x"], map_pts["y"], map_pts["z"]], axis=1).astype(np.float64)
    pre = np.abs(xyz[:, 0] - truth_pos[0]) <= clip_far
    pre &= np.abs(xyz[:, 1] - truth_pos[1]) <= clip_far
    cand = xyz[pre]
    R = rot_matrix(truth_quat)
    local = (cand - np.asarray(truth_pos, dtype=np.float64)) @ R  # R^T (p - t)
    r2 = local[:, 0] ** 2 + local[:, 1] ** 2
    ok = (r2 <= clip_far ** 2) & (r2 >= clip_near ** 2) & (local[:, 2] >= clip_z[0]) & (local[:, 2] <= clip_z[1])
    local = local[ok]
    if len(local) == 0:
        raise ValueError("no map points inside the clip window")
    idx = rng.choice(len(local), size=n, replace=len(local) < n)
    pts = local[idx] + rng.normal(0.0, sigma, (n, 3))
    labels = rng.integers(0, n_origins, n) if n_origins > 1 else np.zeros(n, dtype=np.uint32)
    return make_points(pts.astype(np.float32), labels)


def sensor_origins(n_origins, seed=4):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-0.3, 0.3, (n_origins, 3))
    o[:, 2] = np.abs(o[:, 2])
    o[0] = (0.0, 0.0, 0.2)
    return o.astype(np.float32)


def tracking_particles(n, truth_pos, truth_rpy, seed=3, sigma_pos=(0.2, 0.2, 0.05), sigma_rpy=(0.02, 0.02, 0.1),
                       quat_scale_jitter=True):
    """pos ~ N(truth, sigma_pos^2), rpy ~ N(truth, sigma_rpy^2)."""
    rng = np.random.default_rng(seed)
    pos = np.asarray(truth_pos) + rng.normal(0, 1, (n, 3)) * np.asarray(sigma_pos)
    rpy = np.asarray(truth_rpy) + rng.normal(0, 1, (n, 3)) * np.asarray(sigma_rpy)
    q = quat_from_rpy(rpy)
    if quat_scale_jitter:  # rot_ is not exactly unit in the node (float drift); keep that path exercised
        q = q * (1.0 + rng.normal(0, 1e-3, (n, 1))).astype(np.float32)
    return make_poses(pos, q)


def spread_particles(n, info, seed=3, z=0.6, div_yaw=12):
    """Uniform over the floor plan x div_yaw headings: the global-localisation layout
    (src/mcl_3dl.cpp:1076-1095) and the HBM-bound case (SURVEY hard part 7)."""
    rng = np.random.default_rng(seed)
    L = info["L"]
    pos = np.stack([rng.uniform(0.5, L - 0.5, n), rng.uniform(0.5, L - 0.5, n),
                    z + rng.normal(0, 0.05, n)], axis=1)
    yaw = (rng.integers(0, div_yaw, n) * (2 * math.pi / div_yaw))
    rpy = np.stack([np.zeros(n), np.zeros(n), yaw], axis=1)
    return make_poses(pos, quat_from_rpy(rpy))


def scene(n_map, n_particles, n_lik, n_beam, spread=False, n_origins=2, seed=0):
    """A full {map, particles, likelihood scan, beam scan, origins} bundle."""
    mp, info = warehouse_map(n_map, seed=seed + 1)
    rng = np.random.default_rng(seed + 10)
    truth_pos = free_space_position(info, rng)
    truth_rpy = np.array([0.0, 0.0, rng.uniform(-math.pi, math.pi)])
    truth_q = quat_from_rpy(truth_rpy)[0]
    lik = make_scan(mp, truth_pos, truth_q, n_lik, 0.5, 10.0, seed=seed + 2) if n_lik else np.zeros(0, POINT)
    beam = (make_scan(mp, truth_pos, truth_q, n_beam, 0.5, 4.0, n_origins=n_origins, seed=seed + 5)
            if n_beam else np.zeros(0, POINT))
    origins = sensor_origins(n_origins, seed=seed + 4)
    if spread:
        particles = spread_particles(n_particles, info, seed=seed + 3)
    else:
        particles = tracking_particles(n_particles, truth_pos, truth_rpy, seed=seed + 3)
    return {"map": mp, "info": info, "particles": particles, "lik": lik, "beam": beam, "origins": origins,
            "truth_pos": truth_pos, "truth_rpy": truth_rpy}
