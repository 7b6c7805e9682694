"""Scan preprocessing on the device (scope row f4, scan half): VoxelGrid downsample, the models' filter() clip and the
uniform sample (include/mcl3dl_b200.h: mcl3dl_scan_prepare).  The clip is exact (oracle: mcl3dl_cpu_filter_clip, pinned
to the reference build); the VoxelGrid is PCL's (third-party, absent) restated from its published algorithm and checked
against a numpy restatement of the same; the sample is checked for being what it claims (draws from the clipped cloud)."""
import numpy as np
import pytest

from mcl_3dl_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng_mod():
    from mcl_3dl_b200 import engine
    engine.load_library()
    return engine


def voxel_grid_numpy(pts, leaf):
    """pcl::VoxelGrid::applyFilter restated: float index arithmetic, ascending voxel index, centroid in input order."""
    xyz = np.stack([pts["x"], pts["y"], pts["z"]], axis=1).astype(np.float32)
    inv = (np.float32(1.0) / np.asarray(leaf, np.float32)).astype(np.float32)
    mn, mx = xyz.min(axis=0), xyz.max(axis=0)
    min_b = np.floor((mn * inv).astype(np.float32)).astype(np.int64)
    max_b = np.floor((mx * inv).astype(np.float32)).astype(np.int64)
    div = max_b - min_b + 1
    ijk = np.floor((xyz * inv).astype(np.float32)).astype(np.int64) - min_b
    key = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(key, kind="stable")
    out = []
    k_sorted = key[order]
    starts = np.flatnonzero(np.concatenate([[True], k_sorted[1:] != k_sorted[:-1]]))
    ends = np.concatenate([starts[1:], [len(order)]])
    for a, b in zip(starts, ends):
        idx = order[a:b]
        s = np.zeros(3, np.float32)
        for i in idx:
            s = (s + xyz[i]).astype(np.float32)
        c = (s / np.float32(len(idx))).astype(np.float32)
        labs, cnt = np.unique(pts["label"][idx], return_counts=True)
        out.append((c[0], c[1], c[2], labs[np.argmax(cnt)]))
    return np.array(out, dtype=synth.POINT)


@pytest.mark.parametrize("seed,n_raw,leaf", [(1, 30000, (0.1, 0.1, 0.1)), (2, 5000, (0.2, 0.1, 0.3)), (3, 777, (0.05, 0.05, 0.05))])
def test_scan_prepare(eng_mod, seed, n_raw, leaf):
    from oracle import cpu_checker as cc
    cc.build("port")
    port = cc.CpuChecker("port")
    rng = np.random.default_rng(seed)
    # a raw scan: points on a few planes around the sensor, two sensors (labels), some far / high points to be clipped
    xyz = np.concatenate([rng.uniform([-12, -12, -0.6], [12, 12, -0.55], (n_raw // 2, 3)),
                          rng.uniform([-6, 3.0, -3], [6, 3.05, 3], (n_raw - n_raw // 2, 3))]).astype(np.float32)
    raw = synth.make_points(xyz, rng.integers(0, 2, n_raw))
    e = eng_mod.Engine((0,))
    sp = eng_mod.ScanParams(downsample=leaf, lik_num_points=200, beam_num_points=17, seed=seed)
    info = e.scan_prepare(raw, sp)
    ds = e.scan_get(0)
    want_ds = voxel_grid_numpy(raw, leaf)
    assert info["n_raw"] == n_raw and info["n_downsampled"] == len(want_ds) == len(ds)
    assert np.array_equal(ds["label"], want_ds["label"])
    for f in ("x", "y", "z"):
        assert np.allclose(ds[f], want_ds[f], rtol=0, atol=1e-6), f
    # the clip against the oracle's filter() (the reference's code), order kept
    for which, clip in ((1, (0.5, 10.0, -2.0, 2.0)), (2, (0.5, 4.0, -2.0, 2.0))):
        keep = port.filter_clip(ds, *clip)
        got = e.scan_get(which)
        assert got.tobytes() == ds[keep.astype(bool)].tobytes()
    lik_clip, beam_clip = e.scan_get(1), e.scan_get(2)
    assert 0 < len(beam_clip) < len(lik_clip) < len(ds)
    # the sample: `num` draws, every one a point of the clipped cloud, spread over it; reproducible per seed
    for which, clipped, num in ((3, lik_clip, 200), (4, beam_clip, 17)):
        got = e.scan_get(which)
        assert len(got) == num
        pool = {p.tobytes() for p in clipped}
        assert all(p.tobytes() in pool for p in got)
    lik = e.scan_get(3)
    assert len({p.tobytes() for p in lik}) > 150                    # with replacement, but not a handful of points
    e.scan_prepare(raw, sp)
    assert e.scan_get(3).tobytes() == lik.tobytes()
    sp.seed = seed + 100
    e.scan_prepare(raw, sp)
    assert e.scan_get(3).tobytes() != lik.tobytes()
    # no downsampling / empty cloud / nothing survives the clip
    sp0 = eng_mod.ScanParams(downsample=(0, 0, 0), lik_num_points=10, beam_num_points=0)
    i0 = e.scan_prepare(raw, sp0)
    assert i0["n_downsampled"] == n_raw and i0["n_lik"] == 10 and i0["n_beam"] == 0
    assert e.scan_prepare(raw[:0], sp0) == {"n_raw": 0, "n_downsampled": 0, "n_lik_clipped": 0, "n_beam_clipped": 0, "n_lik": 0, "n_beam": 0}
    far = synth.make_points(np.full((50, 3), 100.0, np.float32))
    i1 = e.scan_prepare(far, sp0)
    assert i1["n_lik_clipped"] == 0 and i1["n_lik"] == 0
    e.close()


def test_prepared_scans_feed_the_resident_update(eng_mod):
    """mcl3dl_particles_measure_update_prepared == mcl3dl_particles_measure_update fed with the scans read back."""
    s = synth.scene(60_000, 500, 64, 8, seed=171)
    e = eng_mod.Engine((0,))
    e.set_map(s["map"], eng_mod.LikParams(dist_weight=(1, 1, 5)), eng_mod.beam_params_from_reference(num_points_default=8))
    rng = np.random.default_rng(5)
    # a raw cloud in the base frame: the scene's scans, repeated with jitter (so that the voxel grid has work to do)
    base = np.concatenate([s["lik"], s["beam"]])
    raw = np.tile(base, 40)
    for f in ("x", "y", "z"):
        raw[f] += rng.normal(0, 0.03, len(raw)).astype(np.float32)
    st = np.zeros(500, dtype=synth.STATE)
    st["pos"] = np.stack([s["particles"]["px"], s["particles"]["py"], s["particles"]["pz"]], axis=1)
    st["rot"] = np.stack([s["particles"][k] for k in ("qx", "qy", "qz", "qw")], axis=1)
    prior = np.full(500, 1.0 / 500, np.float32)
    sp = eng_mod.ScanParams(lik_num_points=96, beam_num_points=6, seed=9)
    info = e.scan_prepare(raw, sp)
    assert info["n_lik"] == 96 and info["n_beam"] == 6 and info["n_downsampled"] < len(raw)
    e.particles_set(st, prior)
    a = e.particles_measure_update_prepared(s["origins"], 0.0)
    _, post_a = e.particles_get()
    e.particles_set(st, prior)
    b = e.particles_measure_update(e.scan_get(3), e.scan_get(4), s["origins"], 0.0)
    _, post_b = e.particles_get()
    assert a == b and np.array_equal(post_a, post_b) and a["kept"] == 1
    e.close()
