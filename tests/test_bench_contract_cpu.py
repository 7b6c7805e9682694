"""bench.py's reference arm runs without a GPU: check the one-JSON-line contract on the smallest workload."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--gpus", "2", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300,
                       cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_rotating_input_sets_of_the_back_to_back_protocol():
    """bench.py --l2 rotate: the stations are distinct places of the same map with their own scan and particle cloud
    (tracking workloads), every rank its own shard; station 0 is the scene itself; both arms name the same protocol."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    s, _, scaling, P_rank = bench.build_scene("c1", 0, 1)
    st = bench.build_stations("c1", 0, 1, s, 8)
    assert len(st) == 8 and st[0]["particles"] is s["particles"] and st[0]["lik"] is s["lik"]
    centres = np.array([[x["particles"]["px"].mean(), x["particles"]["py"].mean()] for x in st])
    d = np.linalg.norm(centres[:, None] - centres[None], axis=2) + np.eye(8) * 1e9
    assert d.min() > 0.5                                     # no two stations at the same place
    for x in st:
        assert len(x["particles"]) == P_rank and len(x["lik"]) == len(s["lik"]) and len(x["beam"]) == len(s["beam"])
        assert x["particles"].dtype == s["particles"].dtype and x["lik"].dtype == s["lik"].dtype
    assert not np.array_equal(st[1]["lik"], st[2]["lik"])
    # weak scaling: another rank draws other particles around the same stations
    st_r1 = bench.build_stations("c1", 1, 2, bench.build_scene("c1", 1, 2)[0], 8)
    assert not np.array_equal(st_r1[1]["particles"], st[1]["particles"])
    assert np.allclose(st_r1[1]["particles"]["px"].mean(), st[1]["particles"]["px"].mean(), atol=0.2)
    assert bench.config_dict("c1", s, 64, 96, 3, False, 0.2)["l2"] == bench.L2_TEXT[bench.L2_MODE]


def test_rotating_input_sets_shard_like_the_scene_under_strong_scaling():
    """c5 (strong scaling, spread particles): every station is one particle draw cut into the ranks' contiguous shards,
    exactly as build_scene cuts the scene itself; the scans are the scene's."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    P = bench.WORKLOADS["c5"][1]
    shards = []
    for rank in range(2):
        s, _, scaling, P_rank = bench.build_scene("c5", rank, 2)
        st = bench.build_stations("c5", rank, 2, s, 3)
        assert scaling == "strong" and P_rank == P // 2 and all(len(x["particles"]) == P_rank for x in st)
        assert st[1]["lik"] is s["lik"] and st[2]["beam"] is s["beam"]
        shards.append(st)
    full = bench.synth.spread_particles(P, bench.build_scene("c5", 0, 1)[0]["info"], seed=3000 + 17)
    assert np.array_equal(np.concatenate([shards[0][1]["particles"], shards[1][1]["particles"]]), full)
    assert not np.array_equal(shards[0][1]["particles"], shards[0][2]["particles"])


def test_e2e_leg_flow_and_its_fallback_to_ordinary_arrays(monkeypatch):
    """The Python side of bench.py's e2e leg with a stand-in engine (no GPU): page-locked caller arrays by default; if
    that path raises, the same measurement runs on ordinary arrays and the line carries a note."""
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from mcl_3dl_b200 import synth
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)

    class Eng:
        def __init__(self, fail_pinned):
            self.fail_pinned, self.pinned_arrays, self.calls = fail_pinned, 0, 0

        def host_array(self, n, dt):
            if self.fail_pinned:
                raise RuntimeError("no page-locked memory")
            self.pinned_arrays += 1
            return np.zeros(n, dt)

        def bind_measure(self, poses, lik, beam, org, out):
            assert poses.dtype == synth.POSE and out.dtype == synth.RESULT and len(out) == len(poses)

            def call():
                self.calls += 1
                out["match_cnt"] = 7
                return out
            return call

        def collect_timing(self, on):
            pass

        def last_timing(self):
            return {"h2d_ms": 0.0}

        def measure_update(self, poses, lik, beam, org, prior):
            return np.asarray(prior), {"entropy": 1.0, "kept": 1}, None

    class Cx:
        pass
    for fail in (False, True):
        cx = Cx()
        cx.world, cx.rank, cx.local, cx.dev, cx.notes = 1, 0, 0, torch.device("cpu"), []
        cx.flush = torch.zeros(8, dtype=torch.uint8)
        s, dda, _, P_rank = bench.build_scene("c1", 0, 1)
        eng = Eng(fail)
        live = {"eng": eng, "scene": s, "n_lik": len(s["lik"]), "n_beam": len(s["beam"]), "unit_pts": len(s["lik"]),
                "P_rank": P_rank}
        out = bench.e2e_leg(cx, "c1", "dda", 4, live)
        assert out["value"] > 0 and out["h2d_bytes_per_step"] == P_rank * 32 + (len(s["lik"]) + len(s["beam"]) + 2) * 16
        assert out["d2h_bytes_per_step"] == P_rank * 24 and out["fused_weight_update"]["kept"] == 1
        assert eng.calls == 3 + 4 + 1 and (live["out_host"]["match_cnt"] == 7).all()
        assert ("page-locked memory from mcl3dl_host_alloc" in out["host_buffers"]) == (not fail)
        assert eng.pinned_arrays == (0 if fail else 2) and len(cx.notes) == (1 if fail else 0)
