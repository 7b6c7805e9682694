#!/usr/bin/env python
"""A/B sweep of engine variants on one GPU, sized for a ~1 minute box slot (no torch import).

    python profiles/ab_variants.py --build-variants                 # here (nvcc, no GPU): the variant libraries
    python profiles/ab_variants.py --out gpurun_out/r01y_ab.jsonl [--budget 60]   # on the GPU box

For every (workload, variant) job: create an engine from the variant's library with the variant's environment,
stage the map, run the host-buffer call (mcl3dl_measure through Engine.bind_measure) a few dozen times and report
  * whether the records are BYTE-IDENTICAL to the baseline variant's on the same inputs,
  * median kernel time (CUDA events of the engine, when its timing events are on) and median e2e wall time.
The work runs in a child process that prints one JSON line per job; if the child dies or hangs in a variant the
parent records that and restarts it on the remaining jobs, so one bad variant cannot take the sweep down.
Numbers: warm L2 (no flush between calls), wall clock around a synchronous call — for A/B ranking only; the
bench.py contract numbers stay the reference.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = os.path.join(ROOT, "mcl_3dl_b200")
LIB = os.path.join(PKG, "libmcl3dl_b200.so")          # the default build
U2_LIB = os.path.join(PKG, "libmcl3dl_b200_nfu2.so")  # lik_kernel_nf with 2 instead of 4 evals per lane in flight
HOST = {"MCL3DL_TIMING": "1", "MCL3DL_ZEROCOPY_OUT": "8192"}  # kernel times wanted: the timing events stay on
VARIANT_BUILDS = {}

# name -> (library, environment).  "base" (the CSR-window kernels of round 1, MCL3DL_NNF=0) must come first: everything
# is compared with its records byte for byte.
VARIANTS = [
    ("base", LIB, dict(HOST, MCL3DL_NNF="0", MCL3DL_BEAM="pl")),   # round-1 search structures, static beam kernel
    ("dflt", LIB, dict(HOST)),                                      # today's defaults
    ("beam_pl", LIB, dict(HOST, MCL3DL_BEAM="pl")),
    ("beam_dq", LIB, dict(HOST, MCL3DL_BEAM="dq")),
    ("kd_nor2", LIB, dict(HOST, MCL3DL_NNF_KD_R2="0")),
    ("fast_host", LIB, {"MCL3DL_TIMING": "0", "MCL3DL_ZEROCOPY_OUT": "8192"}),
    ("group", LIB, dict(HOST, MCL3DL_MAPPING="group")),
]
# workload -> (bench workload, raycaster, spread override)
WORKLOADS = [("c2", "c2", "dda", False), ("c3kd", "c3", "kd", False), ("c5", "c5", "dda", False),
             ("c1kd", "c1", "kd", False), ("c2s", "c2", "dda", True), ("c3", "c3", "dda", False), ("c2iso", "c2", "dda", False)]
ISO_WORKLOADS = {"c2iso"}  # dist_weight (1,1,1) instead of the node's (1,1,5): ~3x more map points per eval
ENV_KEYS = ["MCL3DL_TIMING", "MCL3DL_ZEROCOPY_OUT", "MCL3DL_NEAR_K", "MCL3DL_NEAR_KD_K", "MCL3DL_NEAR_MAX_MB",
            "MCL3DL_MAPPING", "MCL3DL_UPDATE_ONE_SYNC", "MCL3DL_NNF", "MCL3DL_NF_STAGE", "MCL3DL_NF_TPP", "MCL3DL_NNF_KD_R2", "MCL3DL_BEAM", "MCL3DL_BEAM_DQ_PPL"]


def jobs_all():
    return [(w[0], v[0]) for w in WORKLOADS for v in VARIANTS]


def child(job_names, calls):
    import bench
    from mcl_3dl_b200 import engine, synth
    variants = {v[0]: v for v in VARIANTS}
    workloads = {w[0]: w for w in WORKLOADS}
    scenes = {}
    for wname, vname in job_names:
        _, bw, caster, spread = workloads[wname]
        _, lib, env = variants[vname]
        if wname not in scenes:
            bench.FORCE_SPREAD = spread
            s, dda, _, _ = bench.build_scene(bw, 0, 1)
            bench.FORCE_SPREAD = False
            scenes = {wname: (s, dda)}  # keep one scene in memory at a time
        s, dda = scenes[wname]
        n_lik, n_beam = len(s["lik"]), len(s["beam"])
        for k in ENV_KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        rec = {"workload": wname, "variant": vname}
        if not os.path.exists(lib):
            rec["error"] = "library not built"
            print(json.dumps(rec), flush=True)
            continue
        print(json.dumps({"start": [wname, vname]}), flush=True)
        eng = engine.Engine((0,), lib_path=lib)
        dw = (1.0, 1.0, 1.0) if wname in ISO_WORKLOADS else bench.DIST_WEIGHT
        lik = engine.LikParams(dist_weight=dw) if (n_lik or caster == "kd") else None
        beam = (engine.beam_params_from_reference(num_points_default=max(n_beam, 1), dda_grid_size=dda,
                                                  use_raycast_using_dda=(caster == "dda")) if n_beam else None)
        eng.set_map(s["map"], stamp=1, lik=lik, beam=beam)
        out = np.zeros(len(s["particles"]), dtype=synth.RESULT)
        poses = np.ascontiguousarray(s["particles"], dtype=synth.POSE)
        call = eng.bind_measure(poses, np.ascontiguousarray(s["lik"], dtype=synth.POINT),
                                np.ascontiguousarray(s["beam"], dtype=synth.POINT),
                                np.ascontiguousarray(s["origins"], dtype=np.float32).reshape(-1, 3), out)
        for _ in range(4):
            call()
        wall, k_lik, k_beam = [], [], []
        timed = env.get("MCL3DL_TIMING", "0") != "0"
        for _ in range(calls):
            t0 = time.perf_counter()
            call()
            wall.append(time.perf_counter() - t0)
            if timed:
                lt = eng.last_timing()
                k_lik.append(lt["lik_ms"])
                k_beam.append(lt["beam_ms"])
        base_path = "/tmp/ab_base_%s.npy" % wname
        if vname == "base":
            np.save(base_path, out)
            rec["identical_to_base"] = True
        elif os.path.exists(base_path):
            ref = np.load(base_path)
            rec["identical_to_base"] = bool(ref.tobytes() == out.tobytes())
            # variants that give a particle another number of lanes sum its likelihood terms in another order: everything
            # integer (and the beam score, a product of identical factors) must still be exact, the likelihood sum to rounding
            rec["exact_fields_identical"] = bool(all(np.array_equal(ref[f], out[f]) for f in
                                                     ("match_cnt", "score_beam", "n_short", "n_hit", "n_long")))
            rec["score_like_max_rel_diff"] = float(np.max(np.abs(ref["score_like"] - out["score_like"]) /
                                                          np.maximum(np.abs(ref["score_like"]), 1e-6))) if len(out) else 0.0
        else:
            rec["identical_to_base"] = None
        # the fused weight update (mcl3dl_measure_update) on the same inputs: posterior compared with the base variant's
        try:  # (added after the r01y run: never executed on a GPU yet, so it must not take the sweep down)
            prior = np.full(len(poses), 1.0 / len(poses), dtype=np.float32)
            upd = []
            for k in range(3 + max(calls // 3, 3)):
                t0 = time.perf_counter()
                post, summ, _ = eng.measure_update(poses, s["lik"], s["beam"], s["origins"], prior)
                if k >= 3:
                    upd.append(time.perf_counter() - t0)
            post_path = "/tmp/ab_base_post_%s.npy" % wname
            if vname == "base":
                np.save(post_path, post)
                rec["posterior_identical_to_base"] = True
            elif os.path.exists(post_path):
                rec["posterior_identical_to_base"] = bool(np.load(post_path).tobytes() == post.tobytes())
            rec["update_us_median"] = 1e6 * float(np.median(upd))
            rec["entropy"] = float(summ["entropy"])
        except Exception as exc:
            rec["update_error"] = "%s: %s" % (type(exc).__name__, exc)
        units = len(poses) * (n_lik if n_lik else n_beam)
        rec.update({"e2e_us_median": 1e6 * float(np.median(wall)), "e2e_us_min": 1e6 * float(np.min(wall)),
                    "e2e_units_per_s": units / float(np.median(wall)),
                    "lik_kernel_us": 1e3 * float(np.median(k_lik)) if k_lik and n_lik else None,
                    "beam_kernel_us": 1e3 * float(np.median(k_beam)) if k_beam and n_beam else None,
                    "near_field": eng.near_field_info(), "build_ms": eng.map_info().build_ms, "map_device_mb": eng.map_info().device_bytes / 1e6,
                    "match_cnt_sum": int(out["match_cnt"].sum()), "n_hit_sum": int(out["n_hit"].sum()), "calls": calls})
        eng.close()
        print(json.dumps(rec), flush=True)


def parent(out_path, budget, calls, only):
    t_start = time.time()
    jobs = [j for j in jobs_all() if not only or j[0] in only]
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    f = open(out_path, "w")
    while jobs and time.time() - t_start < budget:
        left = budget - (time.time() - t_start)
        p = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--child", json.dumps(jobs), "--calls", str(calls)],
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        try:
            so, se = p.communicate(timeout=max(left, 1.0))
            died = p.returncode != 0
        except subprocess.TimeoutExpired:
            p.kill()
            so, se = p.communicate()
            died = True
        started = None
        for line in so.splitlines():
            try:
                r = json.loads(line)
            except ValueError:
                continue
            if "start" in r:
                started = tuple(r["start"])
                continue
            f.write(json.dumps(r) + "\n")
            f.flush()
            jobs.remove((r["workload"], r["variant"]))
            started = None
        if died:
            bad = started if started in jobs else (jobs[0] if jobs else None)
            if bad:
                f.write(json.dumps({"workload": bad[0], "variant": bad[1], "error": "child died or timed out",
                                    "stderr_tail": se[-600:]}) + "\n")
                f.flush()
                jobs.remove(bad)
    for j in jobs:
        f.write(json.dumps({"workload": j[0], "variant": j[1], "error": "not run (time budget)"}) + "\n")
    f.write(json.dumps({"elapsed_s": time.time() - t_start}) + "\n")
    f.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ab_variants.jsonl"))
    ap.add_argument("--budget", type=float, default=60.0, help="seconds of wall clock for the whole sweep")
    ap.add_argument("--calls", type=int, default=30)
    ap.add_argument("--only", default="", help="comma-separated workload names")
    ap.add_argument("--child", default=None)
    ap.add_argument("--build-variants", action="store_true", help="nvcc-build the variant libraries and exit (no GPU needed)")
    a = ap.parse_args()
    if a.build_variants:
        from mcl_3dl_b200 import build as _build
        for lib, defines in VARIANT_BUILDS.items():
            print(_build.build(defines=defines, out=lib))
    elif a.child is not None:
        child([tuple(j) for j in json.loads(a.child)], a.calls)
    else:
        parent(a.out, a.budget, a.calls, [w for w in a.only.split(",") if w])
