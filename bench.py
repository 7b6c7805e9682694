#!/usr/bin/env python
"""bench.py — measurement-update throughput of the B200 engine (and, with --impl reference, of the
reference's own CPU path) on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one measurement update (the hot path) over one batch: every particle x every sampled
scan point.  Prints ONE JSON line on rank 0 (see the contract in the task statement):
  value     whole-job particle x point evals/s with all inputs resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the host-buffer C-ABI call mcl3dl_measure (H2D + kernels + D2H inside)
  roofline  dominant kernel: algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference CPU path (oracle/_ref if present, else the port) on a bounded sample, 1 thread
Workloads (BASELINE.json configs): c1 64x(96+3)/50k map, c2 1024x512 likelihood/1M map (default, the
metric's config), c3 4096x256 beam DDA/1M map, c4 16384x1024 lik+beam/10M map, c5 65536 spread lik+beam.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mcl_3dl_b200 import synth  # noqa: E402

WORKLOADS = {
    #        map pts   P      n_lik n_beam spread dda   scaling
    "c1": (50_000, 64, 96, 3, False, 0.2, "weak"),
    "c2": (1_000_000, 1024, 512, 0, False, 0.2, "weak"),
    "c3": (1_000_000, 4096, 0, 256, False, 0.2, "weak"),
    "c4": (10_000_000, 16384, 1024, 1024, False, 0.1, "strong"),
    "c5": (1_000_000, 65536, 64, 8, True, 0.2, "strong"),
}
FORCE_SPREAD = False
DIST_WEIGHT = (1.0, 1.0, 5.0)  # the node's default metric (src/parameters.cpp:108-111)
MAP_VOXEL = 0.1


def bytes_per_eval_model(match_dist_min=0.2, w=DIST_WEIGHT, h=MAP_VOXEL):
    """SURVEY.md §8(d), exact mode: 16 B x cells of the h-lattice inside the search ellipsoid's box."""
    cells = 1
    for k in range(3):
        cells *= 2 * math.ceil(match_dist_min / (w[k] * h) - 1e-9) + 1
    return 16 * cells, cells


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "span": "warm-up, timed steps, ~0.5 s of the same step untimed, roofline and e2e loops (100 ms period)"}


def as_u8(a):
    return np.ascontiguousarray(a).view(np.uint8).reshape(-1)


def build_scene(workload, rank, world):
    n_map, P, n_lik, n_beam, spread, dda, scaling = WORKLOADS[workload]
    spread = spread or FORCE_SPREAD
    P_rank = P if scaling == "weak" else P // world
    s = synth.scene(n_map, P, n_lik, n_beam, spread=spread, n_origins=2, seed=1000)
    if scaling == "weak":
        # same map and scan, this rank's own particle draw
        if spread:
            s["particles"] = synth.spread_particles(P, s["info"], seed=2000 + rank)
        else:
            s["particles"] = synth.tracking_particles(P, s["truth_pos"], s["truth_rpy"], seed=2000 + rank)
    else:
        s["particles"] = s["particles"][rank * P_rank:(rank + 1) * P_rank]
    return s, dda, scaling, P_rank


USE_DDA = True


def cpu_arm(workload, s, dda, n_lik, n_beam, target_s, threads, want_kind=None):
    """Time the reference CPU path on a bounded particle sample.  Returns (evals/s, meta)."""
    from oracle import cpu_checker as cc
    kind = "reference" if cc.available("reference") else "port"
    if want_kind:
        kind = want_kind
    if kind == "port":
        cc.build("port")
    chk = cc.CpuChecker(kind)
    t0 = time.perf_counter()
    cpu = chk.create(s["map"], cc.lik_params(dist_weight=DIST_WEIGHT),
                     cc.beam_raw(num_points_default=max(n_beam, 1), dda_grid_size=dda, use_raycast_using_dda=USE_DDA),
                     20.0, 0.4)
    build_s = time.perf_counter() - t0
    cpu.set_tally(False)
    per_particle = max(n_lik + n_beam, 1)
    probe = s["particles"][:max(threads, 4)]
    t0 = time.perf_counter()
    cpu.measure(probe, s["lik"], s["beam"], s["origins"], n_threads=threads)  # also builds the lazy DDA grid
    t0 = time.perf_counter()
    cpu.measure(probe, s["lik"], s["beam"], s["origins"], n_threads=threads)
    dt = max(time.perf_counter() - t0, 1e-6)
    rate = len(probe) * per_particle / dt
    n_sample = int(min(len(s["particles"]), max(len(probe), rate * target_s / per_particle)))
    sample = s["particles"][:n_sample]
    reps = int(max(1, min(200, round(target_s / max(n_sample * per_particle / rate, 1e-6)))))
    t0 = time.perf_counter()
    for _ in range(reps):
        cpu.measure(sample, s["lik"], s["beam"], s["origins"], n_threads=threads)
    dt = (time.perf_counter() - t0) / reps
    meta = {"kind": kind, "cores": threads, "index_build_s": round(build_s, 3), "seconds": round(dt * reps, 3),
            "sample": "%d of %d particles x (%d lik + %d beam) pts x %d repeats, same map/scan, %d thread(s)"
                      % (n_sample, len(s["particles"]), n_lik, n_beam, reps, threads)}
    return n_sample, dt, cpu, meta


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_map, P, n_lik, n_beam, spread, dda, scaling = WORKLOADS[args.workload]
    s, dda, scaling, P_rank = build_scene(args.workload, 0, 1)
    threads = os.cpu_count() or 1
    unit_pts = n_lik if n_lik else n_beam
    # give every thread enough particles to amortise its start-up: tile the particle set to >= 64 per thread
    need = threads * 64
    if len(s["particles"]) < need:
        s["particles"] = np.tile(s["particles"], (need + len(s["particles"]) - 1) // len(s["particles"]))
    n_sample, dt, cpu, meta = cpu_arm(args.workload, s, dda, n_lik, n_beam, 2.0, threads)
    sample = s["particles"][:n_sample]
    for _ in range(args.warmup):
        cpu.measure(sample, s["lik"], s["beam"], s["origins"], n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu.measure(sample, s["lik"], s["beam"], s["origins"], n_threads=threads)
    dt = time.perf_counter() - t0
    value = n_sample * unit_pts * args.steps / dt
    meta["value"] = value
    meta["unit"] = "evals/s"
    line = {"impl": "reference", "metric": metric_name(n_lik), "value": value, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.workload, s, P, n_lik, n_beam, spread, dda, None),
            "cpu_baseline": meta,
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def metric_name(n_lik):
    return "particle x point likelihood evals/s" if n_lik else "particle x ray beam-model evals/s"


def config_dict(workload, s, P, n_lik, n_beam, spread, dda, info):
    d = {"workload": "%s: %d particles x (%d likelihood pts + %d beam rays), %d-pt map @%.1f m voxel, %s particles"
                     % (workload, P, n_lik, n_beam, len(s["map"]), MAP_VOXEL, "spread" if spread else "tracking"),
         "dist_weight": list(DIST_WEIGHT), "dda_grid_size": dda, "match_dist_min": 0.2,
         "raycaster": "RaycastUsingDDA" if USE_DDA else "RaycastUsingKDTree",
         "l2": "flushed (256 MiB write) before every timed step"}
    if info is not None:
        d["nn_grid"] = list(info.nn_dims)
        d["dda_grid"] = list(info.dda_dims)
        d["map_device_bytes"] = int(info.device_bytes)
        d["map_build_ms"] = round(info.build_ms, 3)
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer"],
                    help="N > 1: how the records are gathered: NCCL all-gather (contract run) or the engine's own "
                         "peer-memory exchange kernel (experiment, not yet run on hardware)")
    ap.add_argument("--graph", action="store_true",
                    help="experiment: replay the device-resident step as one CUDA graph (not part of the contract run)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--raycaster", default="dda", choices=["dda", "kd"],
                    help="beam raycaster: RaycastUsingDDA (north_star) or RaycastUsingKDTree (the node's default)")
    ap.add_argument("--spread", action="store_true", help="spread (global-localisation style) particles: the HBM-bound variant")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    global FORCE_SPREAD, USE_DDA
    FORCE_SPREAD = args.spread
    USE_DDA = args.raycaster == "dda"
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from mcl_3dl_b200 import engine, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = os.environ.get("MCL3DL_NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    n_map, P, n_lik, n_beam, spread, _, _ = WORKLOADS[args.workload]
    spread = spread or args.spread
    s, dda, scaling, P_rank = build_scene(args.workload, rank, world)
    eng = engine.Engine((local,))  # no fallback: raises without the CUDA library / device
    lik = engine.LikParams(dist_weight=DIST_WEIGHT)
    use_dda = args.raycaster == "dda"
    beam = engine.beam_params_from_reference(num_points_default=max(n_beam, 1), dda_grid_size=dda,
                                             use_raycast_using_dda=use_dda)
    eng.set_map(s["map"], lik if (n_lik or not use_dda) else None, beam if n_beam else None)
    info = eng.map_info()

    particles = s["particles"]
    d_p = torch.from_numpy(as_u8(particles)).to(dev)
    d_l = torch.from_numpy(as_u8(s["lik"])).to(dev) if n_lik else torch.zeros(16, dtype=torch.uint8, device=dev)
    d_b = torch.from_numpy(as_u8(s["beam"])).to(dev) if n_beam else torch.zeros(16, dtype=torch.uint8, device=dev)
    d_o = torch.from_numpy(np.ascontiguousarray(s["origins"], dtype=np.float32)).to(dev)
    d_out = torch.zeros(P_rank * 24, dtype=torch.uint8, device=dev)
    d_all = torch.zeros(world * P_rank * 24, dtype=torch.uint8, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    peer_all = None
    if world > 1 and args.exchange == "peer":
        # EXPERIMENT (never run): every rank's records are stored into every rank's buffer by one kernel over NVLink;
        # the CUDA IPC handles of the buffers are the only thing that goes through torch.distributed
        handle = eng.exchange_create(P_rank, world, rank)
        peer_all = 0  # device address of the last gathered array, returned by every exchange_records call
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        gathered = torch.empty(world * len(handle), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(gathered, mine)
        eng.exchange_open(bytes(gathered.cpu().numpy().tobytes()))
        dist.barrier()

    def step_eager():
        nonlocal peer_all
        st = torch.cuda.current_stream().cuda_stream
        eng.measure_device(d_p.data_ptr(), P_rank, d_l.data_ptr(), n_lik, d_b.data_ptr(), n_beam,
                           d_o.data_ptr(), len(s["origins"]), d_out.data_ptr(), st)
        if world > 1:
            if peer_all is not None:
                peer_all = eng.exchange_records(d_out.data_ptr(), P_rank, st)
            else:
                # the one exchange of the path: all-gather of the per-particle records over NVLink (NCCL)
                sharding.gather_records_device(d_out, d_all)

    step = step_eager
    if args.graph and peer_all is not None:
        raise SystemExit("--graph cannot replay --exchange peer: the step counter is a kernel argument")
    if args.graph:
        # EXPERIMENT (added without GPU time left in round 1, never run): replay the step (both kernels, their
        # fork/join events and the NCCL all-gather) as one CUDA graph, to take the per-step launch work off the CPU
        for _ in range(3):
            step_eager()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step_eager()
        step = graph.replay

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        barrier()
        for a, b in evs:
            flush.fill_(1)          # L2 flush, outside the timed span
            a.record()
            fn()
            b.record()
        barrier()
        tot = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([tot], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # clocks are sampled from the warm-up to the end of the e2e loop: the timed region alone lasts a few
    # milliseconds, shorter than one nvidia-smi sampling period
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = eng.kernel_launches()
    total_ms = timed(step, args.steps)
    launches = eng.kernel_launches() - l0
    # the timed region lasts a few milliseconds, shorter than one nvidia-smi period: keep the same step running
    # (untimed, same count on every rank) for ~0.5 s so that the clock record holds samples taken under this load
    n_extra = int(min(20000, max(0, 500.0 / max(total_ms / args.steps, 1e-3))))
    for _ in range(n_extra):
        step()
    barrier()

    exchange_info = None
    if world > 1:
        exchange_info = {"mode": args.exchange, "graph": bool(args.graph)}
        if peer_all is not None:
            # check the experiment against NCCL on the same records (outside every timed region)
            class _DevView:  # raw device memory as a torch tensor (CUDA array interface)
                def __init__(self, ptr, nbytes):
                    self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
            step_eager()
            sharding.gather_records_device(d_out, d_all)
            torch.cuda.synchronize()
            t_peer = torch.as_tensor(_DevView(peer_all, world * P_rank * 24), device=dev)
            exchange_info["matches_nccl"] = bool(torch.equal(t_peer, d_all))
            exchange_info["peer_wait_timed_out"] = eng.exchange_failed()

    unit_pts = n_lik if n_lik else n_beam
    evals_step = world * P_rank * unit_pts
    ms_per_step = total_ms / args.steps
    value = evals_step / (ms_per_step * 1e-3)

    # ---- dominant-kernel roofline: time each model's kernel alone (one launch per call)
    kern = {}
    if n_lik:
        def lik_only():
            eng.measure_device(d_p.data_ptr(), P_rank, d_l.data_ptr(), n_lik, 0, 0, 0, 0, d_out.data_ptr(), stream)
        kern["lik"] = timed(lik_only, args.steps) / args.steps
    if n_beam:
        def beam_only():
            eng.measure_device(d_p.data_ptr(), P_rank, 0, 0, d_b.data_ptr(), n_beam, d_o.data_ptr(), len(s["origins"]),
                               d_out.data_ptr(), stream)
        kern["beam"] = timed(beam_only, args.steps) / args.steps
    peak, peak_src = peaks()
    dom = max(kern, key=kern.get)
    # exact work counters of one (untimed) step: what this layout's algorithm must read, no reuse assumed
    eng.collect_stats(True)
    step()
    ws = eng.read_stats()
    eng.collect_stats(False)
    io_bytes = P_rank * 32 + P_rank * 24
    bpe, cells = bytes_per_eval_model()
    near = eng.near_field_info()  # [(k, bytes)] likelihood / KD-caster screens; k = 0: not staged
    if dom == "lik":
        alg_bytes = ws["lik_index_rows"] * 8 + ws["lik_points_scanned"] * 16 + io_bytes + n_lik * 16
        note = ("counted: %.1f CSR rows x 8 B + %.1f map points x 16 B per eval (+ poses/scan/records)"
                % (ws["lik_index_rows"] / max(P_rank * n_lik, 1), ws["lik_points_scanned"] / max(P_rank * n_lik, 1)))
        if near[0][0]:
            alg_bytes += P_rank * n_lik * 4
            note += " + one 4 B near-field word per eval (k=%d, %.0f MB)" % (near[0][0], near[0][1] / 1e6)
        survey_bytes = P_rank * n_lik * bpe + io_bytes + n_lik * 16
        survey_note = "SURVEY 8d exact mode, dense float4 voxel grid: %d B/eval = 16 B x %d cells" % (bpe, cells)
    else:
        alg_bytes = (ws["beam_cells_stepped"] * 4 + ws["beam_cells_occupied"] * 8 + ws["beam_points_tested"] * 16
                     + io_bytes + n_beam * 16)
        note = ("counted: %.1f cells stepped x 4 B occupancy word + %.2f occupied x 8 B CSR + %.2f pts x 16 B per ray"
                % (ws["beam_cells_stepped"] / max(P_rank * n_beam, 1), ws["beam_cells_occupied"] / max(P_rank * n_beam, 1),
                   ws["beam_points_tested"] / max(P_rank * n_beam, 1)))
        survey_bytes = (ws["beam_cells_stepped"] * 1 + ws["beam_cells_occupied"] * 8 + ws["beam_points_tested"] * 16
                        + io_bytes + n_beam * 16)
        survey_note = "SURVEY 8d beam: 1 B per cell stepped + 8 B CSR + 16 B per point tested"
    achieved = alg_bytes / (kern[dom] * 1e-3) / 1e9
    traffic, traffic_note = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            ncu = json.load(f)
        traffic = ncu.get(args.workload, {}).get(dom + "_dram_bytes_per_launch")
        traffic_note = ncu.get("_note")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom + "_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic, "traffic_note": traffic_note,
                "kernel_ms": kern[dom], "algorithmic_bytes_per_launch": alg_bytes, "model": note,
                "kernel_ms_all": kern, "work_counters_per_step": ws,
                "survey_8d_model": {"bytes_per_launch": survey_bytes, "achieved": survey_bytes / (kern[dom] * 1e-3) / 1e9,
                                    "frac": survey_bytes / (kern[dom] * 1e-3) / 1e9 / peak, "note": survey_note}}

    # ---- e2e: the host-buffer C-ABI call (pinned staging + H2D + kernels + D2H inside the call)
    out_host = np.zeros(P_rank, dtype=synth.RESULT)
    # the buffer addresses are resolved once (Engine.bind_measure), as a C++ caller's would be: every timed call is
    # exactly one mcl3dl_measure(host pointers) = staging + H2D + kernels + D2H + synchronise
    h_in = [np.ascontiguousarray(a, dtype=dt) for a, dt in ((particles, synth.POSE), (s["lik"], synth.POINT),
                                                             (s["beam"], synth.POINT))]
    h_org = np.ascontiguousarray(s["origins"], dtype=np.float32).reshape(-1, 3)
    e2e_call = eng.bind_measure(h_in[0], h_in[1], h_in[2], h_org, out_host)
    for _ in range(3):
        e2e_call()
    barrier()
    e2e_tot = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_call()
        e2e_tot += time.perf_counter() - t0
    eng.collect_timing(True)   # device-side breakdown of one extra, untimed call (the events cost ~28 us per call,
    e2e_call()                 # so they are off while the e2e loop above is timed)
    last_call_device_ms = eng.last_timing()
    eng.collect_timing(False)
    t = torch.tensor([e2e_tot], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e = {"value": evals_step * args.steps / e2e_s, "unit": "evals/s",
           "h2d_bytes_per_step": world * (P_rank * 32 + n_lik * 16 + n_beam * 16 + len(s["origins"]) * 16),
           "d2h_bytes_per_step": world * P_rank * 24, "ms_per_step": 1e3 * e2e_s / args.steps,
           "timing": "host wall clock around the synchronous call", "last_call_device_ms": last_call_device_ms,
           "d2h_mode": ("kernels store the records straight into the pinned host block" if P_rank <= 8192
                        else "one D2H copy of the records")}

    # ---- e2e with the fused weight update (scope row f2): priors up, posteriors (4 B/particle) back
    prior = np.full(P_rank, 1.0 / max(P_rank, 1), dtype=np.float32)
    for _ in range(3):
        eng.measure_update(particles, s["lik"], s["beam"], s["origins"], prior)
    fused_tot = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        post, summ, _ = eng.measure_update(particles, s["lik"], s["beam"], s["origins"], prior)
        fused_tot += time.perf_counter() - t0
    t = torch.tensor([fused_tot], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e["fused_weight_update"] = {"value": evals_step * args.steps / float(t.item()), "unit": "evals/s",
                                  "ms_per_step": 1e3 * float(t.item()) / args.steps,
                                  "h2d_bytes_per_step": e2e["h2d_bytes_per_step"] + world * P_rank * 4,
                                  "d2h_bytes_per_step": world * P_rank * 4,
                                  "entropy": summ["entropy"], "kept": summ["kept"]}
    clk = clocks.stop() if rank == 0 else None

    # ---- sanity: device-resident records == host-path records
    got = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=synth.RESULT)
    if n_lik and n_beam:
        assert np.array_equal(got, out_host), "device-resident and host entry points disagree"

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            n_sample, dt, cpu, meta = cpu_arm(args.workload, s, dda, n_lik, n_beam, args.cpu_seconds, 1)
            meta["value"] = n_sample * unit_pts / dt
            meta["unit"] = "evals/s"
            # parity spot check of this very workload against the checker (first particles of rank 0)
            cpu.set_tally(True)
            chk = cpu.measure(particles[:16], s["lik"], s["beam"], s["origins"])
            ok = all(np.array_equal(chk[f], out_host[:16][f]) for f in ("match_cnt", "n_short", "n_hit", "n_long"))
            ok = ok and np.allclose(chk["score_like"], out_host[:16]["score_like"], rtol=1e-4, atol=1e-6)
            meta["parity_spot_check"] = bool(ok)
            cpu_baseline = meta
        except Exception as exc:  # the GPU line must still be printed if the checker cannot be built / loaded
            cpu_baseline = {"error": "%s: %s" % (type(exc).__name__, exc)}

    if rank == 0:
        line = {"metric": metric_name(n_lik), "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(args.workload, s, P if scaling == "strong" else P_rank * world, n_lik, n_beam,
                                      spread, dda, info),
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu_baseline, "exchange": exchange_info,
                "match_ratio_mean": float(out_host["match_cnt"].mean() / max(n_lik, 1)),
                "beam_tallies_mean": [float(out_host[f].mean()) for f in ("n_short", "n_hit", "n_long")]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
