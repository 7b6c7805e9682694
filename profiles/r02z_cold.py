"""Where do the 20 us of the c2 likelihood launch go?  (diagnostic, not a bench value)

Times one graph-replayed launch of the c2 likelihood kernel with CUDA events under five cache states:
  cold        L2 flushed (256 MiB write) before the launch            -- bench.py's protocol
  code_warm   flushed, then the SAME kernel run on 2 poses far outside the map (touches no map data), then the launch
  warm        the launch repeated without a flush
  far_cold    flushed, then the far launch alone (fixed cost of a cold launch that loads no map data)
  far_warm    the far launch repeated without a flush (launch floor)
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MCL3DL_NF_TPP", "128")  # the same kernel instance for the 2-pose launch
import bench  # noqa: E402
from mcl_3dl_b200 import engine, synth  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    dev = torch.device("cuda:0")
    n_map, P, n_lik, n_beam, spread, _, _ = bench.WORKLOADS[wl]
    s, dda, _, P_rank = bench.build_scene(wl, 0, 1)
    eng = engine.Engine((0,))
    eng.set_map(s["map"], engine.LikParams(dist_weight=bench.DIST_WEIGHT), None)
    u8 = bench.as_u8
    d_p = torch.from_numpy(u8(s["particles"])).to(dev)
    far = synth.make_poses(np.full((2, 3), 1e5), np.tile([0, 0, 0, 1.0], (2, 1)))
    d_far = torch.from_numpy(u8(far)).to(dev)
    d_l = torch.from_numpy(u8(s["lik"])).to(dev)
    d_o = torch.zeros(16, dtype=torch.float32, device=dev)
    d_out = torch.zeros(P_rank * 24, dtype=torch.uint8, device=dev)
    d_out2 = torch.zeros(2 * 24, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def launch_main():
        eng.measure_device(d_p.data_ptr(), P_rank, d_l.data_ptr(), n_lik, d_o.data_ptr(), 0, d_o.data_ptr(), 1, d_out.data_ptr(),
                           torch.cuda.current_stream().cuda_stream)

    def launch_far():
        eng.measure_device(d_far.data_ptr(), 2, d_l.data_ptr(), n_lik, d_o.data_ptr(), 0, d_o.data_ptr(), 1, d_out2.data_ptr(),
                           torch.cuda.current_stream().cuda_stream)

    for _ in range(3):
        launch_main()
        launch_far()
    torch.cuda.synchronize()
    g_main, g_far = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_main):
        launch_main()
    with torch.cuda.graph(g_far):
        launch_far()

    def timed(pre, fn, k=200):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        torch.cuda.synchronize()
        for a, b in evs:
            pre()
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        per = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
        return {"min_us": round(per[0], 2), "median_us": round(per[len(per) // 2], 2), "p90_us": round(per[int(0.9 * len(per))], 2)}

    def fl():
        flush.fill_(1)

    def fl_far():
        flush.fill_(1)
        g_far.replay()

    res = {"workload": wl, "P": P_rank, "n_lik": n_lik}
    res["cold"] = timed(fl, g_main.replay)
    res["code_warm"] = timed(fl_far, g_main.replay)
    res["warm"] = timed(g_main.replay, g_main.replay)
    res["far_cold"] = timed(fl, g_far.replay)
    res["far_warm"] = timed(g_far.replay, g_far.replay)
    res["empty_events"] = timed(lambda: None, lambda: None)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
