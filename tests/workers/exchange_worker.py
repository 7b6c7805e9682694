"""Worker of tests/test_gpu_exchange.py: one process per GPU (torchrun).  Every rank measures its own particle shard
with the record exchange folded into the kernels (mcl3dl_measure_exchange_device: peer-memory stores + signal kernel)
and checks the gathered array, byte for byte, against an NCCL all-gather of the plain mcl3dl_measure_device records —
eagerly over changing inputs, and replayed as a CUDA graph.  Prints EXCHANGE_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mcl_3dl_b200 import engine, synth  # noqa: E402


class DevView:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    P_rank, n_lik, n_beam = 777, 96, 24
    s = synth.scene(60_000, P_rank * world, n_lik, n_beam, seed=5)
    eng = engine.Engine((local,))
    eng.set_map(s["map"], engine.LikParams(dist_weight=(1, 1, 5)),
                engine.beam_params_from_reference(num_points_default=n_beam, dda_grid_size=0.2))
    u8 = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).to(dev)  # noqa: E731
    d_l, d_b = u8(s["lik"]), u8(s["beam"])
    d_o = torch.from_numpy(np.ascontiguousarray(s["origins"], dtype=np.float32)).to(dev)
    d_out = torch.zeros(P_rank * 24, dtype=torch.uint8, device=dev)
    d_all = torch.zeros(world * P_rank * 24, dtype=torch.uint8, device=dev)
    handle = eng.exchange_create(P_rank, world, rank)
    gathered = torch.empty(world * len(handle), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered, torch.tensor(list(handle), dtype=torch.uint8, device=dev))
    eng.exchange_open(bytes(gathered.cpu().numpy().tobytes()))
    dist.barrier()
    n_org = len(s["origins"])
    st = torch.cuda.current_stream().cuda_stream

    def check(d_p, ptr, tag):
        eng.measure_device(d_p.data_ptr(), P_rank, d_l.data_ptr(), n_lik, d_b.data_ptr(), n_beam, d_o.data_ptr(), n_org,
                           d_out.data_ptr(), st)
        dist.all_gather_into_tensor(d_all, d_out)
        torch.cuda.synchronize()
        got = torch.as_tensor(DevView(ptr, world * P_rank * 24), device=dev)
        assert torch.equal(got, d_all), "rank %d: peer exchange != NCCL all-gather (%s)" % (rank, tag)

    # eager, the particle shard changes every step (ranks drift apart in time: rank r sleeps r ms every other step)
    for step in range(12):
        parts = synth.tracking_particles(P_rank * world, s["truth_pos"], s["truth_rpy"], seed=100 + step)
        d_p = u8(parts[rank * P_rank:(rank + 1) * P_rank])
        if step % 2:
            torch.cuda._sleep(int(2e6) * (rank + 1))
        ptr = eng.measure_exchange_device(d_p.data_ptr(), P_rank, d_l.data_ptr(), n_lik, d_b.data_ptr(), n_beam,
                                          d_o.data_ptr(), n_org, st)
        check(d_p, ptr, "eager step %d" % step)
    # one CUDA graph replayed with the inputs rewritten in place between replays
    d_p = u8(s["particles"][rank * P_rank:(rank + 1) * P_rank])
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng.measure_exchange_device(d_p.data_ptr(), P_rank, d_l.data_ptr(), n_lik, d_b.data_ptr(), n_beam, d_o.data_ptr(), n_org,
                                    torch.cuda.current_stream().cuda_stream)
    for step in range(9):
        parts = synth.tracking_particles(P_rank * world, s["truth_pos"], s["truth_rpy"], seed=300 + step)
        d_p.copy_(u8(parts[rank * P_rank:(rank + 1) * P_rank]))
        for _ in range(1 + step % 3):
            g.replay()
        ptr, failed = eng.exchange_current(st)
        assert not failed
        check(d_p, ptr, "graph replay %d" % step)
    dist.barrier()
    if rank == 0:
        print("EXCHANGE_OK world=%d" % world)
    dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
