"""The N>1 path on CPU: world_size-2 gloo processes, each computing its particle shard (with the CPU
oracle standing in for the device), one all-gather of the records, host-side weight update."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_particles, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from mcl_3dl_b200 import sharding, synth
    from oracle import cpu_checker as cc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = synth.scene(20_000, n_particles, 48, 12, seed=77)
    port_chk = cc.CpuChecker("port")
    cpu = port_chk.create(s["map"], cc.lik_params(dist_weight=(1, 1, 5)), cc.beam_raw(num_points_default=12))
    b, e = sharding.shard_bounds(n_particles, world)[rank]
    local = cpu.measure(s["particles"][b:e], s["lik"], s["beam"], s["origins"])
    full = sharding.gather_records(local, n_particles)
    if rank == 0:
        want = cpu.measure(s["particles"], s["lik"], s["beam"], s["origins"])
        prior = np.full(n_particles, 1.0 / n_particles, dtype=np.float32)
        post, ent, kept, rmin, rmax = sharding.posterior(prior, full, 48)
        like = (want["score_beam"] * want["score_like"]).astype(np.float32)
        ref_post, ref_ent, ref_kept = port_chk.pf_update(prior, like)
        q.put((np.array_equal(full, want), bool(kept) == bool(ref_kept), float(np.abs(post - ref_post).max()),
               abs(ent - ref_ent), rmin, rmax, float((want["match_cnt"] / 48.0).min())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_particles", [64, 101])
def test_two_rank_gather_and_weight_update(port, n_particles):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, p, n_particles, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    same, kept_same, dpost, dent, rmin, rmax, want_min = res
    assert same and kept_same
    assert dpost == 0.0 and dent < 1e-6
    assert abs(rmin - want_min) < 1e-7 and rmax >= rmin


def test_shard_bounds_cover_everything():
    from mcl_3dl_b200 import sharding
    for n in (0, 1, 7, 64, 65536, 100003):
        for w in (1, 2, 4, 8):
            b = sharding.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1


def test_posterior_restores_when_all_zero():
    from mcl_3dl_b200 import sharding, synth
    rec = np.zeros(5, dtype=synth.RESULT)
    rec["score_beam"] = 1.0
    prior = np.full(5, 0.2, np.float32)
    post, ent, kept, _, _ = sharding.posterior(prior, rec, 10)
    assert not kept and np.array_equal(post, prior) and ent is None
