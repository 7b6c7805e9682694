"""Multi-GPU plumbing of the measurement update: contiguous particle shards, one all-gather of the
per-particle records, and the unchanged host-side weight update.

Particles are independent units (include/mcl_3dl/pf.h:256-260 loops them serially); the only
cross-particle work — normalisation, entropy, resampling (pf.h:261-279,182-225) — happens on the
host after every record exists.  So: map replicated per GPU, particles split into contiguous blocks
[r*P/G, (r+1)*P/G), ONE all-gather of the 24-byte records (NCCL over NVLink when the records live on
the device, gloo for CPU tensors in tests), then `posterior()`.
"""
import numpy as np

from .synth import RESULT


def shard_bounds(n_particles, world):
    """Same split as the C engine uses across its devices (engine.cu: p0[d] = P*d/G)."""
    return [(n_particles * r // world, n_particles * (r + 1) // world) for r in range(world)]


def gather_records(local, n_total, group=None):
    """All-gather RESULT records held as a numpy array (CPU tensors -> works with the gloo backend).
    Shards may be uneven; every rank returns the full [n_total] array in particle order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bounds = shard_bounds(n_total, world)
    longest = max(e - b for b, e in bounds)
    buf = np.zeros(longest, dtype=RESULT)
    buf[:len(local)] = local
    t = torch.from_numpy(buf.view(np.uint8).reshape(-1).copy())
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    full = np.zeros(n_total, dtype=RESULT)
    for (b, e), o in zip(bounds, outs):
        full[b:e] = np.frombuffer(o.numpy().tobytes(), dtype=RESULT)[:e - b]
    return full


def gather_records_device(d_local_u8, d_all_u8, group=None):
    """Equal-shard device-resident variant: torch uint8 CUDA tensors, NCCL all_gather_into_tensor."""
    import torch.distributed as dist
    dist.all_gather_into_tensor(d_all_u8, d_local_u8, group=group)
    return d_all_u8


def posterior(prob, records, n_lik, extra_likelihood=None):
    """pf::ParticleFilter::measure's weight update (pf.h:252-279) fed by the node's lambda
    (src/mcl_3dl.cpp:402-425): per particle likelihood = beam * likelihood (map-key order) [* odom term],
    float32 throughout, sequential sums.  Returns (prob, entropy, kept, match_ratio_min, match_ratio_max)."""
    prob = np.array(prob, dtype=np.float32)
    like = np.float32(1.0) * records["score_beam"].astype(np.float32)
    like = (like * records["score_like"].astype(np.float32)).astype(np.float32)
    if extra_likelihood is not None:
        like = (like * np.asarray(extra_likelihood, dtype=np.float32)).astype(np.float32)
    quality = (records["match_cnt"].astype(np.float32) / np.float32(max(n_lik, 1))) if n_lik else np.zeros(len(prob), np.float32)
    new = (prob * like).astype(np.float32)
    total = np.float32(0)
    for v in new:  # sequential float32 accumulate, as the reference's loop
        total = np.float32(total + v)
    ratio_min = float(min(1.0, quality.min())) if len(quality) else 1.0
    ratio_max = float(max(0.0, quality.max())) if len(quality) else 0.0
    if not total > 0:
        return prob, None, False, ratio_min, ratio_max
    new = (new / total).astype(np.float32)
    ent = np.float32(0)
    for v in new:
        if v > 0:
            ent = np.float32(ent + np.float32(v * np.log(v)))
    return new, float(-ent), True, ratio_min, ratio_max
