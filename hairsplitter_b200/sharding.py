"""Contig-chunk sharding over ranks: chunks are independent, so ranks take whole
chunks -- heaviest first onto the least loaded rank -- and the data path has no collective; only the
timing (max over ranks) and the unit counts (sum) are reduced. The HS_call_variants executable shards
contigs over GPUs with the same rule (hairsplitter_b200/host/call_variants_main.cpp)."""
from __future__ import annotations


def lpt_assign(weights, n_bins: int):
    """longest-processing-time-first: returns n_bins lists of item indices (each ascending)"""
    order = sorted(range(len(weights)), key=lambda i: (-weights[i], i))
    load = [0.0] * n_bins
    bins = [[] for _ in range(n_bins)]
    for i in order:
        g = min(range(n_bins), key=lambda b: (load[b], b))
        bins[g].append(i)
        load[g] += weights[i]
    return [sorted(b) for b in bins]


def chunk_weight(chunk) -> float:
    """columns + aligned cells, the quantity both hot kernels scale with"""
    return float(chunk.length) + float(chunk.read_off[-1])


def reduce_step(ms: float, units: float, device=None):
    """(max over ranks of the step time, sum over ranks of the units processed); no-op without a process group"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ms, units
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    u = torch.tensor([units], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])
