"""N > 1 host logic on CPU (gloo, world size 2): chunks are sharded with no data-path collective, every
chunk is processed by exactly one rank, the union of the per-rank results equals the single-process
result, and the step metric is reduced as max(time) / sum(units)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from hairsplitter_b200 import sharding  # noqa: E402


def _chunks():
    import cases
    return [cases.small_case(seed=s, length=ln, depth=d, mean_len=400) for s, ln, d in
            [(1, 3000, 20), (2, 800, 10), (3, 1500, 25), (4, 600, 8), (5, 2200, 12)]]


def _digest(oracle, cb):
    o = oracle.pileup(cb)
    return (int(o["code"].shape[0]), int(o["stats"][0]), int(o["stats"][1]), int(np.bitwise_xor.reduce(o["code"].astype(np.int64) * 31 + 7)))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.pyoracle import Oracle
    chunks = _chunks()
    mine = sharding.lpt_assign([sharding.chunk_weight(c) for c in chunks], world)[rank]
    O = Oracle()
    local = {i: _digest(O, chunks[i]) for i in mine}   # the stand-in for the per-rank GPU work
    gathered = [None] * world
    dist.all_gather_object(gathered, local)              # host-side gather of results, not a data-path collective
    ms, units = sharding.reduce_step(10.0 + 5.0 * rank, float(sum(chunks[i].length for i in mine)))
    if rank == 0:
        out.put((gathered, ms, units))
    dist.barrier()
    dist.destroy_process_group()


def test_lpt_assign_is_a_partition_and_balances():
    w = [9, 1, 8, 2, 7, 3, 6, 4, 5, 5]
    bins = sharding.lpt_assign(w, 3)
    assert sorted(i for b in bins for i in b) == list(range(len(w)))
    loads = [sum(w[i] for i in b) for b in bins]
    assert max(loads) - min(loads) <= max(w)
    assert sharding.lpt_assign(w, 1) == [list(range(len(w)))]
    assert sharding.lpt_assign([], 2) == [[], []]


@pytest.mark.timeout(300)
def test_two_ranks_shard_chunks_without_collectives_on_the_data_path():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, ms, units = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle.pyoracle import Oracle
    chunks = _chunks()
    O = Oracle()
    want = {i: _digest(O, c) for i, c in enumerate(chunks)}
    got = {}
    for part in gathered:
        for i, d in part.items():
            assert i not in got, "a chunk was processed by two ranks"
            got[i] = d
    assert got == want
    assert ms == 15.0                                   # max over ranks
    assert units == float(sum(c.length for c in chunks))  # sum over ranks
