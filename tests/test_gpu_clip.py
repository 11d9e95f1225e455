"""GPU parity of hsgpu_clip_reads (read clipping / window extraction of modify_GFA, reference
src/create_new_contigs.cpp:383-447) against the oracle and the committed golden vectors of the reference's own loop.
Integer outputs: tolerance 0."""
import gzip
import json
import os
import sys

import numpy as np
import pytest

import cases
from hairsplitter_b200 import api

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def _run(gpu_ctx, reads, items):
    """reads: [(ops, pos)], items: [(read, left, right)]"""
    off = np.zeros(len(reads) + 1, np.int64)
    off[1:] = np.cumsum([len(r[0]) for r in reads])
    cigar = np.concatenate([np.asarray(r[0], np.uint32) for r in reads]) if off[-1] else np.zeros(0, np.uint32)
    return gpu_ctx.clip_reads(cigar, off, [r[1] for r in reads], [i[0] for i in items], [i[1] for i in items], [i[2] for i in items])


def _check(oracle, reads, items, got):
    n_ok = 0
    for (ri, left, right), g in zip(items, got):
        ops, pos = reads[ri]
        st, out = oracle.clip_read(ops, pos, left, right)
        assert int(g["status"]) == st, (ops, pos, left, right)
        if st == 0:
            assert [int(g[k]) for k in ("read_start", "read_end", "cigar_start", "cigar_end")] == [int(x) for x in out], (ops, pos, left, right)
            n_ok += 1
    return n_ok


def test_golden_vectors(gpu_ctx, oracle):
    vec = json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "clip_vectors.json.gz")).read())
    reads = [(r["ops"], r["pos"]) for r in vec]
    items = [(i, r["left"], r["right"]) for i, r in enumerate(vec)]
    got = _run(gpu_ctx, reads, items)
    for r, g in zip(vec, got):
        assert int(g["status"]) == r["status"], r
        if r["status"] == 0:
            assert [int(g[k]) for k in ("read_start", "read_end", "cigar_start", "cigar_end")] == r["out"], r
            want = r["clipped"] if r["out"][3] > r["out"][2] else ""
            assert (api.clipped_cigar(r["ops"], g) if r["out"][3] > r["out"][2] else "") == want, r


def test_windows_of_a_synthetic_contig(gpu_ctx, oracle):
    """every read of a contig against every 2 kb window it overlaps (plus its neighbours): long CIGARs (several
    chunks of 32 ops per warp), clips, reads starting inside / after the interval, hard clips"""
    cb = cases.medium_case(seed=321)
    reads = [(cb.cigar[cb.cigar_off[i]:cb.cigar_off[i + 1]].tolist(), int(cb.start[i])) for i in range(cb.n_reads)]
    ends = [r[1] + sum(o >> 4 for o in r[0] if (o & 15) in (0, 2, 7, 8)) for r in reads]
    items = []
    for i, (ops, pos) in enumerate(reads):
        for w0 in range(max(0, pos - 2000) // 2000 * 2000, ends[i] + 2000, 2000):
            items.append((i, max(0, w0 - 10), min(cb.length - 1, w0 + 2000 + 10)))
    assert len(items) > 1000
    got = _run(gpu_ctx, reads, items)
    assert _check(oracle, reads, items, got) > 300


def test_random_and_degenerate_cases(gpu_ctx, oracle):
    import make_golden_clip
    rng = np.random.default_rng(99)
    reads, items = [], []
    for i in range(4000):
        ops, pos, left, right = make_golden_clip.random_case(rng, realistic=i % 2 == 0)
        reads.append((ops, pos))
        items.append((i, left, right))
    reads.append(([], 5))            # "*": no ops at all
    items.append((len(reads) - 1, 0, 10))
    reads.append(([(70000 << 4) | 0, (3 << 4) | 2, (65000 << 4) | 0], 100))  # ops far longer than a window
    for left, right in ((0, 50), (100, 100), (60000, 61000), (70099, 70103), (140000, 140500), (200000, 200001)):
        items.append((len(reads) - 1, left, right))
    got = _run(gpu_ctx, reads, items)
    assert _check(oracle, reads, items, got) > 500
    assert gpu_ctx.clip_reads(np.zeros(0, np.uint32), np.zeros(1, np.int64), [], [], [], []).size == 0
