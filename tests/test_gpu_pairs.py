"""GPU parity of the read x read SNP agreement counts (hsgpu_pairs_*, hsgpu_read_pair_counts) against the
oracle's restatement of list_similarities_and_differences_between_reads3 (reference
src/separate_reads.cpp:374-433) and against the dense definition sim = 3 A At + R Rt, diff = A Rt + R At.
Bit-exact (int32 counts): tolerance 0."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def snp_columns(oracle, cb, mean_error=0.08):
    """the suspect columns of a synthetic contig as (n_reads, snp_off, read_idx, code, ref_base, second_base)"""
    o = oracle.pileup(cb)
    oc = oracle.call_variants(o["col_off"], o["code"], mean_error)
    pos = oc["suspect_pos"]
    snp_off = np.zeros(pos.size + 1, np.int64)
    idx, code = [np.zeros(0, np.uint32)], [np.zeros(0, np.uint8)]
    for j, q in enumerate(pos):
        a, b = o["col_off"][q], o["col_off"][q + 1]
        idx.append(o["read_idx"][a:b])
        code.append(o["code"][a:b])
        snp_off[j + 1] = snp_off[j] + (b - a)
    return (cb.n_reads, snp_off, np.concatenate(idx).astype(np.uint32), np.concatenate(code).astype(np.uint8),
            oc["ref_base"][pos].astype(np.uint8), oc["second_base"][pos].astype(np.uint8))


def random_columns(rng, n_reads, n_snps, span, sorted_reads):
    """reads cover a window of SNPs each (like reads on a contig); codes: ref / alt / a third allele"""
    first = rng.integers(0, max(1, n_snps - span // 2), n_reads)
    if sorted_reads:
        first.sort()
    length = rng.integers(span // 2, span + 1, n_reads)
    snp_off = np.zeros(n_snps + 1, np.int64)
    idx, code = [], []
    rb = rng.integers(33, 158, n_snps).astype(np.uint8)
    sb = ((rb - 33 + rng.integers(1, 124, n_snps)) % 125 + 33).astype(np.uint8)
    for s in range(n_snps):
        r = np.nonzero((first <= s) & (s < first + length) & (rng.random(n_reads) < 0.9))[0]
        u = rng.random(r.size)
        c = np.where(u < 0.55, rb[s], np.where(u < 0.9, sb[s], np.uint8(33 + (int(rb[s]) + 7) % 125))).astype(np.uint8)
        idx.append(r.astype(np.uint32))
        code.append(c)
        snp_off[s + 1] = snp_off[s] + r.size
    idx = np.concatenate(idx)
    if sorted_reads:
        # renumber the reads by the first SNP they actually have a cell at (what a position-sorted SAM gives)
        seen = np.full(n_reads, n_snps, np.int64)
        np.minimum.at(seen, idx, np.repeat(np.arange(n_snps), np.diff(snp_off)))
        rank = np.empty(n_reads, np.uint32)
        rank[np.argsort(seen, kind="stable")] = np.arange(n_reads, dtype=np.uint32)
        idx = rank[idx]
        for s in range(n_snps):  # columns list their reads in ascending order
            idx[snp_off[s]:snp_off[s + 1]].sort()
    return (n_reads, snp_off, idx, np.concatenate(code), rb, sb)


def dense_definition(col):
    n, snp_off, idx, code, rb, sb = col
    S = snp_off.size - 1
    A, R = np.zeros((n, S), np.int64), np.zeros((n, S), np.int64)
    for s in range(S):
        a, b = snp_off[s], snp_off[s + 1]
        r, c = idx[a:b], code[a:b]
        is_ref = c == rb[s]
        R[r[is_ref], s] = 1
        A[r[(~is_ref) & (c == sb[s])], s] = 1
    sim = 3 * A @ A.T + R @ R.T
    diff = A @ R.T + R @ A.T
    np.fill_diagonal(sim, 0)
    np.fill_diagonal(diff, 0)
    return sim.astype(np.int32), diff.astype(np.int32)


def test_single_contig_all_modes(gpu_ctx, oracle):
    from hairsplitter_b200 import api
    col = snp_columns(oracle, cases.medium_case(seed=101))
    assert col[1].size > 50
    want_sim, want_diff = oracle.read_pair_counts(*col)
    assert want_sim.max() > 0 and want_diff.max() > 0
    sim, diff = gpu_ctx.read_pair_counts(*col)  # the reference-shaped entry point
    assert np.array_equal(sim, want_sim) and np.array_equal(diff, want_diff)
    for flags in (0, api.PAIRS_KEEP_ORDER, api.PAIRS_DENSE, api.PAIRS_DENSE | api.PAIRS_KEEP_ORDER, api.PAIRS_SIMT):
        p = api.Pairs(gpu_ctx, [col], flags)
        p.compute()
        p.compute()  # repeatable
        sim, diff = p.fetch(0)
        assert np.array_equal(sim, want_sim), flags
        assert np.array_equal(diff, want_diff), flags
        p.close()


def test_batch_of_ragged_contigs(gpu_ctx, oracle):
    """several contigs in one launch: fewer reads than one tile, no SNPs, no reads, reads without cells"""
    from hairsplitter_b200 import api
    rng = np.random.default_rng(5)
    cols = [
        snp_columns(oracle, cases.small_case(seed=7, length=5000, depth=30, mean_len=1500)),
        (5, np.zeros(1, np.int64), np.zeros(0, np.uint32), np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(0, np.uint8)),
        random_columns(rng, 130, 300, 60, sorted_reads=False),
        (0, np.zeros(1, np.int64), np.zeros(0, np.uint32), np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(0, np.uint8)),
        random_columns(rng, 400, 140, 30, sorted_reads=True),
        snp_columns(oracle, cases.hifi_case(seed=31), mean_error=0.004),
    ]
    p = api.Pairs(gpu_ctx, cols)
    p.compute()
    for c, col in enumerate(cols):
        sim, diff = p.fetch(c)
        if col[0] == 0:
            assert sim.size == 0
            continue
        want_sim, want_diff = oracle.read_pair_counts(*col)
        assert np.array_equal(sim, want_sim), c
        assert np.array_equal(diff, want_diff), c
    info = p.info()
    assert info["tile_pairs"] <= info["tile_pairs_dense"] and info["kblocks"] <= info["kblocks_dense"]
    p.close()


@pytest.mark.parametrize("sorted_reads", [True, False])
def test_many_tiles_per_cta(gpu_ctx, sorted_reads):
    """more work items than SMs and more SNP blocks than pipeline stages: every barrier phase wraps and the
    TMEM accumulators are reused; checked against the dense definition"""
    from hairsplitter_b200 import api
    rng = np.random.default_rng(17 + sorted_reads)
    col = random_columns(rng, 2400, 700, 260, sorted_reads)
    want_sim, want_diff = dense_definition(col)
    for flags in (api.PAIRS_DENSE, 0):
        p = api.Pairs(gpu_ctx, [col], flags)
        info = p.info()
        assert info["identity"] == int(sorted_reads)
        if flags & api.PAIRS_DENSE:
            assert info["tile_pairs"] == 19 * 20 // 2 and info["kblocks"] == info["kblocks_dense"]
        else:
            assert info["kblocks"] < info["kblocks_dense"] // 2  # the band, not the square
        p.compute()
        sim, diff = p.fetch(0)
        assert np.array_equal(sim, want_sim), flags
        assert np.array_equal(diff, want_diff), flags
        p.close()


def test_argument_errors(gpu_ctx):
    from hairsplitter_b200 import api
    bad = (3, np.array([0, 2], np.int64), np.array([0, 7], np.uint32), np.array([40, 41], np.uint8),
           np.array([40], np.uint8), np.array([41], np.uint8))
    with pytest.raises(api.HsgpuError):
        api.Pairs(gpu_ctx, [bad])
