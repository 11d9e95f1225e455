"""CPU tests: the oracle (oracle/hs_oracle.c) against the reference itself (oracle/_ref, when built here)
and against the committed golden vectors that were generated from the reference."""
import glob
import os

import numpy as np
import pytest

import cases
from golden import make_golden


def test_oracle_matches_golden_vectors(oracle):
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "cv_*.npz")))
    assert len(files) >= 3
    for f in files:
        g = np.load(f)
        cb = make_golden.batch_from_npz(g)
        p = oracle.pileup(cb)
        assert np.array_equal(p["col_off"], g["col_off"])
        assert np.array_equal(p["read_idx"], g["read_idx"])
        assert np.array_equal(p["code"], g["code"])
        assert np.array_equal(p["read_end"], g["read_end"])
        md = oracle.mean_distance(*p["stats"])
        assert np.float32(md).tobytes() == np.float32(g["mean_distance"]).tobytes()
        assert np.array_equal(oracle.ref_codes(cb.contig), g["newref"])
        oc = oracle.call_variants(p["col_off"], p["code"], md)
        assert np.array_equal(oc["ref_base"], g["ref_base"])
        assert np.array_equal(oc["second_base"], g["second_base"])
        assert np.array_equal(oc["suspect_pos"], g["suspect_pos"])
        assert np.array_equal(oc["suspect_pos"][oc["suspect_is_auto"] == 1], g["automatic_pos"])
        assert np.float32(oc["depth_sum"] / cb.length) == np.float32(g["depth"])
        parts = make_golden.parts_from_npz(g)
        kept = oracle.robust_filter(p["col_off"], p["read_idx"], p["code"], oc["ref_base"], oc["second_base"], parts,
                                    oc["suspect_pos"])
        assert np.array_equal(kept, g["filtered_pos"])


def test_rh_order_and_sort_match_reference(oracle, refcv):
    rng = np.random.default_rng(1)
    for it in range(4000):
        n = int(rng.integers(1, 40 if it % 20 else 130))
        keys = rng.permutation(np.arange(0, 160))[:n].astype(np.uint8)
        assert np.array_equal(oracle.rh_order(keys), refcv.rh_order(keys))
    for it in range(4000):
        n = int(rng.integers(1, 130))
        keys = rng.permutation(256)[:n].astype(np.uint8)
        counts = rng.integers(0, rng.integers(1, 8), n).astype(np.int32)
        a, b = oracle.sort_desc(keys, counts), refcv.sort_desc(keys, counts)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_rh_order_long_probe_chains(oracle, refcv):
    """keys that share a home bucket force displacements >= 6, i.e. robin_hood's info-byte overflow path"""
    M64 = (1 << 64) - 1

    def h(key, k):
        x = (key * 0xff51afd7ed558ccd) & M64
        x ^= x >> 33
        x = (x * ((0xc4ceb9fe1a85ec53 + k * 0xc4ceb9fe1a85ec54) & M64)) & M64
        return x ^ (x >> 33)

    rng = np.random.default_rng(2)
    for k, mask in [(0, 7), (1, 15), (2, 31), (3, 63)]:
        for home in range(mask + 1):
            same = [key for key in range(256) if ((h(key, k) >> 5) & mask) in (home, (home + 1) & mask)]
            for rep in range(20):
                n = int(rng.integers(2, min(len(same), 40) + 1))
                pre = rng.permutation(256)[: int(rng.integers(0, 8))]
                keys = np.array(list(dict.fromkeys(list(pre) + list(rng.permutation(same)[:n]))), dtype=np.uint8)
                assert np.array_equal(oracle.rh_order(keys), refcv.rh_order(keys))


def test_chi_square_matches_reference(oracle, refcv):
    rng = np.random.default_rng(3)
    for n00 in range(0, 7):
        for n01 in range(0, 7):
            for n10 in range(0, 7):
                for n11 in range(0, 7):
                    a, b = oracle.chi_square(n00, n01, n10, n11), refcv.chi_square(n00, n01, n10, n11)
                    assert np.float32(a).tobytes() == np.float32(b).tobytes()
    for _ in range(20000):
        t = rng.integers(0, rng.integers(1, 300), 4)
        a, b = oracle.chi_square(*t), refcv.chi_square(*t)
        assert np.float32(a).tobytes() == np.float32(b).tobytes()


@pytest.mark.parametrize("case", ["small", "small_eqx", "hifi", "deep", "ragged"])
def test_pipeline_matches_reference(oracle, refcv, case):
    cbs = {"small": lambda: [cases.small_case()], "small_eqx": lambda: [cases.small_case(seed=12, eqx=True)],
           "hifi": lambda: [cases.hifi_case()], "deep": lambda: [cases.deep_case()],
           "ragged": cases.ragged_cases}[case]()
    rng = np.random.default_rng(4)
    for cb in cbs:
        R = refcv(cb)
        p, rp = oracle.pileup(cb), R.pileup()
        for k in ("col_off", "read_idx", "code"):
            assert np.array_equal(p[k], rp[k]), k
        md = oracle.mean_distance(*p["stats"])
        assert np.float32(md).tobytes() == np.float32(R.mean_distance()).tobytes()
        assert np.array_equal(R.read_limits()[1], p["read_end"])
        assert np.array_equal(oracle.ref_codes(cb.contig), R.newref())
        rc = R.call_variants()
        oc = oracle.call_variants(p["col_off"], p["code"], md)
        assert np.array_equal(oc["ref_base"], rc["ref_base"]) and np.array_equal(oc["second_base"], rc["second_base"])
        assert np.array_equal(oc["suspect_pos"], rc["suspects"]["pos"])
        assert np.array_equal(oc["suspect_pos"][oc["suspect_is_auto"] == 1], rc["automatic"]["pos"])
        assert np.float32(oc["depth_sum"] / cb.length) == np.float32(rc["depth"])
        parts, filt, merged = R.robust()
        kept = oracle.robust_filter(p["col_off"], p["read_idx"], p["code"], oc["ref_base"], oc["second_base"], parts,
                                    oc["suspect_pos"])
        assert np.array_equal(kept, filt["pos"])
        # distance() on the reference's partitions and on hand-made ones that contain every state
        if cb.n_reads == 0:
            continue
        for q in rng.integers(0, cb.length, 60):
            a, b = p["col_off"][q], p["col_off"][q + 1]
            for pi in range(min(len(parts), 4)):
                P = parts[pi]
                x = oracle.distance(P["read_idx"], P["state"], P["more"], P["less"], p["read_idx"][a:b],
                                    p["code"][a:b], rc["ref_base"][q])
                assert np.array_equal(x, R.distance(pi, int(q), rc["ref_base"][q]))
            idx = np.unique(rng.integers(0, cb.n_reads, max(2, cb.n_reads // 2))).astype(np.int32)
            st = rng.choice(np.array([1, -1, 0, -2], dtype=np.int16), size=idx.size)
            more = rng.integers(0, 6, idx.size).astype(np.int32)
            less = rng.integers(0, 3, idx.size).astype(np.int32)
            for ref_base in (rc["ref_base"][q], rc["second_base"][q], 133, 157):
                x = oracle.distance(idx, st, more, less, p["read_idx"][a:b], p["code"][a:b], ref_base)
                assert np.array_equal(x, R.distance_custom(idx, st, more, less, int(q), ref_base))


def test_deletion_majority_codes_hit_the_signed_char_quirk(oracle, refcv):
    """ref_base >= 128 (deletion is the consensus): `char != unsigned char` is always true (:838)"""
    cb = cases.small_case(seed=13, indel_frac=0.9, n_strains=2)
    R = refcv(cb)
    p = oracle.pileup(cb)
    rc = R.call_variants()
    big = np.nonzero(rc["ref_base"] >= 128)[0]
    rng = np.random.default_rng(5)
    idx = np.arange(cb.n_reads, dtype=np.int32)
    n = 0
    for q in list(big[:50]) + list(rng.integers(0, cb.length, 50)):
        a, b = p["col_off"][q], p["col_off"][q + 1]
        st = rng.choice(np.array([1, -1, 0, -2], dtype=np.int16), size=idx.size)
        x = oracle.distance(idx, st, np.ones_like(idx) * 3, np.zeros_like(idx), p["read_idx"][a:b], p["code"][a:b],
                            rc["ref_base"][q])
        assert np.array_equal(x, R.distance_custom(idx, st, np.ones_like(idx) * 3, np.zeros_like(idx), int(q),
                                                   rc["ref_base"][q]))
        n += 1
    assert n >= 50


def test_read_pair_counts_definition(oracle):
    """hso_read_pair_counts against the dense definition sim = 3 A At + R Rt, diff = A Rt + R At"""
    rng = np.random.default_rng(6)
    R_, S = 40, 25
    snp_off = np.zeros(S + 1, np.int64)
    idx, code = [], []
    rb = rng.integers(33, 158, S).astype(np.uint8)
    sb = ((rb - 33 + rng.integers(1, 125, S)) % 125 + 33).astype(np.uint8)
    A = np.zeros((R_, S), np.int64)
    Rm = np.zeros((R_, S), np.int64)
    for s in range(S):
        reads = np.sort(rng.choice(R_, size=int(rng.integers(0, R_)), replace=False))
        c = rng.choice([rb[s], sb[s], 40], size=reads.size, p=[0.6, 0.3, 0.1]).astype(np.uint8)
        idx.append(reads.astype(np.uint32))
        code.append(c)
        snp_off[s + 1] = snp_off[s] + reads.size
        Rm[reads[c == rb[s]], s] = 1
        A[reads[(c == sb[s]) & (c != rb[s])], s] = 1
    sim, diff = oracle.read_pair_counts(R_, snp_off, np.concatenate(idx), np.concatenate(code), rb, sb)
    ws = 3 * A @ A.T + Rm @ Rm.T
    wd = A @ Rm.T + Rm @ A.T
    np.fill_diagonal(ws, 0)
    np.fill_diagonal(wd, 0)
    assert np.array_equal(sim, ws) and np.array_equal(diff, wd)
    # ... and against the reference's own Eigen implementation (oracle/_ref/libhsref_sr.so)
    from oracle import pyoracle
    if pyoracle.RefSR.available():
        rs, rd = pyoracle.RefSR.read_pair_counts(R_, snp_off, np.concatenate(idx), np.concatenate(code), rb, sb)
        assert np.array_equal(sim, rs) and np.array_equal(diff, rd)
        col = cases_snp_columns(oracle)
        a = oracle.read_pair_counts(*col)
        b = pyoracle.RefSR.read_pair_counts(*col)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and b[0].max() > 0


def cases_snp_columns(oracle):
    """suspect columns of a synthetic contig in the .col layout (same helper as the GPU pair tests)"""
    o = oracle.pileup(cases.medium_case(seed=101))
    oc = oracle.call_variants(o["col_off"], o["code"], 0.08)
    pos = oc["suspect_pos"]
    snp_off = np.zeros(pos.size + 1, np.int64)
    idx, code = [], []
    for j, q in enumerate(pos):
        a, b = o["col_off"][q], o["col_off"][q + 1]
        idx.append(o["read_idx"][a:b])
        code.append(o["code"][a:b])
        snp_off[j + 1] = snp_off[j] + (b - a)
    return (o["read_end"].size, snp_off, np.concatenate(idx).astype(np.uint32), np.concatenate(code).astype(np.uint8),
            oc["ref_base"][pos].astype(np.uint8), oc["second_base"][pos].astype(np.uint8))


def test_splitting_long_cigar_ops_does_not_change_the_pileup(oracle):
    """the 16-bit CIGAR form of the C ABI splits ops longer than 4095; the oracle must not see a difference"""
    import copy
    from hairsplitter_b200 import api
    cb = cases.walk_edge_case()
    long_cb = copy.deepcopy(cb)
    # stretch one read's match so that it needs splitting, on a longer contig
    c16, off16 = api.compact_cigar(np.array([(9000 << 4) | 0, (5000 << 4) | 2, (4096 << 4) | 1, (3 << 4) | 0], np.uint32),
                                   np.array([0, 4], np.int64))
    assert [(int(x) >> 4, int(x) & 15) for x in c16] == [(4095, 0), (4095, 0), (810, 0), (4095, 2), (905, 2), (4095, 1), (1, 1), (3, 0)]
    assert list(off16) == [0, 8]
    split, split_off = api.compact_cigar(cb.cigar, cb.cigar_off)
    long_cb.cigar = split.astype(np.uint32)
    long_cb.cigar_off = split_off
    a, b = oracle.pileup(cb), oracle.pileup(long_cb)
    for k in ("col_off", "read_idx", "code", "read_end"):
        assert np.array_equal(a[k], b[k])
    assert list(a["stats"]) == list(b["stats"])


def test_columns_deeper_than_a_short_are_counted_like_the_reference(oracle, refcv):
    """call_variants counts a column with a `short` index (src/call_variants.cpp:479): only its first 32768 cells ever
    reach the histogram and the depth. A 40000-deep column whose variant is carried by the reads beyond the 32768th must
    stay invisible, one carried by early reads must be called; depth = 32768 per column. The oracle follows the
    reference here. The CUDA library counts every cell of such a column -- the one known deviation, DESIGN.md 7b (no
    BASELINE config comes near: the amplicon contig is 2000x)."""
    from hairsplitter_b200 import synth
    L, R, counted = 40, 40000, 32768
    rng = np.random.default_rng(0)
    contig = rng.integers(0, 4, L).astype(np.uint8)
    reads = np.tile(contig, (R, 1))
    reads[counted:, 20] = (contig[20] + 1) % 4   # seen by nobody: these cells are never counted
    reads[:10, 10] = (contig[10] + 2) % 4        # ten early reads: counted
    cb = synth.ContigBatch(contig=contig, read_bases=reads.reshape(-1), read_off=np.arange(R + 1, dtype=np.int64) * L,
                           cigar=np.full(R, (L << 4) | 0, np.uint32), cigar_off=np.arange(R + 1, dtype=np.int64),
                           start=np.zeros(R, np.int32), strand=np.ones(R, np.uint8), strain=np.zeros(R, np.int32))
    ref = refcv(cb)
    md = ref.mean_distance()
    rv = ref.call_variants(md)
    o = oracle.pileup(cb)
    ov = oracle.call_variants(o["col_off"], o["code"], md)
    assert np.array_equal(rv["ref_base"], ov["ref_base"]) and np.array_equal(rv["second_base"], ov["second_base"])
    assert np.array_equal(rv["suspects"]["pos"], ov["suspect_pos"]) and ov["suspect_pos"].size >= 1
    assert rv["depth"] == counted and ov["depth_sum"] == counted * L
    assert ov["second_base"][20] == 0  # the late variant never entered the histogram: rank 1 is one of the three filler keys
    ref.close()
