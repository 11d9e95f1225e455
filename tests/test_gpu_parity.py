"""GPU parity tests: libhsgpu.so (CUDA, through the C ABI) against the oracle on identical seeded inputs.

Bit-exact integer/byte comparison everywhere; the chi-square float is compared bit-for-bit too.
Where oracle/_ref (the compiled reference itself) travelled with the snapshot it is used as a second
witness. Nothing here reads /root/reference.
"""
import numpy as np
import pytest

import cases
from hairsplitter_b200 import api

pytestmark = pytest.mark.gpu


def _build(gpu_ctx, chunks, mean_error=None):
    pk = api.PackedBatch(chunks)
    pu = api.Pileup(gpu_ctx, pk)
    pu.build()
    pu.column_rank(mean_error)
    return pk, pu


def _check_contig(oracle, pu, ci, cb):
    o = oracle.pileup(cb)
    cells, dist, alen = pu.stats()
    assert int(cells[ci]) == o["code"].shape[0]
    assert int(dist[ci]) == int(o["stats"][0]) and int(alen[ci]) == int(o["stats"][1])
    md = pu.mean_distance(dist[ci], alen[ci])
    assert md == oracle.mean_distance(*o["stats"])
    e = pu.export(ci)
    assert np.array_equal(e["col_off"], o["col_off"])
    assert np.array_equal(e["read_idx"], o["read_idx"])
    assert np.array_equal(e["code"], o["code"])
    oc = oracle.call_variants(o["col_off"], o["code"], md)
    s = pu.column_summary(ci)
    assert np.array_equal(s["ref_base"], oc["ref_base"])
    assert np.array_equal(s["second_base"], oc["second_base"])
    assert np.array_equal(s["depth"].astype(np.int64), np.diff(o["col_off"]))
    pos, au = pu.suspects(ci)
    assert np.array_equal(pos, oc["suspect_pos"])
    assert np.array_equal(au, oc["suspect_is_auto"])
    ns, ds = pu.column_counts()
    assert int(ds[ci]) == oc["depth_sum"]
    return o, oc, md


@pytest.mark.parametrize("case", ["small", "small_eqx", "medium", "hifi", "deep", "walk_edges"])
def test_pileup_and_ranking_match_oracle(gpu_ctx, oracle, case):
    cb = {"small": cases.small_case, "small_eqx": lambda: cases.small_case(seed=12, eqx=True),
          "medium": cases.medium_case, "hifi": cases.hifi_case, "deep": cases.deep_case,
          "walk_edges": cases.walk_edge_case}[case]()
    pk, pu = _build(gpu_ctx, [cb])
    o, oc, md = _check_contig(oracle, pu, 0, cb)
    off = 0
    ends = pu.read_ends()
    assert np.array_equal(ends, o["read_end"])
    # column extraction of the suspect columns (what the .col writer needs)
    pos = oc["suspect_pos"]
    eo, ei, ec = pu.extract_columns(0, pos)
    for j, q in enumerate(pos):
        a, b = o["col_off"][q], o["col_off"][q + 1]
        assert np.array_equal(ei[eo[j]:eo[j + 1]], o["read_idx"][a:b])
        assert np.array_equal(ec[eo[j]:eo[j + 1]], o["code"][a:b])
    pu.close()


def test_batch_of_ragged_contigs(gpu_ctx, oracle):
    chunks = cases.ragged_cases() + [cases.small_case(seed=61), cases.walk_edge_case(seed=72, length=3100)]
    pk, pu = _build(gpu_ctx, chunks)
    for ci, cb in enumerate(chunks):
        _check_contig(oracle, pu, ci, cb)
    pu.close()


def test_mean_error_override_switches_min_reads(gpu_ctx, oracle):
    cb = cases.small_case(seed=71, error=0.03)
    for me in (0.01, 0.2):
        pk, pu = _build(gpu_ctx, [cb], mean_error=[me])
        o = oracle.pileup(cb)
        oc = oracle.call_variants(o["col_off"], o["code"], me)
        pos, au = pu.suspects(0)
        assert np.array_equal(pos, oc["suspect_pos"]) and np.array_equal(au, oc["suspect_is_auto"])
        pu.close()


def _random_partitions(rng, cb, o, n_parts):
    """partitions with every state (1,-1,0,-2), built from real columns so that they overlap"""
    parts = []
    R = cb.n_reads
    for _ in range(n_parts):
        q = int(rng.integers(0, cb.length))
        a, b = o["col_off"][q], o["col_off"][min(cb.length, q + int(rng.integers(1, 4000)))]
        idx = np.unique(o["read_idx"][a:b]).astype(np.int32)
        if idx.size == 0:
            idx = np.array([int(rng.integers(0, max(R, 1)))], dtype=np.int32)
        idx = idx[rng.random(idx.size) < 0.8] if idx.size > 3 else idx
        st = rng.choice(np.array([1, -1, 0, -2], dtype=np.int16), size=idx.size, p=[0.45, 0.35, 0.12, 0.08])
        more = rng.integers(0, 6, idx.size).astype(np.int32)
        less = rng.integers(0, 3, idx.size).astype(np.int32)
        parts.append(dict(read_idx=idx, state=st, more=more, less=less))
    return parts


def _check_tables(oracle, pu, cb, o, oc, parts, pos):
    t = pu.partition_tables(0, parts, pos)
    for j, q in enumerate(pos):
        a, b = o["col_off"][q], o["col_off"][q + 1]
        for pi, P in enumerate(parts):
            d = oracle.distance(P["read_idx"], P["state"], P["more"], P["less"], o["read_idx"][a:b], o["code"][a:b],
                                oc["ref_base"][q])
            g = t[j, pi]
            got = [g["n00"], g["n01"], g["n10"], g["n11"], g["solid00"], g["solid01"], g["solid10"], g["solid11"],
                   g["second_base"], g["augmented"]]
            assert list(d) == [int(x) for x in got], (q, pi)
            chi = oracle.chi_square(d[0], d[1], d[2], d[3])
            assert np.float32(chi).tobytes() == np.float32(g["chi_square"]).tobytes(), (q, pi, d)


@pytest.mark.parametrize("case", ["small", "deep"])
def test_partition_tables_match_oracle(gpu_ctx, oracle, case):
    cb = cases.small_case(seed=81) if case == "small" else cases.deep_case(seed=82)
    pk, pu = _build(gpu_ctx, [cb])
    o, oc, md = _check_contig(oracle, pu, 0, cb)
    rng = np.random.default_rng(5)
    parts = _random_partitions(rng, cb, o, 9 if case == "small" else 3)
    pos = np.unique(np.concatenate([rng.integers(0, cb.length, 150), oc["suspect_pos"][:40]])).astype(np.int32)
    _check_tables(oracle, pu, cb, o, oc, parts, pos)
    pu.close()


def test_many_partitions_tables(gpu_ctx, oracle):
    """more than 128 partitions: exercises the partition-chunk loop"""
    cb = cases.small_case(seed=83, length=3000, depth=25)
    pk, pu = _build(gpu_ctx, [cb])
    o, oc, md = _check_contig(oracle, pu, 0, cb)
    rng = np.random.default_rng(6)
    parts = _random_partitions(rng, cb, o, 140)
    pos = np.unique(rng.integers(0, cb.length, 25)).astype(np.int32)
    _check_tables(oracle, pu, cb, o, oc, parts, pos)
    pu.close()


@pytest.mark.parametrize("case", ["small", "medium", "hifi", "deep", "many_parts", "over128_parts", "noisy"])
def test_robust_filter_matches_oracle(gpu_ctx, oracle, case):
    cb = {"small": lambda: cases.small_case(seed=91), "medium": lambda: cases.medium_case(seed=92),
          "hifi": lambda: cases.hifi_case(seed=93), "deep": lambda: cases.deep_case(seed=94),
          "many_parts": lambda: cases.small_case(seed=95),
          "over128_parts": lambda: cases.small_case(seed=96, length=4000, depth=40),
          # 35 % error at depth ~100: a fifth of the columns have more than 24 distinct codes (up to 45), which the
          # lane-per-partition kernel hands over to the lane-per-read kernel
          "noisy": lambda: cases.small_case(seed=97, length=3000, depth=110, mean_len=1500, error=0.35)}[case]()
    pk, pu = _build(gpu_ctx, [cb])
    o, oc, md = _check_contig(oracle, pu, 0, cb)
    rng = np.random.default_rng(7)
    parts = None
    from oracle import pyoracle
    if pyoracle.ref_available() and case not in ("many_parts", "over128_parts", "noisy"):
        R = pyoracle.RefCV(cb)
        rc = R.call_variants()
        assert np.array_equal(rc["suspects"]["pos"], oc["suspect_pos"])
        parts, filt, merged = R.robust()
        kept = pu.robust_filter(0, parts, oc["suspect_pos"])
        assert np.array_equal(kept, filt["pos"])  # the reference's own snps_out
    if parts is None or len(parts) == 0:
        # over128_parts: the kernel takes partitions 128 at a time (presence masks, transposed state rows of 160)
        parts = _random_partitions(rng, cb, o, {"many_parts": 70, "over128_parts": 150, "noisy": 40}.get(case, 12))
    want = oracle.robust_filter(o["col_off"], o["read_idx"], o["code"], oc["ref_base"], oc["second_base"], parts,
                                oc["suspect_pos"])
    kept = pu.robust_filter(0, parts, oc["suspect_pos"])
    assert np.array_equal(kept, want)
    # no partitions: nothing is kept (:640-642)
    assert pu.robust_filter(0, [], oc["suspect_pos"]).size == 0
    pu.close()


def test_read_pair_counts_match_oracle(gpu_ctx, oracle):
    cb = cases.medium_case(seed=101)
    o = oracle.pileup(cb)
    oc = oracle.call_variants(o["col_off"], o["code"], 0.08)
    pos = oc["suspect_pos"]
    assert pos.size > 50
    snp_off = np.zeros(pos.size + 1, np.int64)
    idx, code = [], []
    for j, q in enumerate(pos):
        a, b = o["col_off"][q], o["col_off"][q + 1]
        idx.append(o["read_idx"][a:b])
        code.append(o["code"][a:b])
        snp_off[j + 1] = snp_off[j] + (b - a)
    idx, code = np.concatenate(idx), np.concatenate(code)
    rb, sb = oc["ref_base"][pos], oc["second_base"][pos]
    want_sim, want_diff = oracle.read_pair_counts(cb.n_reads, snp_off, idx, code, rb, sb)
    sim, diff = gpu_ctx.read_pair_counts(cb.n_reads, snp_off, idx, code, rb, sb)
    assert np.array_equal(sim, want_sim) and np.array_equal(diff, want_diff)
    assert want_sim.max() > 0 and want_diff.max() > 0
    # degenerate inputs
    sim, diff = gpu_ctx.read_pair_counts(5, np.zeros(1, np.int64), np.zeros(0, np.uint32), np.zeros(0, np.uint8),
                                         np.zeros(0, np.uint8), np.zeros(0, np.uint8))
    assert not sim.any() and not diff.any()


def test_golden_fixtures(gpu_ctx):
    """outputs of the compiled reference itself, committed under tests/golden/"""
    import glob
    import os
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "cv_*.npz")))
    assert files, "golden fixtures missing"
    from golden import make_golden
    for f in files:
        g = np.load(f)
        cb = make_golden.batch_from_npz(g)
        pk, pu = _build(gpu_ctx, [cb])
        cells, dist, alen = pu.stats()
        assert np.float32(pu.mean_distance(dist[0], alen[0])).tobytes() == np.float32(g["mean_distance"]).tobytes()
        e = pu.export(0)
        assert np.array_equal(e["col_off"], g["col_off"]) and np.array_equal(e["read_idx"], g["read_idx"])
        assert np.array_equal(e["code"], g["code"])
        s = pu.column_summary(0)
        assert np.array_equal(s["ref_base"], g["ref_base"]) and np.array_equal(s["second_base"], g["second_base"])
        pos, au = pu.suspects(0)
        assert np.array_equal(pos, g["suspect_pos"])
        assert np.array_equal(pos[au == 1], g["automatic_pos"])
        parts = make_golden.parts_from_npz(g)
        kept = pu.robust_filter(0, parts, pos)
        assert np.array_equal(kept, g["filtered_pos"])
        pu.close()


@pytest.mark.parametrize("case", ["hifi", "walk_edges"])
def test_compact_cigar_input_gives_the_same_pileup(gpu_ctx, oracle, case):
    """hsgpu_pileup_input.cigar16 (ops longer than 4095 split) is only another encoding of the same alignment"""
    cb = {"hifi": cases.hifi_case, "walk_edges": cases.walk_edge_case}[case]()
    pk = api.PackedBatch([cb]).use_compact_cigar()
    assert pk.cigar16.dtype == np.uint16
    if case == "walk_edges":  # ops longer than 4095 are split
        assert pk.cigar16.shape[0] > pk.cigar.shape[0]
    pu = api.Pileup(gpu_ctx, pk)
    pu.build()
    pu.column_rank()
    _check_contig(oracle, pu, 0, cb)
    pu.close()


@pytest.mark.parametrize("case", ["hifi", "walk_edges", "clips"])
def test_cigar8_input_gives_the_same_pileup(gpu_ctx, oracle, case):
    """hsgpu_pileup_input.cigar8 (M/=/X, I, D, S/H classes; ops longer than 63 split) encodes the same alignment"""
    if case == "hifi":
        cb = cases.hifi_case()
    elif case == "clips":
        cb = cases.small_case(seed=91, hard=0.6, eqx=True)  # H and S clips, =/X instead of M
    else:
        cb = cases.walk_edge_case()
        ty = cb.cigar & 15
        keep = (ty != 3) & (ty != 6)  # N and P have no 8-bit form; drop them from the hand-made CIGARs
        cb.cigar_off = np.concatenate([[0], np.cumsum(keep)])[cb.cigar_off].astype(np.int64)
        cb.cigar = cb.cigar[keep]
    pk = api.PackedBatch([cb]).use_cigar8()
    assert pk.cigar8.dtype == np.uint8 and pk.cigar8.shape[0] >= pk.cigar.shape[0]
    assert int(pk.cigar8_off[-1]) == pk.cigar8.shape[0] or pk.cigar.shape[0] == 0
    pu = api.Pileup(gpu_ctx, pk)
    pu.build()
    pu.column_rank()
    _check_contig(oracle, pu, 0, cb)
    pu.close()


def test_suspects_all_equals_the_per_contig_calls(gpu_ctx):
    """hsgpu_suspects_all: every contig's list (and the depth numerators) in one call, ragged batch incl. an empty contig"""
    batch = [cases.small_case(seed=21, length=9000, depth=40, mean_len=1500),
             cases.small_case(seed=22, length=300, depth=0, mean_len=200),
             cases.small_case(seed=23, length=20000, depth=25, mean_len=3000),
             cases.small_case(seed=24, length=4097, depth=60, mean_len=900)]
    pu = api.Pileup(gpu_ctx, api.PackedBatch(batch))
    pu.build()
    pu.column_rank()
    pos, au, off, ds = pu.suspects_all()
    ns, ds1 = pu.column_counts()
    assert np.array_equal(np.diff(off), ns) and np.array_equal(ds, ds1) and off[0] == 0
    assert int(ns.sum()) > 0
    for ci in range(len(batch)):
        p1, a1 = pu.suspects(ci)
        assert np.array_equal(pos[off[ci]:off[ci + 1]], p1) and np.array_equal(au[off[ci]:off[ci + 1]], a1)
    # a buffer that is too small: HSGPU_ERR_CAPACITY, with the offsets filled in so the call can be retried
    off2 = np.zeros(len(batch) + 1, np.int64)
    small = np.zeros(1, np.int32)
    rc = pu.lib.hsgpu_suspects_all(pu.h, 1, small.ctypes.data, None, off2.ctypes.data, None)
    assert rc == -4 and np.array_equal(off2, off)
    # only the counts
    rc = pu.lib.hsgpu_suspects_all(pu.h, 0, None, None, off2.ctypes.data, None)
    assert rc == 0 and np.array_equal(off2, off)
    pu.close()


def test_invalid_batch_is_refused_and_leaves_the_context_usable(gpu_ctx, oracle):
    """hsgpu_pileup_create validates the batch (error code + message, nothing leaked) and the context keeps working"""
    cb = cases.small_case(seed=97, length=2000, depth=10, mean_len=600)
    pk = api.PackedBatch([cb])
    bad = pk.read_start.copy()
    bad[0] = -5
    good, pk.read_start = pk.read_start, bad
    with pytest.raises(api.HsgpuError) as e:
        api.Pileup(gpu_ctx, pk)
    assert "negative read start" in str(e.value)
    pk.read_start = good
    pu = api.Pileup(gpu_ctx, pk)
    pu.build()
    pu.column_rank()
    _check_contig(oracle, pu, 0, cb)
    pu.close()


def test_robust_filter_all_contigs_in_one_launch(gpu_ctx, oracle):
    """hsgpu_partitions_set + hsgpu_robust_filter_all: loops 3+4 of keep_only_robust_variants for a whole batch
    (a deep amplicon-like contig, a contig without reads, one without partitions, > 128 partitions), snps_in =
    the pileup's own suspect columns; against the oracle contig by contig, and against the reference's own
    partitions / snps_out where the compiled reference travelled"""
    from oracle import pyoracle
    chunks = [cases.small_case(seed=191), cases.ragged_cases()[1], cases.medium_case(seed=192), cases.deep_case(seed=193),
              cases.hifi_case(seed=194), cases.small_case(seed=195, length=4000, depth=40), cases.small_case(seed=196)]
    pk, pu = _build(gpu_ctx, chunks)
    rng = np.random.default_rng(8)
    all_parts, want = [], []
    for ci, cb in enumerate(chunks):
        if cb.n_reads == 0:
            all_parts.append([])
            want.append(np.zeros(0, np.int32))
            continue
        o = oracle.pileup(cb)
        _, dist, alen = pu.stats()
        oc = oracle.call_variants(o["col_off"], o["code"], pu.mean_distance(dist[ci], alen[ci]))
        parts = None
        if ci == 6:
            parts = []  # no partition: nothing is kept (:640-642)
        elif ci == 5:
            parts = _random_partitions(rng, cb, o, 150)
        elif pyoracle.ref_available():
            R = pyoracle.RefCV(cb)
            R.call_variants()
            parts, filt, _ = R.robust()
            if len(parts):
                want_ref = filt["pos"]
        if parts is None or (len(parts) == 0 and ci not in (6,)):
            parts = _random_partitions(rng, cb, o, 10)
        all_parts.append(parts)
        w = oracle.robust_filter(o["col_off"], o["read_idx"], o["code"], oc["ref_base"], oc["second_base"], parts,
                                 oc["suspect_pos"]) if len(parts) else np.zeros(0, np.int32)
        want.append(w)
    pu.partitions_set(all_parts)
    for _ in range(2):  # the call may be repeated on resident partitions
        kept, off = pu.robust_filter_all()
        assert off.shape[0] == len(chunks) + 1
        for ci in range(len(chunks)):
            assert np.array_equal(kept[off[ci]:off[ci + 1]], want[ci]), ci
    assert sum(w.size for w in want) > 0
    # the single-contig call gives the same list and leaves the batch's partitions replaced: set them again
    one = pu.robust_filter(2, all_parts[2], pu.suspects(2)[0])
    assert np.array_equal(one, want[2])
    pu.close()
