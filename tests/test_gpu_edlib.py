"""GPU parity of hsgpu_edlib_align_batch against the edlib oracle (oracle/hs_oracle_edlib.c, pinned to the
vendored edlib) and, where it travelled with the snapshot, against the vendored edlib itself."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ALPHA = b"ACGT"


def _rnd(rng, n, alpha=ALPHA):
    return bytes(rng.choice(list(alpha), n).tolist())


def _mutate(rng, s, e):
    out = bytearray()
    for c in s:
        u = rng.random()
        if u < e / 3:
            out.append(int(rng.choice(list(ALPHA))))
        elif u < 2 * e / 3:
            pass
        elif u < e:
            out.append(c)
            out.append(int(rng.choice(list(ALPHA))))
        else:
            out.append(c)
    return bytes(out)


def _check_batch(gpu_ctx, oracle, queries, targets, k, mode, task, use_ref=True):
    from oracle import pyoracle
    res, ends, starts, aln = gpu_ctx.edlib_align_batch(queries, targets, k=k, mode=mode, task=task)
    ref = pyoracle.RefEdlib if (use_ref and pyoracle.RefEdlib.available()) else None
    for i, (q, t) in enumerate(zip(queries, targets)):
        want = oracle.edlib_align(q, t, k, mode, task)
        r = res[i]
        ctx = (i, len(q), len(t), k, mode, task)
        assert int(r["edit_distance"]) == want["edit_distance"], ctx
        assert int(r["alphabet_length"]) == want["alphabet_length"], ctx
        lo, nl = int(r["loc_off"]), int(r["n_locations"])
        assert np.array_equal(ends[lo:lo + nl], want["end_locations"]), ctx
        if want["start_locations"] is not None:
            assert int(r["has_start_locations"]) == 1, ctx
            assert np.array_equal(starts[lo:lo + nl], want["start_locations"]), ctx
        else:
            assert int(r["has_start_locations"]) == 0, ctx
        if task == 2 and want["edit_distance"] >= 0 and len(q) and len(t):
            assert int(r["status"]) == want["status"], ctx
            if want["status"] == 0:
                ao, al = int(r["aln_off"]), int(r["alignment_length"])
                assert np.array_equal(aln[ao:ao + al], want["alignment"]), ctx
        if ref is not None:
            b = ref.align(q, t, k, mode, task)
            assert int(r["edit_distance"]) == b["edit_distance"], ctx
            assert np.array_equal(ends[lo:lo + nl], b["end_locations"]), ctx
            if b["start_locations"] is not None:
                assert np.array_equal(starts[lo:lo + nl], b["start_locations"]), ctx
            if task == 2 and b["edit_distance"] >= 0 and int(r["status"]) == 0 and len(q) and len(t):
                ao, al = int(r["aln_off"]), int(r["alignment_length"])
                assert np.array_equal(aln[ao:ao + al], b["alignment"]), ctx


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("task", [0, 1, 2])
def test_random_pairs_all_modes(gpu_ctx, oracle, mode, task):
    rng = np.random.default_rng(100 + 10 * mode + task)
    for k in (-1, 0, 5, 40, 1000):
        qs, ts = [], []
        for it in range(60):
            qlen = int(rng.choice([0, 1, 5, 30, 63, 64, 65, 100, 127, 128, 129, 200, 300, 640]))
            tlen = int(rng.choice([0, 1, 10, 64, 150, 400, 700]))
            t = _rnd(rng, tlen)
            if rng.random() < 0.7 and tlen > qlen > 0:
                s = int(rng.integers(0, tlen - qlen + 1))
                q = _mutate(rng, t[s:s + qlen], float(rng.choice([0, 0.05, 0.2])))
            else:
                q = _rnd(rng, qlen, ALPHA if rng.random() < 0.8 else b"ACGTNRY-*")
            qs.append(q)
            ts.append(t)
        _check_batch(gpu_ctx, oracle, qs, ts, k, mode, task)


def test_in_pipeline_shape_hw_path(gpu_ctx, oracle):
    """the four call sites of the reference: <=300 bp query, HW, PATH, ~2.3 kb target
    (src/create_new_contigs.cpp:557-630, src/tools.cpp:508-536)"""
    rng = np.random.default_rng(7)
    qs, ts = [], []
    for it in range(40):
        t = _rnd(rng, 2300)
        s = int(rng.integers(0, 2000))
        q = _mutate(rng, t[s:s + int(rng.choice([200, 300]))], float(rng.choice([0.0, 0.05, 0.15])))
        if it % 10 == 0:
            q = _rnd(rng, 300)  # unrelated query: distance close to the query length, many tied locations
        qs.append(q)
        ts.append(t)
    _check_batch(gpu_ctx, oracle, qs, ts, -1, 2, 2)


def test_benchmark_shape_read_chunk_on_window(gpu_ctx, oracle):
    """1536-base read chunk on its contig window + 15 % slack, 10 % error"""
    rng = np.random.default_rng(8)
    qs, ts = [], []
    for it in range(12):
        t = _rnd(rng, 1766)
        q = _mutate(rng, t[115:115 + 1536], 0.10)[:1536]
        qs.append(q)
        ts.append(t)
    _check_batch(gpu_ctx, oracle, qs, ts, -1, 2, 2, use_ref=True)
    _check_batch(gpu_ctx, oracle, qs, ts, -1, 0, 2, use_ref=True)


def _check_vectors(gpu_ctx, vec):
    groups = {}
    for v in vec:
        groups.setdefault((v["k"], v["mode"], v["task"]), []).append(v)
    for (k, mode, task), vs in groups.items():
        qs = [v["q"].encode("latin1") for v in vs]
        ts = [v["t"].encode("latin1") for v in vs]
        res, ends, starts, aln = gpu_ctx.edlib_align_batch(qs, ts, k=k, mode=mode, task=task)
        for i, v in enumerate(vs):
            r = res[i]
            ctx = (i, len(qs[i]), len(ts[i]), k, mode, task)
            assert int(r["status"]) == v["status"], ctx
            assert int(r["edit_distance"]) == v["edit_distance"], ctx
            assert int(r["alphabet_length"]) == v["alphabet_length"], ctx
            lo, nl = int(r["loc_off"]), int(r["n_locations"])
            assert ends[lo:lo + nl].tolist() == v["end_locations"], ctx
            if v["start_locations"] is None:
                assert int(r["has_start_locations"]) == 0, ctx
            else:
                assert starts[lo:lo + nl].tolist() == v["start_locations"], ctx
            if task == 2 and v["alignment"] is not None and v["edit_distance"] >= 0:
                ao, al = int(r["aln_off"]), int(r["alignment_length"])
                assert aln[ao:ao + al].tolist() == list(v["alignment"]), ctx


def test_long_queries_and_hirschberg_golden_vectors(gpu_ctx):
    """queries of more than 2048 rows (strips of 32 blocks) and paths at or above edlib's 1 MiB switch
    (obtainAlignmentHirschberg, src/edlib/src/edlib.cpp:1236-1401) against vectors of the vendored edlib"""
    from test_oracle_edlib import golden_vectors_long
    _check_vectors(gpu_ctx, golden_vectors_long())


def test_long_and_short_pairs_in_one_batch(gpu_ctx, oracle):
    """one batch holding ordinary pairs, pairs in the Hirschberg regime and queries of several strips: the LONG launch
    redoes only its own pairs"""
    rng = np.random.default_rng(21)
    qs, ts = [], []
    for i in range(24):
        if i % 6 == 0:
            t = _rnd(rng, 3000)
            q = _mutate(rng, t[200:2800], 0.08)            # long query, Hirschberg
        elif i % 6 == 1:
            t = _rnd(rng, 2600)
            q = _mutate(rng, t[:2000], 0.05)[:2048]        # one strip, Hirschberg
        elif i % 6 == 2:
            t = _rnd(rng, 200)
            q = _rnd(rng, 2300)                            # long query, traceback
        else:
            t = _rnd(rng, 900)
            q = _mutate(rng, t[300:600], 0.1)
        qs.append(q)
        ts.append(t)
    for mode in (0, 2):
        _check_batch(gpu_ctx, oracle, qs, ts, -1, mode, 2, use_ref=True)
    _check_batch(gpu_ctx, oracle, qs, ts, -1, 1, 1, use_ref=True)


def test_query_longer_than_limit_fails_loudly(gpu_ctx):
    from hairsplitter_b200 import api
    with pytest.raises(api.HsgpuError):
        gpu_ctx.edlib_align_batch([b"A" * ((1 << 20) + 1)], [b"ACGT" * 10], k=-1, mode=0, task=0)


def test_single_pair_shim_has_edlibs_result_semantics(gpu_ctx, oracle):
    """hsgpu_edlibAlign: edlib's calling convention for one pair (the one-line change at the reference's call sites):
    same fields, NULL pointers where edlib leaves them NULL, arrays released by hsgpu_edlibFreeAlignResult"""
    from hairsplitter_b200 import api
    from oracle import pyoracle
    rng = np.random.default_rng(77)
    ref = pyoracle.RefEdlib if pyoracle.RefEdlib.available() else None
    cases_ = []
    for it in range(24):
        t = _rnd(rng, int(rng.choice([0, 1, 40, 300, 2300])))
        qlen = int(rng.choice([0, 1, 20, 64, 150, 300]))
        if len(t) > qlen > 0 and rng.random() < 0.7:
            s0 = int(rng.integers(0, len(t) - qlen + 1))
            q = _mutate(rng, t[s0:s0 + qlen], 0.1)
        else:
            q = _rnd(rng, qlen)
        cases_.append((q, t))
    for q, t in cases_:
        for k, mode, task in ((-1, 2, 2), (3, 2, 2), (-1, 0, 1), (-1, 1, 0)):
            got = gpu_ctx.edlib_align(q, t, k=k, mode=mode, task=task)
            want = ref.align(q, t, k, mode, task) if ref is not None else oracle.edlib_align(q, t, k, mode, task)
            ctx = (len(q), len(t), k, mode, task)
            assert got["status"] == 0 and got["edit_distance"] == want["edit_distance"], ctx
            assert got["alphabet_length"] == want["alphabet_length"], ctx
            for key in ("end_locations", "start_locations"):
                if want[key] is None or len(want[key]) == 0:
                    assert got[key] is None, (key, ctx)
                else:
                    assert np.array_equal(got[key], want[key]), (key, ctx)
            if task == 2 and want["edit_distance"] >= 0 and want.get("alignment") is not None and len(want["alignment"]):
                assert np.array_equal(got["alignment"], want["alignment"]), ctx
            elif task != 2 or want["edit_distance"] < 0:
                assert got["alignment"] is None, ctx
    # additional equalities are refused, not ignored
    cfg = api.EdlibAlignConfig(-1, 2, 2, None, 1)
    r = gpu_ctx.lib.hsgpu_edlibAlign(gpu_ctx.h, b"ACGT", 4, b"ACGT", 4, cfg)
    assert r.status == 1 and not r.endLocations


def test_golden_vectors_of_the_vendored_edlib(gpu_ctx):
    """committed vectors made from the reference's vendored edlib (tests/golden/make_golden_edlib.py): the batch
    kernel reproduces distance, locations and path without oracle/_ref at run time"""
    import gzip
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    vec = json.loads(gzip.open(os.path.join(root, "tests", "golden", "edlib_vectors.json.gz")).read())
    _check_vectors(gpu_ctx, vec)
