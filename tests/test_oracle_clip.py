"""CPU: the oracle's restatement of modify_GFA's read-clipping walk (oracle/hs_oracle.c: hso_clip_read, reference
src/create_new_contigs.cpp:392-447) against the committed golden vectors (made by the reference's own loop body,
tests/golden/make_golden_clip.py) and, where oracle/_ref is present, against that loop body on fresh random cases;
and the host-side reconstruction of the clipped CIGAR string from the op range the C ABI returns."""
import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def _vectors():
    return json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "clip_vectors.json.gz")).read())


def op_range(ops, a, b):
    """(op_first, op_first_skip, op_last, op_last_take) of the expanded range [a, b), as hsgpu_clip reports it"""
    at, first, last = 0, None, None
    for k, o in enumerate(ops):
        n = o >> 4
        if first is None and a < at + n:
            first = (k, a - at)
        if last is None and b <= at + n and first is not None and n > 0 and b > at:
            last = (k, b - at)
        at += n
    if first is None:
        first = (len(ops), 0)
    if last is None:
        last = (len(ops), 0)
    return first + last


def test_oracle_matches_golden_vectors(oracle):
    vec = _vectors()
    assert len(vec) >= 1000
    n_ok = 0
    for r in vec:
        st, out = oracle.clip_read(r["ops"], r["pos"], r["left"], r["right"])
        assert st == r["status"], r
        if st == 0:
            assert [int(x) for x in out] == r["out"], r
            n_ok += 1
    assert 200 < n_ok < len(vec)


def test_clipped_cigar_string_from_op_range():
    from hairsplitter_b200 import api
    for r in _vectors():
        if r["status"] != 0:
            continue
        a, b = r["out"][2], r["out"][3]
        f, fs, l, lt = op_range(r["ops"], a, b)
        clip = {"op_first": f, "op_first_skip": fs, "op_last": l, "op_last_take": lt}
        got = api.clipped_cigar(r["ops"], clip) if b > a else ""
        want = r["clipped"] if b > a else ""
        assert got == want, (r, clip)


def test_oracle_matches_reference_loop_on_fresh_cases(oracle):
    from oracle.pyoracle import RefClip
    if not RefClip.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    import make_golden_clip
    rng = np.random.default_rng(5)
    for i in range(3000):
        ops, pos, left, right = make_golden_clip.random_case(rng, realistic=i % 2 == 0)
        cigar = "".join(f"{o >> 4}{'MIDNSHP=X'[o & 15]}" for o in ops)
        st, out = oracle.clip_read(ops, pos, left, right)
        rs, rout = RefClip.clip_read(cigar, pos, left, right)
        assert st == rs, (cigar, pos, left, right)
        if st == 0:
            assert np.array_equal(out, rout), (cigar, pos, left, right)
