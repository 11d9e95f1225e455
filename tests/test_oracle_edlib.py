"""CPU tests of the edlib restatement (oracle/hs_oracle_edlib.c): against the committed golden vectors made from the
reference's vendored edlib (tests/golden/make_golden_edlib.py) and, where oracle/_ref travelled, against that edlib
itself on fresh random pairs."""
import gzip
import json
import os

import numpy as np
import pytest

from oracle.pyoracle import Oracle, RefEdlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def golden_vectors():
    return json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "edlib_vectors.json.gz")).read())


def golden_vectors_long():
    vec = json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "edlib_vectors_long.json.gz")).read())
    for v in vec:
        if v["alignment"] is not None:
            v["alignment"] = [int(c) for c in v["alignment"]]
    return vec


def check_against(want, got, ctx):
    """`got` in the oracle's dict form, `want` a golden vector or a RefEdlib.align result"""
    assert got["edit_distance"] == want["edit_distance"], ctx
    assert got["alphabet_length"] == want["alphabet_length"], ctx
    assert list(got["end_locations"]) == list(want["end_locations"]), ctx
    if want["start_locations"] is None:
        assert got["start_locations"] is None, ctx
    else:
        assert list(got["start_locations"]) == list(want["start_locations"]), ctx
    assert got["status"] == want["status"], ctx
    if want["alignment"] is not None and want["edit_distance"] >= 0:
        assert list(got["alignment"]) == list(want["alignment"]), ctx


def test_oracle_matches_golden_vectors(oracle):
    vec = golden_vectors()
    assert len(vec) >= 1000
    for i, v in enumerate(vec):
        q, t = v["q"].encode("latin1"), v["t"].encode("latin1")
        got = oracle.edlib_align(q, t, v["k"], v["mode"], v["task"])
        check_against(v, got, (i, len(q), len(t), v["k"], v["mode"], v["task"]))


def test_oracle_matches_golden_vectors_long_queries_and_hirschberg(oracle):
    """queries of more than 2048 rows and paths at or above edlib's 1 MiB switch (obtainAlignmentHirschberg)"""
    vec = golden_vectors_long()
    assert len(vec) >= 17
    n_hirschberg = 0
    for i, v in enumerate(vec):
        q, t = v["q"].encode("latin1"), v["t"].encode("latin1")
        got = oracle.edlib_align(q, t, v["k"], v["mode"], v["task"])
        check_against(v, got, (i, len(q), len(t), v["k"], v["mode"], v["task"]))
        if v["task"] == 2 and v["edit_distance"] >= 0:
            an = v["end_locations"][0] - v["start_locations"][0] + 1
            n_hirschberg += (20 * ((len(q) + 63) // 64) + 8) * an >= 1 << 20
    assert n_hirschberg >= 10


@pytest.mark.skipif(not RefEdlib.available(), reason="oracle/_ref not built")
def test_oracle_matches_vendored_edlib_in_the_hirschberg_regime(oracle):
    rng = np.random.default_rng(5)
    alpha = list(b"ACGT")
    for it in range(10):
        n = int(rng.integers(1500, 4000))
        m = int(rng.integers(1200, 3500))
        t = bytes(rng.choice(alpha[: int(rng.integers(2, 5))], n).tolist())
        q = bytearray(t[:m] if rng.random() < 0.7 else bytes(rng.choice(alpha, m).tolist()))
        for j in range(len(q)):
            if rng.random() < 0.1:
                q[j] = int(rng.choice(alpha))
        mode = int(rng.choice([0, 2]))
        want = RefEdlib.align(bytes(q), t, -1, mode, 2)
        got = oracle.edlib_align(bytes(q), t, -1, mode, 2)
        check_against(want, got, (it, len(q), n, mode))


@pytest.mark.skipif(not RefEdlib.available(), reason="oracle/_ref not built")
def test_oracle_matches_vendored_edlib_on_random_pairs(oracle):
    rng = np.random.default_rng(99)
    alpha = list(b"ACGT")
    for it in range(300):
        tlen = int(rng.choice([0, 1, 7, 64, 65, 200, 500, 900]))
        qlen = int(rng.choice([0, 1, 3, 63, 64, 65, 127, 128, 300, 450]))
        t = bytes(rng.choice(alpha, tlen).tolist())
        if tlen > qlen > 0 and rng.random() < 0.6:
            s = int(rng.integers(0, tlen - qlen + 1))
            q = bytearray(t[s:s + qlen])
            for j in range(qlen):
                if rng.random() < 0.08:
                    q[j] = int(rng.choice(alpha))
            q = bytes(q)
        else:
            q = bytes(rng.choice(alpha, qlen).tolist())
        mode, task = int(rng.integers(0, 3)), int(rng.integers(0, 3))
        k = int(rng.choice([-1, 0, 3, 30, 1000]))
        want = RefEdlib.align(q, t, k, mode, task)
        got = oracle.edlib_align(q, t, k, mode, task)
        check_against(want, got, (it, qlen, tlen, k, mode, task))
