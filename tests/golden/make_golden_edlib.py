"""Golden vectors of edlibAlign from the VENDORED edlib of the reference (oracle/_ref/libhsref_edlib.so, built from
/root/reference/src/edlib by oracle/Makefile). Run:  python tests/golden/make_golden_edlib.py
Writes edlib_vectors.json.gz: a list of {q, t, k, mode, task, status, edit_distance, alphabet_length, end_locations,
start_locations (null where edlib leaves it NULL), alignment (null unless task == PATH)}."""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

ALPHA = b"ACGT"


def rnd(rng, n, alpha=ALPHA):
    return bytes(rng.choice(list(alpha), n).tolist())


def mutate(rng, s, e):
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


def pairs(rng):
    out = []
    for qlen in (0, 1, 5, 63, 64, 65, 128, 129, 200, 300, 640):
        for tlen in (0, 1, 10, 64, 150, 400, 700):
            t = rnd(rng, tlen)
            if tlen > qlen > 0 and rng.random() < 0.7:
                s = int(rng.integers(0, tlen - qlen + 1))
                q = mutate(rng, t[s:s + qlen], float(rng.choice([0, 0.05, 0.2])))
            else:
                q = rnd(rng, qlen, ALPHA if rng.random() < 0.8 else b"ACGTNRY-*")
            out.append((q, t))
    # the in-pipeline shape: <= 300 bp query, ~2.3 kb target (src/create_new_contigs.cpp:557-630)
    for _ in range(6):
        t = rnd(rng, 2300)
        s = int(rng.integers(0, 2000))
        out.append((mutate(rng, t[s:s + 280], 0.1), t))
    # repeats: many equally good end positions
    out.append((b"ACAC" * 8, b"AC" * 200))
    out.append((b"A" * 70, b"A" * 300))
    return out


def main():
    from oracle.pyoracle import RefEdlib
    rng = np.random.default_rng(4242)
    vec = []
    for q, t in pairs(rng):
        for mode in (0, 1, 2):
            for task, k in ((2, -1), (1, -1), (0, 5), (2, 40)):
                r = RefEdlib.align(q, t, k, mode, task)
                vec.append(dict(q=q.decode("latin1"), t=t.decode("latin1"), k=k, mode=mode, task=task, status=int(r["status"]),
                                edit_distance=int(r["edit_distance"]), alphabet_length=int(r["alphabet_length"]),
                                end_locations=[int(x) for x in r["end_locations"]],
                                start_locations=None if r["start_locations"] is None else [int(x) for x in r["start_locations"]],
                                alignment=None if r["alignment"] is None else [int(x) for x in r["alignment"]]))
    with gzip.GzipFile(os.path.join(HERE, "edlib_vectors.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(vec, separators=(",", ":")).encode())
    print(len(vec), "vectors,", os.path.getsize(os.path.join(HERE, "edlib_vectors.json.gz")), "bytes")


if __name__ == "__main__":
    main()
