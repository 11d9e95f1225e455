"""Golden vectors of edlibAlign from the VENDORED edlib of the reference (oracle/_ref/libhsref_edlib.so, built from
/root/reference/src/edlib by oracle/Makefile). Run:  python tests/golden/make_golden_edlib.py
Writes edlib_vectors.json.gz: a list of {q, t, k, mode, task, status, edit_distance, alphabet_length, end_locations,
start_locations (null where edlib leaves it NULL), alignment (null unless task == PATH)}, and
edlib_vectors_long.json.gz: the same for queries longer than 2048 and for paths at or above edlib's 1 MiB switch to
Hirschberg (src/edlib/src/edlib.cpp:1193-1195), alignment as a string of digits."""
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


def pairs_long(rng):
    """(q, t, mode, task, k): long queries and/or the Hirschberg regime"""
    out = []
    t = rnd(rng, 2600)
    out.append((mutate(rng, t[:2040], 0.02)[:2048], t, 0, 2, -1))           # 32 blocks, just above the switch
    t = rnd(rng, 3000)
    out.append((mutate(rng, t, 0.1), t, 0, 2, -1))
    out.append((mutate(rng, t, 0.1), t, 0, 2, 100))                         # k too small: -1
    t = rnd(rng, 6000)
    out.append((mutate(rng, t[1700:4200], 0.15), t, 2, 2, -1))              # HW, query of 40 blocks
    out.append((mutate(rng, t[1700:4200], 0.05), t, 2, 1, -1))
    out.append((mutate(rng, t[:2300], 0.1), t, 1, 2, -1))                   # SHW
    t = rnd(rng, 1200)
    out.append((mutate(rng, t + rnd(rng, 3800), 0.1), t, 0, 2, -1))         # 5000 x 1200
    t = rnd(rng, 9000)
    out.append((mutate(rng, t[4000:4700], 0.2), t, 0, 2, -1))               # 700 x 9000: 11 blocks, switch by columns
    out.append((rnd(rng, 2500), rnd(rng, 300), 0, 2, -1))                   # long query below the switch, unrelated
    t = rnd(rng, 350)
    out.append((mutate(rng, t, 0.1) + rnd(rng, 2200), t, 0, 2, -1))
    out.append((b"AC" * 1100, b"AC" * 1500, 2, 2, -1))                      # many end and start locations
    out.append((b"A" * 3000, b"A" * 3500, 0, 2, -1))                        # every path is optimal: tie rules only
    out.append((rnd(rng, 2100, b"AC"), rnd(rng, 2300, b"AC"), 0, 2, -1))    # two letters, unrelated: long ties
    t = rnd(rng, 2300, b"ACDEFGHIKLMNPQRSTVWY")
    out.append((mutate(rng, t, 0.3)[:2100], t, 0, 2, -1))                   # 20 letters: Peq beyond shared memory
    t = rnd(rng, 10000)
    out.append((mutate(rng, t, 0.12), t, 0, 2, -1))                         # 10 kb x 10 kb: several levels, 5 strips
    out.append((mutate(rng, t[2000:7000], 0.1), t, 2, 2, -1))
    out.append((mutate(rng, t[2000:7000], 0.1), t, 2, 0, 800))
    return out


def main():
    from oracle.pyoracle import RefEdlib
    rng = np.random.default_rng(777)
    vec = []
    for q, t, mode, task, k in pairs_long(rng):
        r = RefEdlib.align(q, t, k, mode, task)
        vec.append(dict(q=q.decode("latin1"), t=t.decode("latin1"), k=k, mode=mode, task=task, status=int(r["status"]),
                        edit_distance=int(r["edit_distance"]), alphabet_length=int(r["alphabet_length"]),
                        end_locations=[int(x) for x in r["end_locations"]],
                        start_locations=None if r["start_locations"] is None else [int(x) for x in r["start_locations"]],
                        alignment=None if r["alignment"] is None else "".join(str(int(x)) for x in r["alignment"])))
    with gzip.GzipFile(os.path.join(HERE, "edlib_vectors_long.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(vec, separators=(",", ":")).encode())
    print(len(vec), "long vectors,", os.path.getsize(os.path.join(HERE, "edlib_vectors_long.json.gz")), "bytes")
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
