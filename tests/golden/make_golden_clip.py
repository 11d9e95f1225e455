"""Golden vectors of the read-clipping walk of modify_GFA (reference src/create_new_contigs.cpp:392-447), produced by
the reference's own loop body (oracle/_ref/libhsref_clip.so, built by oracle/Makefile from the reference source).
Run in the build container:  python tests/golden/make_golden_clip.py  -> tests/golden/clip_vectors.json.gz"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
LETTERS = "MIDNSHP=X"


def random_case(rng, realistic):
    ops = []
    if realistic:  # an aligned read: clips at the ends, M runs broken by short indels
        if rng.random() < 0.5:
            ops.append((int(rng.integers(1, 300)) << 4) | int(rng.choice([4, 5])))
        for _ in range(int(rng.integers(3, 60))):
            ops.append((int(rng.integers(1, 40)) << 4) | 0)
            ops.append((int(rng.integers(1, 4)) << 4) | int(rng.choice([1, 2])))
        ops.append((int(rng.integers(1, 40)) << 4) | 0)
        if rng.random() < 0.5:
            ops.append((int(rng.integers(1, 300)) << 4) | int(rng.choice([4, 5])))
        span = sum(o >> 4 for o in ops if (o & 15) in (0, 2))
        pos = int(rng.integers(0, 5000))
        left = pos + int(rng.integers(-50, span + 50))
        right = left + int(rng.integers(0, 400))
    else:  # anything goes: every letter, empty ops, clips in the middle, degenerate intervals
        for _ in range(int(rng.integers(1, 14))):
            ty = int(rng.choice([0, 0, 0, 1, 2, 4, 5, 7, 8, 3, 6]))
            ln = int(rng.integers(0, 9)) if rng.random() < 0.8 else int(rng.integers(0, 70))
            ops.append((ln << 4) | ty)
        pos = int(rng.integers(0, 30))
        left = int(rng.integers(0, 70))
        right = left + int(rng.integers(-3, 50))
    return ops, pos, left, right


def main():
    from oracle.pyoracle import RefClip
    rng = np.random.default_rng(77)
    out = []
    for i in range(1500):
        ops, pos, left, right = random_case(rng, realistic=i % 3 != 0)
        cigar = "".join(f"{o >> 4}{LETTERS[o & 15]}" for o in ops)
        st, v = RefClip.clip_read(cigar, pos, left, right)
        rec = {"ops": ops, "pos": pos, "left": left, "right": right, "status": st}
        if st == 0:
            rec["out"] = [int(x) for x in v]
            rec["clipped"] = RefClip.clip_cigar(cigar, v[2], v[3])
        out.append(rec)
    with gzip.GzipFile(os.path.join(HERE, "clip_vectors.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    print(len(out), "vectors,", sum(1 for r in out if r["status"] == 0), "with status 0")


if __name__ == "__main__":
    main()
