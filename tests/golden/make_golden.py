"""Generates the committed golden fixtures from the UNMODIFIED reference (needs oracle/_ref, i.e. the
build container with /root/reference). Run:  python tests/golden/make_golden.py

Each cv_*.npz holds a small seeded input (contig, reads, CIGARs) and what the reference's own
generate_msa / call_variants / keep_only_robust_variants produced for it (src/call_variants.cpp).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))


def batch_from_npz(g):
    from hairsplitter_b200 import synth
    return synth.ContigBatch(contig=g["contig"], read_bases=g["read_bases"], read_off=g["read_off"], cigar=g["cigar"],
                             cigar_off=g["cigar_off"], start=g["start"], strand=g["strand"],
                             strain=np.zeros(g["start"].shape[0], np.int32), name="golden")


def parts_from_npz(g):
    off = g["part_off"]
    return [dict(read_idx=g["part_idx"][off[i]:off[i + 1]], state=g["part_state"][off[i]:off[i + 1]],
                 more=g["part_more"][off[i]:off[i + 1]], less=g["part_less"][off[i]:off[i + 1]])
            for i in range(off.shape[0] - 1)]


def main():
    import cases
    from oracle.pyoracle import RefCV
    todo = {
        "cv_ont_small": cases.small_case(seed=1001, length=4000, depth=25, mean_len=1200),
        "cv_hifi_small": cases.small_case(seed=1002, length=4000, depth=30, mean_len=1500, error=0.005, hard=0.0),
        "cv_eqx_small": cases.small_case(seed=1003, length=3000, depth=20, mean_len=900, eqx=True),
    }
    for name, cb in todo.items():
        R = RefCV(cb)
        p = R.pileup()
        rc = R.call_variants()
        parts, filt, merged = R.robust()
        off = np.zeros(len(parts) + 1, np.int64)
        if parts:
            np.cumsum([len(q["read_idx"]) for q in parts], out=off[1:])
        cat = lambda k, dt: np.concatenate([q[k] for q in parts]).astype(dt) if parts else np.zeros(0, dt)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            contig=cb.contig, read_bases=cb.read_bases, read_off=cb.read_off, cigar=cb.cigar, cigar_off=cb.cigar_off,
            start=cb.start, strand=cb.strand,
            col_off=p["col_off"], read_idx=p["read_idx"], code=p["code"], mean_distance=np.float32(R.mean_distance()),
            newref=R.newref(), read_end=R.read_limits()[1], ref_base=rc["ref_base"], second_base=rc["second_base"],
            suspect_pos=rc["suspects"]["pos"], automatic_pos=rc["automatic"]["pos"], depth=np.float32(rc["depth"]),
            part_off=off, part_idx=cat("read_idx", np.int32), part_state=cat("state", np.int16),
            part_more=cat("more", np.int32), part_less=cat("less", np.int32),
            filtered_pos=filt["pos"], merged_pos=merged["pos"],
        )
        print(name, cb.n_reads, "reads", p["code"].shape[0], "cells", len(parts), "partitions", filt["pos"].shape[0],
              "kept")


if __name__ == "__main__":
    main()
