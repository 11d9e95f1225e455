#!/usr/bin/env python
"""How long HS_separate_reads' first phase takes on the .col text and on its binary sidecar.

CPU only (no GPU code runs): the inputs of a BASELINE config are generated at full size, the reference's own
HS_call_variants (oracle/_ref) writes the .col, the drop-in's writer (write_outputs through libhshost.so) prints it
again -- checked block for block against the reference's text -- and leaves the sidecar; then the drop-in's
parse_column_file is timed on both routes, and the structures it leaves are compared (one digest over every field).

  python scripts/sidecar_parse_time.py --config 2 [--scale 1.0] [--repeat 3]  ->  profiles/sidecar_parse_config2.json
"""
import argparse
import ctypes as C
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import full_config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    cores = full_config.host_cores()
    L = C.CDLL(os.path.join(ROOT, "hairsplitter_b200", "libhshost.so"))
    L.hshost_col_digest.restype = C.c_double
    L.hshost_col_digest.argtypes = [C.c_char_p, C.c_int, C.c_float, C.c_int, C.POINTER(C.c_uint64)]
    L.hshost_rewrite_col.restype = C.c_int
    L.hshost_rewrite_col.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
    tmp = tempfile.mkdtemp(prefix="hs_sidecar_")
    try:
        prefix = os.path.join(tmp, "in")
        info = full_config.write_inputs(args.config, args.scale, prefix, cores)
        ref_col, col, vcf, err = [os.path.join(tmp, n) for n in ("ref.col", "out.col", "out.vcf", "out.err")]
        cv = os.path.join(ROOT, "oracle", "_ref", "HS_call_variants")
        threads = min(cores, info["contigs"])
        t_ref, _ = full_config.run([cv, prefix + ".gfa", prefix + ".fasta", prefix + ".sam", str(threads), tmp, err,
                                    str(info["amplicon"]), "0", ref_col, os.path.join(tmp, "ref.vcf"), "0.33"])
        # what the sidecar adds to the writer: the same write_outputs call with and without it
        t_write = {}
        for flag in ("0", "1", "0", "1"):
            os.environ["HS_SIDECAR"] = flag
            t0 = time.perf_counter()
            assert L.hshost_rewrite_col(ref_col.encode(), col.encode(), vcf.encode()) == 0
            t_write.setdefault(flag, []).append(round(time.perf_counter() - t0, 3))
        del os.environ["HS_SIDECAR"]
        same_text = full_config.block_hashes(ref_col) == full_config.block_hashes(col)

        def parse(route):
            out = (C.c_uint64 * 5)()
            dt = L.hshost_col_digest(col.encode(), 1000000000, 0.0, route, out)
            assert dt >= 0, "route %d refused" % route
            return dt, tuple(out)

        text, side = [], []
        for _ in range(args.repeat):
            dt, d_text = parse(0)
            text.append(dt)
            dt, d_side = parse(1)
            side.append(dt)
        assert d_side[4] == 1 and d_text[4] == 0
        res = {"config": args.config, "scale": args.scale, "workload": info["description"], "host_cores": cores,
               "contigs": int(d_text[1]), "snps": int(d_text[2]), "cells": int(d_text[3]),
               "col_text_mb": round(os.path.getsize(col) / 1e6, 1), "sidecar_mb": round(os.path.getsize(col + ".hsb") / 1e6, 1),
               "reference_call_variants_s": round(t_ref, 1),
               "parse_text_plus_write_outputs_s": {"without_sidecar": t_write["0"], "with_sidecar": t_write["1"]},
               "text_identical_to_reference": same_text, "structures_identical": d_text[:4] == d_side[:4],
               "parse_text_s": [round(x, 4) for x in text], "parse_sidecar_s": [round(x, 4) for x in side],
               "parse_text_median_s": round(sorted(text)[len(text) // 2], 4),
               "parse_sidecar_median_s": round(sorted(side)[len(side) // 2], 4),
               "how": "python scripts/sidecar_parse_time.py --config %d --scale %g (CPU only; page cache warm on every run)" % (args.config, args.scale)}
        out = args.out or os.path.join(ROOT, "profiles", "sidecar_parse_config%d%s.json" % (args.config, "" if args.scale == 1.0 else "_x%g" % args.scale))
        with open(out, "w") as f:
            json.dump(res, f, indent=1)
        print(json.dumps(res))
        if not (same_text and res["structures_identical"]):
            raise SystemExit("sidecar and text differ")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
