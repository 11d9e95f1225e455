#!/usr/bin/env python
"""The host side of bin/HS_call_variants on a BASELINE config (scaled), without a GPU: the product executable runs with
oracle/_ref/mock_for_glue_test/libhsgpu.so (test infrastructure: the C-ABI calls answered by the oracle) in front of the
real library for that one subprocess, the reference's HS_call_variants runs next to it, and the .col / .vcf files are
compared contig block by contig block. Then both .col files go through HS_separate_reads -- the pinned reference and
oracle/sr_hostcheck (the drop-in's host pipeline with the oracle as its stages, reading the sidecar) -- and the .gro
files are compared the same way. What this exercises at size is the drop-in's HOST code: parsers, packing, batching,
partition building, merge, writers, sidecar, window walk, clustering post-processing.

  python scripts/host_side_check.py --config 2 --scale 0.2   ->  profiles/host_side_check_config2_x0.2.json
"""
import argparse
import json
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import full_config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--scale", type=float, default=0.2)
    args = ap.parse_args()
    from oracle.pyoracle import PIN_SEED
    cores = full_config.host_cores()
    mock = os.path.join(ROOT, "oracle", "_ref", "mock_for_glue_test")
    tmp = tempfile.mkdtemp(prefix="hs_hostcheck_")
    try:
        prefix = os.path.join(tmp, "in")
        info = full_config.write_inputs(args.config, args.scale, prefix, cores)
        files = [prefix + e for e in (".gfa", ".fasta", ".sam")]
        amp = str(info["amplicon"])
        threads = str(min(cores, info["contigs"]))
        out = {}
        for tag, exe, env in (("ref", os.path.join(ROOT, "oracle", "_ref", "HS_call_variants"), None),
                              ("ours", os.path.join(ROOT, "hairsplitter_b200", "bin", "HS_call_variants"),
                               dict(os.environ, LD_LIBRARY_PATH=mock, HS_TIMING="1"))):
            col, vcf, err = [os.path.join(tmp, f"{tag}.{e}") for e in ("col", "vcf", "err")]
            dt, log = full_config.run([exe, *files, threads, tmp, err, amp, "0", col, vcf, "0.33"], env)
            out[tag] = dict(col=col, vcf=vcf, err=err, seconds=round(dt, 1), log=log)
        col_same = full_config.block_hashes(out["ref"]["col"]) == full_config.block_hashes(out["ours"]["col"])
        vcf_same = sorted(open(out["ref"]["vcf"], "rb").read().splitlines()) == sorted(open(out["ours"]["vcf"], "rb").read().splitlines())
        error_rate = open(out["ref"]["err"]).read().split()[0]
        gro = {}
        for tag, exe, env in (("ref", os.path.join(ROOT, "oracle", "_ref", "HS_separate_reads_pinned"), None),
                              ("ours", os.path.join(ROOT, "oracle", "sr_hostcheck"),
                               dict(os.environ, HS_PIN_SEED=str(PIN_SEED), HS_TIMING="1"))):
            path = os.path.join(tmp, tag + ".gro")
            dt, log = full_config.run([exe, out[tag]["col"], threads, error_rate, "none", "0", "0", amp, path, "0"], env)
            gro[tag] = dict(path=path, seconds=round(dt, 1), log=log)
        gro_same = full_config.block_hashes(gro["ref"]["path"]) == full_config.block_hashes(gro["ours"]["path"])
        res = {"config": args.config, "scale": args.scale, "workload": info["description"], "contigs": info["contigs"],
               "reads": info["reads"], "columns": info["columns"], "host_cores": cores,
               "snps": full_config.count_lines(out["ref"]["col"], b"SNPS\t"), "groups": full_config.count_lines(gro["ref"]["path"], b"GROUP\t"),
               "col_blocks_identical_to_reference": col_same, "vcf_lines_identical_to_reference": vcf_same,
               "error_rate": [open(out[t]["err"]).read().split()[0] for t in ("ref", "ours")],
               "gro_blocks_identical_to_pinned_reference": gro_same,
               "sidecar_used_by_the_host_pipeline": "binary sidecar" in gro["ours"]["log"],
               "host_phases_call_variants": [l.replace("[hs timing]", "").strip() for l in out["ours"]["log"].splitlines()
                                             if l.startswith("[hs timing]") and not any(k in l for k in ("gpu ", "robust filter", "upload", "context"))],
               "note": "no GPU: the C-ABI calls of the drop-in executable are answered by the oracle (oracle/mock_hsgpu.c) in this "
                       "subprocess only; phase times are those of a virtual machine with 8 cores",
               "how": "python scripts/host_side_check.py --config %d --scale %g" % (args.config, args.scale)}
        path = os.path.join(ROOT, "profiles", "host_side_check_config%d_x%g.json" % (args.config, args.scale))
        with open(path, "w") as f:
            json.dump(res, f, indent=1)
        print(json.dumps(res))
        if not (col_same and vcf_same and gro_same):
            raise SystemExit("host side differs from the reference")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
