#!/usr/bin/env python
"""The two drop-in executables on a BASELINE config at FULL size, against the unmodified reference.

  python scripts/full_config.py --config K --mode golden
      build container (needs oracle/_ref): generates the config's input files, runs the reference's
      HS_call_variants and the RNG-pinned reference HS_separate_reads (oracle/Makefile) on all host cores and
      commits the SHA-1 of every contig's block of the .col and .gro files to tests/golden/full_config_K.json
      (the files themselves are hundreds of megabytes).
  python scripts/full_config.py --config K --mode check
      GPU box: generates the same files (the generator is seeded and the box runs the same image), runs
      hairsplitter_b200/bin/HS_call_variants and HS_separate_reads (HS_PIN_SEED = the reference's pinned
      std::random_device value) and compares every contig's block with the committed hashes.
      Writes gpurun_out/full_config_K.json (parity booleans, wall times) and exits non-zero on any difference.

Contig blocks are compared one by one because the reference writes the contigs in the order its OpenMP threads
finish them.
"""
import argparse
import hashlib
import json
import multiprocessing as mp
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from hairsplitter_b200 import synth  # noqa: E402


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _job(job):
    config, seed, ci, length, special = job
    if special == "amplicon":
        cb = synth.amplicon_contig(np.random.default_rng(seed + 7919))
    elif special == "config":  # configs generated through one generator (1 and 2): made in the parent
        raise RuntimeError("not a pool job")
    else:
        cb = synth._make_chunk_job((config, seed, ci, length))
    return synth.chunk_text(cb) + (cb.n_reads, cb.length)


def write_inputs(config, scale, prefix, workers):
    """streams the chunks of a config into <prefix>.gfa/.fasta/.sam without keeping them in memory"""
    t0 = time.perf_counter()
    n_reads = n_cols = 0
    names = []
    if config in (1, 2):
        chunks, info = synth.make_config(config, scale=scale, seed=config)
        synth.write_files(chunks, prefix, links=info["links"])
        n_reads, n_cols = sum(c.n_reads for c in chunks), sum(c.length for c in chunks)
        names = [c.name for c in chunks]
        del chunks
    else:
        spec = synth.CONFIG_SPEC[config]
        total = int(spec["genome"] * scale)
        lengths = [min(300_000, total - o) for o in range(0, total, 300_000)]
        jobs = [(config, config, ci, l, None) for ci, l in enumerate(lengths)]
        if config == 5:
            jobs.append((config, config, len(lengths), 10_000, "amplicon"))
        _, info = synth.make_config(config, scale=scale, seed=config, n_chunks=0)
        with open(prefix + ".gfa", "wb") as gfa, open(prefix + ".fasta", "wb") as fa, open(prefix + ".sam", "wb") as sam, \
                open(prefix + ".sam.body", "wb") as body:
            with mp.get_context("fork").Pool(min(workers, len(jobs))) as pool:
                for t in pool.imap(_job, jobs, chunksize=1):
                    gfa.write(t[0])
                    sam.write(t[1])   # @SQ header lines first ...
                    fa.write(t[2])
                    body.write(t[3])  # ... the records after all of them
                    n_reads += t[4]
                    n_cols += t[5]
                    names.append(t[0].split(b"\t")[1].decode())
            body.flush()
            with open(prefix + ".sam.body", "rb") as b:
                shutil.copyfileobj(b, sam, 1 << 24)
        os.remove(prefix + ".sam.body")
        info["chunks"] = len(jobs)
    info.update(reads=n_reads, columns=n_cols, contigs=len(names), generation_s=round(time.perf_counter() - t0, 1),
                input_mb=round(sum(os.path.getsize(prefix + e) for e in (".gfa", ".fasta", ".sam")) / 1e6))
    return info


def block_hashes(path):
    """contig name -> SHA-1 of its block (CONTIG line up to the next CONTIG line, blank lines dropped)"""
    out, cur, h = {}, None, None
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b"CONTIG\t"):
                if cur is not None:
                    out[cur] = h.hexdigest()
                cur = line.split(b"\t")[1].decode()
                h = hashlib.sha1()
            if cur is not None and line.strip():
                h.update(line)
    if cur is not None:
        out[cur] = h.hexdigest()
    return out


def count_lines(path, prefix):
    n = 0
    with open(path, "rb") as f:
        for line in f:
            n += line.startswith(prefix)
    return n


def run(cmd, env=None):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise SystemExit(f"{cmd[0]} failed ({r.returncode}): {r.stderr.decode()[-2000:]}")
    return dt, r.stderr.decode()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True)
    ap.add_argument("--mode", choices=["golden", "check"], required=True)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--tmp", default=None)
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--gpus", type=int, default=1)
    args = ap.parse_args()
    from oracle.pyoracle import PIN_SEED
    cores = host_cores()
    tag = f"full_config_{args.config}" + ("" if args.scale == 1.0 else f"_x{args.scale:g}")
    golden_path = os.path.join(ROOT, "tests", "golden", tag + ".json")
    tmp = tempfile.mkdtemp(prefix="hs_full_", dir=args.tmp)
    try:
        prefix = os.path.join(tmp, "in")
        info = write_inputs(args.config, args.scale, prefix, cores)
        gfa, reads, sam = prefix + ".gfa", prefix + ".fasta", prefix + ".sam"
        amp = str(info["amplicon"])
        col, vcf, err, gro = [os.path.join(tmp, "out." + e) for e in ("col", "vcf", "err", "gro")]
        if args.mode == "golden":
            cv = os.path.join(ROOT, "oracle", "_ref", "HS_call_variants")
            sr = os.path.join(ROOT, "oracle", "_ref", "HS_separate_reads_pinned")
            threads = min(cores, info["contigs"])
            env = dict(os.environ)
        else:
            cv = os.path.join(ROOT, "hairsplitter_b200", "bin", "HS_call_variants")
            sr = os.path.join(ROOT, "hairsplitter_b200", "bin", "HS_separate_reads")
            threads = cores
            env = dict(os.environ, HS_PIN_SEED=str(PIN_SEED), HS_TIMING="1", HSGPU_NGPUS=str(args.gpus))
        t_cv, log_cv = run([cv, gfa, reads, sam, str(threads), tmp, err, amp, "0", col, vcf, "0.33"], env)
        error_rate = open(err).read().split()[0]
        our_error_rate = error_rate
        if args.mode == "check":
            # the reference sums the per-contig values in the order its threads finish (float), so its last printed
            # digit can differ from ours; both HS_separate_reads runs get the same <error_rate> argument
            error_rate = json.load(open(golden_path))["error_rate"]
        t_sr, log_sr = run([sr, col, str(threads), error_rate, os.path.join(tmp, "no_ploidy"), "0", "0", amp, gro, "0"], env)
        res = {"config": args.config, "scale": args.scale, "workload": info["description"], "contigs": info["contigs"],
               "reads": info["reads"], "columns": info["columns"], "input_mb": info["input_mb"], "amplicon": info["amplicon"],
               "snps": count_lines(col, b"SNPS\t"), "groups": count_lines(gro, b"GROUP\t"), "error_rate": our_error_rate,
               "threads": threads, "call_variants_s": round(t_cv, 2), "separate_reads_s": round(t_sr, 2),
               "generation_s": info["generation_s"]}
        hc, hg = block_hashes(col), block_hashes(gro)
        if args.mode == "golden":
            res.update(host=f"build container, {cores} host cores", col_sha1=hc, gro_sha1=hg,
                       how="python scripts/full_config.py --config %d --mode golden" % args.config +
                           ("" if args.scale == 1.0 else f" --scale {args.scale:g}"))
            with open(golden_path, "w") as f:
                json.dump(res, f, indent=1, sort_keys=True)
            print(json.dumps({k: v for k, v in res.items() if not k.endswith("_sha1")}))
        else:
            gold = json.load(open(golden_path))
            bad_col = sorted(k for k in set(hc) | set(gold["col_sha1"]) if hc.get(k) != gold["col_sha1"].get(k))
            bad_gro = sorted(k for k in set(hg) | set(gold["gro_sha1"]) if hg.get(k) != gold["gro_sha1"].get(k))
            res.update(gpus=args.gpus, col_identical_to_reference=not bad_col, gro_identical_to_pinned_reference=not bad_gro,
                       col_blocks_differing=bad_col[:10], gro_blocks_differing=bad_gro[:10],
                       error_rate_reference=gold["error_rate"],
                       reference_call_variants_s=gold["call_variants_s"], reference_separate_reads_s=gold["separate_reads_s"],
                       reference_host=gold["host"], reference_threads=gold["threads"],
                       timing_call_variants=[l.replace("[hs timing]", "").strip() for l in log_cv.splitlines() if l.startswith("[hs timing]")],
                       timing_separate_reads=[l.replace("[hs timing]", "").strip() for l in log_sr.splitlines() if l.startswith("[hs timing]")])
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", tag + ".json"), "w") as f:
                json.dump(res, f, indent=1, sort_keys=True)
            print(json.dumps({k: v for k, v in res.items() if not k.startswith("timing_")}))
            if bad_col or bad_gro:
                raise SystemExit(f"{tag}: {len(bad_col)} .col blocks and {len(bad_gro)} .gro blocks differ from the reference")
    finally:
        if not args.keep:
            shutil.rmtree(tmp, ignore_errors=True)
        else:
            print("kept", tmp)


if __name__ == "__main__":
    main()
