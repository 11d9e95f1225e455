#!/usr/bin/env python
"""Benchmark of the hot path on B200 (contract: see the task statement / DESIGN.md section "Measurement").

A step = one pass of the per-chunk call_variants path -- generate_msa, call_variants (allele counting / ranking)
and loops 3+4 of keep_only_robust_variants, hot loops A+B+C of SURVEY.md section 8 -- over one batch of synthetic
input: BASELINE.json configs[1] unless --config says otherwise (a 5 Mb bacterial genome cut into 17 contig chunks
<= 300 kb, 2 strains 1 % apart, ONT-like reads 10 kb mean, 10 % error, 60x).
Metric: pileup windows/s, window = 2000 columns x depth, value = columns / 2000 / s.

  value   inputs resident in HBM (reads, CIGARs, final partitions), CUDA-event timed on the library's stream:
          hsgpu_pileup_build + hsgpu_column_rank + hsgpu_robust_filter_all per step
  e2e     the same through the C ABI from pinned HOST buffers: hsgpu_pileup_create (H2D) + build + rank +
          hsgpu_partitions_set (H2D) + hsgpu_robust_filter_all + D2H of the per-contig results (suspect lists,
          snps_out, counts), every step. The final partitions (loops 1-2, sequential host C++ in
          hairsplitter_b200/host) are an input of the C ABI and are built once outside the timed region.
  roofline / kernels   per-kernel event timing (hsgpu_profile_enable) over a second pass of the same steps
  cpu_baseline         the reference's own generate_msa + call_variants + keep_only_robust_variants (oracle/_ref,
                       all four loops) or, when the compiled reference did not travel, the C oracle
                       port, on the host cores, on a bounded sample of the same workload

`--impl reference` times only the CPU reference arm and prints the same JSON shape.
Multi-GPU (torchrun): contig chunks shard with no data-path collective; every rank processes its own
genome of the same shape (weak scaling); value = columns of all ranks / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WINDOW = 2000  # sizeOfWindow, reference src/separate_reads.cpp:1485


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--scale", type=float, default=1.0, help="shrinks the genome (testing only)")
    ap.add_argument("--cpu-sample-chunks", type=int, default=0, help="chunks in the CPU baseline sample (0 = auto)")
    ap.add_argument("--e2e-lanes", type=int, default=0, help="hsgpu contexts (host threads) of the e2e path; 0 = by the "
                    "host cores per rank: 5 on a box of its own (measured: 3 / 4 / 5 / 6 lanes = 3.2 / 2.8 / 2.6 / 3.1 ms per step), fewer "
                    "when several ranks share the host")
    ap.add_argument("--e2e-groups", type=int, default=1, help="groups of contig chunks the e2e path cuts a step's batch into")
    ap.add_argument("--strong", action="store_true", help="strong scaling: ONE workload (same seed on every rank) dealt "
                    "to the ranks by sharding.lpt_assign; value = all columns / max-over-ranks time")
    ap.add_argument("--only-realign", action="store_true", help="development: run the realign stage alone and print it")
    ap.add_argument("--no-stages", action="store_true", help="skip the side stages (realign is always measured)")
    ap.add_argument("--wall-ref-runs", type=int, default=3, help="runs of the reference HS_call_variants (median reported)")
    ap.add_argument("--wall-chunks", type=int, default=0, help="chunks in the HS_call_variants wall-time stage (0 = all, -1 = skip)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for i, n in enumerate(names):
                    if r[5 + i].lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_rate(chunks, n_threads, repeat=1):
    """generate_msa + call_variants + keep_only_robust_variants of the reference (oracle/_ref) -- or the C oracle
    port of the first two when the compiled reference did not travel -- over `chunks`, one chunk per thread at a
    time. Returns (windows/s, kind, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle
    use_ref = pyoracle.ref_available()
    if use_ref:
        import ctypes as C
        L = pyoracle.RefCV.lib()
        prepared = []
        for cb in chunks:  # string marshalling is not part of the measured work
            n = cb.n_reads
            reads = (C.c_char_p * n)(*[cb.read_str(i).encode() for i in range(n)])
            cigars = (C.c_char_p * n)(*[cb.cigar_str(i).encode() for i in range(n)])
            st = np.ascontiguousarray(cb.start, np.int32)
            sd = np.ascontiguousarray(cb.strand, np.uint8)
            prepared.append((cb.contig_str().encode(), n, reads, cigars, st, sd))

        def work(a):
            h = L.hsref_cv_create(a[0], a[1], a[2], a[3], a[4].ctypes.data, a[5].ctypes.data)  # generate_msa
            L.hsref_cv_call_variants(h, -1.0, 0.33)                                             # call_variants
            L.hsref_cv_robust(h, -1.0)                                                          # keep_only_robust_variants
            L.hsref_cv_destroy(h)
    else:
        O = pyoracle.Oracle()
        prepared = chunks

        def work(cb):
            p = O.pileup(cb)
            O.call_variants(p["col_off"], p["code"], O.mean_distance(*p["stats"]))
    cols = sum(cb.length for cb in chunks) * repeat
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=n_threads) as ex:
        for _ in range(repeat):
            list(ex.map(work, prepared))
    dt = time.perf_counter() - t0
    return cols / WINDOW / dt, ("reference" if use_ref else "port"), dt


def make_realign_pairs(rng, contig, n_pairs, qlen=1536, slack=0.15, err=0.10):
    """read-chunk x contig-window pairs of the realignment benchmark shape: a `qlen`-base
    chunk of an ONT-like read (10 % error, 1/3 each mismatch/insertion/deletion) against the contig window
    it came from plus 15 % slack. Returns (list of query bytes, list of target bytes)."""
    L = contig.shape[0]
    tlen = int(round(qlen * (1 + slack)))
    seg = qlen + 160
    margin = (tlen - qlen) // 2
    starts = rng.integers(0, L - seg - tlen, n_pairs)
    src = contig[starts[:, None] + margin + np.arange(seg)[None, :]]
    u = rng.random((n_pairs, seg))
    mism = u < err / 3
    dele = (u >= err / 3) & (u < 2 * err / 3)
    ins = (u >= 2 * err / 3) & (u < err)
    obs = np.where(mism, (src + rng.integers(1, 4, src.shape)) & 3, src).astype(np.uint8)
    vals = np.stack([obs, rng.integers(0, 4, src.shape, dtype=np.uint8)], axis=2).reshape(n_pairs, 2 * seg)
    valid = np.stack([~dele, ins], axis=2).reshape(n_pairs, 2 * seg)
    keep = valid & (np.cumsum(valid, axis=1) <= qlen)
    assert (keep.sum(axis=1) == qlen).all()
    ascii_ = np.frombuffer(b"ACGT", dtype=np.uint8)
    q = ascii_[vals[keep].reshape(n_pairs, qlen)]
    t = ascii_[contig[starts[:, None] + np.arange(tlen)[None, :]]]
    return [row.tobytes() for row in q], [row.tobytes() for row in t]


def load_peaks():
    """MEASURED_PEAKS.json (driver-written: HBM, bf16) and profiles/peaks_int.json (scripts/peaks_int.py: INT32 pipes,
    int8 tensor) -- measured denominators for every roofline fraction"""
    peaks, ipeaks = {}, {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    try:
        ipeaks = json.load(open(os.path.join(ROOT, "profiles", "peaks_int.json")))
    except Exception:
        pass
    return peaks, ipeaks


def run_realign(ctx, stream, chunks, args, verify):
    """realignment (edlib-compatible Myers kernel): HW + PATH on the two shapes SURVEY.md 8d names -- 1536-base read
    chunks on 1766-base windows, and the in-pipeline 300 x 2300 shape -- each with its CPU baseline and parity check
    against the vendored edlib when `verify` (rank 0 of a 1-GPU run)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from hairsplitter_b200 import api
    from oracle import pyoracle
    cores = host_cores()
    rng = np.random.default_rng(12345)
    _, ipeaks = load_peaks()
    out = {}
    for tag, qlen, slack, n_full in (("chunk_1536x1766", 1536, 0.15, 20000), ("pipeline_300x2300", 300, 6.6667, 60000)):
        n_pairs = max(64, int(n_full * min(1.0, args.scale * 4)))
        qs, ts = make_realign_pairs(rng, chunks[0].contig, n_pairs, qlen=qlen, slack=slack)
        cells = float(sum(len(q) * len(t) for q, t in zip(qs, ts)))
        ctx.edlib_align_batch(qs[:256], ts[:256], k=-1, mode=2, task=2)  # warm-up (allocations, code load)
        # the C ABI's own arguments in pinned host memory: concatenated sequences + offsets in, results out
        def pinned(arr):
            tns = torch.empty(arr.shape, dtype=torch.from_numpy(np.empty(0, arr.dtype)).dtype, pin_memory=True)
            v = tns.numpy()
            v[...] = arr
            return tns, v
        qo = np.zeros(n_pairs + 1, np.int64)
        to = np.zeros(n_pairs + 1, np.int64)
        np.cumsum([len(q) for q in qs], out=qo[1:])
        np.cumsum([len(t) for t in ts], out=to[1:])
        keep = []
        host = []
        for arr in (np.frombuffer(b"".join(qs), np.uint8), qo, np.frombuffer(b"".join(ts), np.uint8), to):
            tns, v = pinned(arr)
            keep.append(tns)
            host.append(v)
        n_loc_cap = 64 * n_pairs + 64
        res_t = torch.zeros(n_pairs * api.EDLIB_RESULT_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
        ends_t = torch.zeros(n_loc_cap, dtype=torch.int32, pin_memory=True)
        starts_t = torch.zeros(n_loc_cap, dtype=torch.int32, pin_memory=True)
        aln_t = torch.zeros(int(qo[-1] + to[-1]) + 8, dtype=torch.uint8, pin_memory=True)
        bufs = (res_t.numpy().view(api.EDLIB_RESULT_DTYPE), ends_t.numpy(), starts_t.numpy(), aln_t.numpy())
        ctx.edlib_align_batch_arrays(*host, k=-1, mode=2, task=2, out=bufs)
        ctx.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_w0 = time.perf_counter()
        e0.record(stream)
        res, ends, starts, aln = ctx.edlib_align_batch_arrays(*host, k=-1, mode=2, task=2, out=bufs)
        e1.record(stream)
        ctx.sync()
        t_w1 = time.perf_counter()
        prof = ctx.profile_report()
        ctx.profile(False)
        ms_e2e = max(e0.elapsed_time(e1), (t_w1 - t_w0) * 1e3)
        ms_kernel = sum(v[1] for kname, v in prof.items() if kname.startswith("edlib_"))
        r = {
            "shape": f"{n_pairs} pairs, query {qlen} x target {len(ts[0])}, HW + PATH (10% error)",
            "kernel_gcups": cells / (ms_kernel * 1e-3) / 1e9, "e2e_gcups": cells / (ms_e2e * 1e-3) / 1e9,
            "kernel_ms": ms_kernel, "e2e_ms": ms_e2e,
            "e2e_path": "hsgpu_edlib_align_batch on pinned host buffers (concatenated sequences + offsets in; results, "
                        "locations and alignments out), every copy inside the timed call",
            "h2d_bytes": int(qo[-1] + to[-1] + 16 * (n_pairs + 1)),
            "d2h_bytes": int(n_pairs * api.EDLIB_RESULT_DTYPE.itemsize + int(res["alignment_length"].sum()) + 8 * int(res["n_locations"].sum())),
            "kernels": {kname: {"launches": v[0], "ms": v[1]} for kname, v in prof.items()},
        }
        # useful logic work: one Myers column step of a 32-row word is ~19 32-bit logic/add operations (SURVEY.md 8d:
        # 14-22 per word step, edlib.cpp:411-446); HW + PATH sweeps the matrix three times (end, start, path)
        useful_ops = cells / 32.0 * 19.0 * 3.0
        if ipeaks.get("int32_peak_tops"):
            r["useful_int32_tops"] = useful_ops / (ms_kernel * 1e-3) / 1e12
            r["frac_of_int32_peak"] = r["useful_int32_tops"] / ipeaks["int32_peak_tops"]
            r["frac_of_alu_pipe_peak"] = r["useful_int32_tops"] / ipeaks["int32_alu_pipe_tops"]
            r["int32_peak_source"] = "profiles/peaks_int.json (lop3+imad chains: both integer pipes; lop3 alone: ALU pipe)"
        if verify:
            n_cpu = min(n_pairs, 250 * cores)
            t0 = time.perf_counter()
            if pyoracle.RefEdlib.available():
                kind = "reference"
                with ThreadPoolExecutor(max_workers=cores) as ex:
                    cpu = list(ex.map(lambda i: pyoracle.RefEdlib.align(qs[i], ts[i], -1, 2, 2), range(n_cpu)))
            else:
                kind = "port"
                O = pyoracle.Oracle()
                n_cpu = min(n_cpu, 8 * cores)
                with ThreadPoolExecutor(max_workers=cores) as ex:
                    cpu = list(ex.map(lambda i: O.edlib_align(qs[i], ts[i], -1, 2, 2), range(n_cpu)))
            dt = time.perf_counter() - t0
            for i in range(n_cpu):  # parity of the timed results
                rr = res[i]
                assert int(rr["edit_distance"]) == cpu[i]["edit_distance"]
                lo, nl = int(rr["loc_off"]), int(rr["n_locations"])
                assert np.array_equal(ends[lo:lo + nl], cpu[i]["end_locations"])
                assert np.array_equal(starts[lo:lo + nl], cpu[i]["start_locations"])
                ao, al = int(rr["aln_off"]), int(rr["alignment_length"])
                assert np.array_equal(aln[ao:ao + al], cpu[i]["alignment"])
            cpu_cells = float(sum(len(qs[i]) * len(ts[i]) for i in range(n_cpu)))
            r["cpu_baseline"] = {"value": cpu_cells / dt / 1e9, "unit": "GCUPS", "cores": cores, "kind": kind,
                                 "sample": f"first {n_cpu} pairs, edlibAlign HW+PATH, one pair per thread at a time",
                                 "verified_pairs": n_cpu}
        r["cells"] = cells
        out[tag] = r
    out["metric"] = "realign_gcups"
    out["unit"] = "GCUPS (|query| x |target| cells per pair, full matrix, counted once)"
    return out


def run_stages(ctx, stream, chunks, args, hbm_peak):
    """the partition x column contingency filter on one chunk against the reference's own snps_out, and the read x read
    counts; both verified against the reference inside the run when oracle/_ref is present."""
    import torch
    from hairsplitter_b200 import api
    from oracle import pyoracle
    out = {}
    cores = host_cores()
    rng = np.random.default_rng(12345)

    # ---- contingency: loops 3+4 of keep_only_robust_variants on one 300 kb chunk ----
    cb = chunks[0]
    pk = api.PackedBatch([cb])
    pu = api.Pileup(ctx, pk)
    pu.build()
    pu.column_rank()
    pos, _ = pu.suspects(0)
    parts = None
    cpu_s = None
    filt = None
    if pyoracle.ref_available():
        R = pyoracle.RefCV(cb)
        rc = R.call_variants()
        t0 = time.perf_counter()
        parts, filt, merged = R.robust()   # the reference's keep_only_robust_variants (85 % of its run time)
        cpu_s = time.perf_counter() - t0
        assert np.array_equal(rc["suspects"]["pos"], pos)
    else:
        # no compiled reference on this box: partitions from the strain of origin of the reads
        ends_ = pu.read_ends()
        parts = []
        for w0 in range(0, cb.length, 3000):
            idx = np.nonzero((cb.start < w0 + 3000) & (ends_ > w0))[0].astype(np.int32)
            st = np.where(cb.strain[idx] == 0, 1, -1).astype(np.int16)
            parts.append(dict(read_idx=idx, state=st, more=np.full(idx.size, 3, np.int32), less=np.zeros(idx.size, np.int32)))
    pu.robust_filter(0, parts, pos)  # warm-up
    ctx.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    kept = pu.robust_filter(0, parts, pos)
    e1.record(stream)
    ctx.sync()
    prof = ctx.profile_report()
    ctx.profile(False)
    if filt is not None:
        assert np.array_equal(kept, filt["pos"]), "snps_out differs from the reference"
    cells0, _, _ = pu.stats()
    kms = sum(v[1] for kname, v in prof.items() if kname.startswith("robust_filter_"))  # lanes kernel + overflow / deep
    out["contingency"] = {
        "metric": "robust_filter (loops 3+4 of keep_only_robust_variants) on one chunk",
        "shape": f"{cb.length} columns, {cb.n_reads} reads, {len(parts)} partitions, {pos.size} suspects -> {kept.size} kept",
        "e2e_ms": e0.elapsed_time(e1), "kernel_ms": kms,
        "kernel_gbs_algorithmic": (float(cells0[0]) + 3 * cb.length) / (kms * 1e-3) / 1e9 if kms else None,
        "frac_of_hbm_peak": ((float(cells0[0]) + 3 * cb.length) / (kms * 1e-3) / 1e9 / hbm_peak) if kms else None,
        "verified_against_reference": filt is not None,
        "cpu_baseline": ({"value": cpu_s, "unit": "s for the whole keep_only_robust_variants (loops 1-4) on the same chunk",
                          "cores": 1, "kind": "reference"} if cpu_s is not None else None),
    }
    pu.close()

    # ---- read x read counts: list_similarities_and_differences_between_reads3 for every chunk ----
    # SNP columns = the suspect columns our own pipeline finds on each chunk (what the .col file hands to
    # HS_separate_reads); all chunks go through one hsgpu_pairs batch.
    pk = api.PackedBatch(chunks)
    pu = api.Pileup(ctx, pk)
    pu.build()
    pu.column_rank()
    cols = []
    for ci, cb in enumerate(chunks):
        pos, _ = pu.suspects(ci)
        off, idx, code = pu.extract_columns(ci, pos)
        summ = pu.column_summary(ci)
        cols.append((cb.n_reads, off, idx, code, summ["ref_base"][pos], summ["second_base"][pos]))
    pu.close()
    peaks, ipeaks = load_peaks()
    # int8 tensor peak: measured with cuBLASLt (torch._int_mm 8192^3) by scripts/peaks_int.py; only if that file is
    # missing, 2 x the measured bf16 burst figure (nominal dense int8 is 2 x bf16)
    if ipeaks.get("int8_peak_tops"):
        int8_peak_tops = float(ipeaks["int8_peak_tops"])
        int8_src = "measured: cuBLASLt int8 GEMM 8192^3 (profiles/peaks_int.json)"
    else:
        int8_peak_tops = 2.0 * float(peaks.get("bf16_tflops", 1590.0))
        int8_src = "assumed 2 x measured bf16 burst (profiles/peaks_int.json missing)"
    pairs = {}
    for tag, flags in (("band", 0), ("dense", api.PAIRS_DENSE)):
        P = api.Pairs(ctx, cols, flags)
        info = P.info()
        for _ in range(3):
            P.compute()
        ctx.sync()
        ctx.profile(True)
        reps = 10 if tag == "band" else 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            P.compute()
        e1.record(stream)
        ctx.sync()
        prof = ctx.profile_report()
        ctx.profile(False)
        kms = prof.get("pair_umma_kernel", (0, 0.0))[1] / reps
        # every 128-SNP block of a tile pair = 2 x 4 tcgen05.mma of M128 x N256 x K32
        macs = info["kblocks"] * 128.0 * 2 * 128 * 256
        pairs[tag] = {
            "ms_per_batch": e0.elapsed_time(e1) / reps, "kernel_ms": kms,
            "tile_pairs": info["tile_pairs"], "snp_blocks": info["kblocks"],
            "executed_tops": 2 * macs / (kms * 1e-3) / 1e12 if kms else None,
            "frac_of_int8_peak": (2 * macs / (kms * 1e-3) / 1e12 / int8_peak_tops) if kms else None,
            "output_gbs": 2 * info["out_elems"] * 4 / (e0.elapsed_time(e1) / reps * 1e-3) / 1e9,
        }
        if tag == "band":
            n_chk = min(2, len(cols))
            cpu_t = None
            for ci in range(n_chk):  # parity of the timed results against the reference's Eigen code
                sim, diff = P.fetch(ci)
                if pyoracle.RefSR.available():
                    t0 = time.perf_counter()
                    rs, rd = pyoracle.RefSR.read_pair_counts(*cols[ci])
                    cpu_t = (cpu_t or 0.0) + time.perf_counter() - t0
                else:
                    rs, rd = pyoracle.Oracle().read_pair_counts(*cols[ci])
                assert np.array_equal(sim, rs) and np.array_equal(diff, rd), "read-pair counts differ from the reference"
            pairs["verified_chunks"] = n_chk
            pairs["cpu_baseline"] = ({"value": cpu_t / n_chk * 1e3, "unit": "ms per chunk (Eigen sparse products + densify)",
                                      "cores": 1, "kind": "reference"} if cpu_t else None)
            pairs["identity_order"] = bool(info["identity"])
            pairs["dense_equivalent_macs"] = 4.0 * sum(float(c[0]) ** 2 * (c[1].size - 1) for c in cols)
        P.close()
    n_snps = sum(c[1].size - 1 for c in cols)
    out["read_pairs"] = {
        "metric": "list_similarities_and_differences_between_reads3 over all chunks (one hsgpu_pairs batch)",
        "shape": f"{len(cols)} chunks, {sum(c[0] for c in cols)} reads, {n_snps} SNP columns, "
                 f"{sum(int(c[1][-1]) for c in cols)} cells",
        "int8_peak_tops": int8_peak_tops,
        "int8_peak_source": int8_src,
        **pairs,
    }
    return out


def col_blocks(path):
    """contig name -> its block of a .col file (the reference's contig order depends on its thread schedule)"""
    blocks, cur = {}, None
    for line in open(path):
        if line.startswith("CONTIG\t"):
            cur = line.split("\t")[1]
            blocks[cur] = []
        if cur is not None and line.strip():
            blocks[cur].append(line)
    return blocks


def run_separate_reads_wall(tmp, col, cores, n_contigs):
    """wall time of the HS_separate_reads executable on the .col of the whole configuration: ours (host C++ over
    libhsgpu: read-pair counts, read graphs and chinese-whispers runs on one GPU) beside the reference's (OpenMP over
    contigs). Both with std::random_device pinned to the same constant (HS_PIN_SEED / oracle/ref_pin_rng.cpp), so
    the .gro files can be compared contig by contig."""
    import subprocess
    ours = os.path.join(ROOT, "hairsplitter_b200", "bin", "HS_separate_reads")
    ref = os.path.join(ROOT, "oracle", "_ref", "HS_separate_reads_pinned")
    if not os.path.exists(ours):
        return {"unavailable": "hairsplitter_b200/bin/HS_separate_reads not built"}
    from oracle.pyoracle import PIN_SEED
    env = dict(os.environ, HS_PIN_SEED=str(PIN_SEED), HS_TIMING="1")

    def run(exe, tag, threads):
        gro = os.path.join(tmp, f"{tag}.gro")
        t0 = time.perf_counter()
        r = subprocess.run([exe, col, str(threads), "0.1", os.path.join(tmp, "no_ploidy"), "0", "0", "0", gro, "0"], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
        return time.perf_counter() - t0, gro, r.stderr.decode()

    run(ours, "sr_warm", cores)
    # CUDA context creation inside the child process took anything from 0.3 to 2.7 s from run to run on the measured
    # boxes (driver side, not ours to tune): three runs, the MEDIAN is reported, all are listed
    runs = [run(ours, "sr_ours", cores) for _ in range(3)]
    t_ours, gro_ours, log = sorted(runs, key=lambda r: r[0])[1]
    out = {"metric": "HS_separate_reads wall time (parse .col + read-pair counts + read graphs + clustering + write .gro)",
           "ours_s": t_ours, "ours_runs_s": [round(r[0], 3) for r in runs], "ours_threads": cores, "ours_gpus": 1,
           "phases": [l.replace("[hs timing]", "").strip() for l in log.splitlines() if l.startswith("[hs timing]")]}
    if os.path.exists(ref):
        threads = min(cores, n_contigs)
        ref_runs = [run(ref, "sr_ref", threads) for _ in range(3)]
        t_ref, gro_ref, _ = sorted(ref_runs, key=lambda r: r[0])[1]
        out["reference_runs_s"] = [round(r[0], 3) for r in ref_runs]
        a, b = col_blocks(gro_ours), col_blocks(gro_ref)
        same = a == b
        out.update({"reference_s": t_ref, "reference_threads": threads, "speedup": t_ref / t_ours,
                    "gro_identical_to_pinned_reference": same,
                    "groups": sum(1 for v in a.values() for l in v if l.startswith("GROUP"))})
        assert same, "our .gro differs from the pinned reference's"
    return out


def run_call_variants_wall(chunks, args):
    """BASELINE's third headline: wall time of the HS_call_variants executable on the whole configuration, ours
    (host C++ over libhsgpu, one GPU) beside the reference's (OpenMP over contigs, all host cores), same files,
    outputs compared contig by contig."""
    import shutil
    import subprocess
    import tempfile
    from hairsplitter_b200 import synth
    ours = os.path.join(ROOT, "hairsplitter_b200", "bin", "HS_call_variants")
    ref = os.path.join(ROOT, "oracle", "_ref", "HS_call_variants")
    if not os.path.exists(ours):
        return {"unavailable": "hairsplitter_b200/bin/HS_call_variants not built"}
    cores = host_cores()
    sample = chunks if args.wall_chunks <= 0 else chunks[:args.wall_chunks]
    tmp = tempfile.mkdtemp(prefix="hs_wall_")
    try:
        t0 = time.perf_counter()
        gfa, reads, sam = synth.write_files(sample, os.path.join(tmp, "in"), links=getattr(args, "links", ()))
        t_write = time.perf_counter() - t0

        def run(exe, tag, threads):
            col, vcf, err = [os.path.join(tmp, f"{tag}.{e}") for e in ("col", "vcf", "err")]
            t0 = time.perf_counter()
            subprocess.run([exe, gfa, reads, sam, str(threads), tmp, err, "0", "0", col, vcf, "0.33"], check=True,
                           stdout=subprocess.DEVNULL)
            return time.perf_counter() - t0, col, err

        run(ours, "warm", cores)  # warms the file cache
        # three runs, median reported, all listed: CUDA context creation in the child varies between 0.3 and 2.7 s
        runs = [run(ours, "ours", cores) for _ in range(3)]
        t_ours, col_ours, err_ours = sorted(runs, key=lambda r: r[0])[1]
        out = {
            "metric": "HS_call_variants wall time (parse SAM/FASTA/GFA + pileup + variant calling + robust filter + write .col/.vcf)",
            "sample": f"{len(sample)} contig chunks, {sum(c.length for c in sample)} columns, {sum(c.n_reads for c in sample)} reads; "
                      f"input files {sum(os.path.getsize(f) for f in (gfa, reads, sam)) / 1e6:.0f} MB",
            "ours_s": t_ours, "ours_runs_s": [round(r[0], 3) for r in runs], "ours_threads": cores, "ours_gpus": 1,
            "write_inputs_s": round(t_write, 1),
        }
        if os.path.exists(ref):
            threads = min(cores, len(sample))
            ref_runs = [run(ref, "ref", threads) for _ in range(max(1, args.wall_ref_runs))]
            t_ref, col_ref, err_ref = sorted(ref_runs, key=lambda r: r[0])[len(ref_runs) // 2]
            a, b = col_blocks(col_ours), col_blocks(col_ref)
            assert a == b, "our .col differs from the reference's"
            assert open(err_ours).read() == open(err_ref).read() or threads > 1  # float sum order varies with threads
            out.update({"reference_s": t_ref, "reference_runs_s": [round(r[0], 3) for r in ref_runs], "reference_threads": threads,
                        "reported": "median of the runs listed, both arms",
                        "speedup": t_ref / t_ours, "col_identical_to_reference": True,
                        "snps": sum(len(v) for v in a.values())})
        out["separate_reads"] = run_separate_reads_wall(tmp, col_ours, cores, len(sample))
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_e2e(chunks, parts_per_contig, local_rank, ctx, n_lanes, n_groups, steps, warmup, barrier):
    """The way a caller drives the library from host buffers: every step's batch (optionally cut into `n_groups`
    groups of contig chunks) goes through one of `n_lanes` hsgpu contexts (one per host thread, as the header
    prescribes, like the reference's OpenMP loop over contigs), steps dealt to the lanes in turn. The library lets
    one context upload at a time, so the kernels of the batch that has its data overlap the upload of the next
    one (double buffering); CIGARs travel in the 8-bit form; the final partitions go up through
    hsgpu_partitions_set; every batch's results (snps_out, suspect lists, depth numerators) come back through
    hsgpu_robust_filter_all and hsgpu_suspects_all. Every step pays its own H2D and D2H; read / CIGAR inputs sit
    in pinned host memory. Returns the device time of the slowest lane and the host wall clock over `steps` steps."""
    import threading
    import torch
    from hairsplitter_b200 import api
    n_groups = max(1, min(n_groups, len(chunks)))
    n_lanes = max(1, n_lanes)
    groups = []
    keep = []
    group_parts = []
    for g in range(n_groups):
        group_parts.append(api.Pileup.prepare_partitions(parts_per_contig[g::n_groups]))
        pb = api.PackedBatch(chunks[g::n_groups]).use_cigar8()
        for name in ("contig_len", "contig_bases", "contig_word_off", "contig_read_off", "read_bases", "read_word_off",
                     "read_len", "cigar8", "cigar8_off", "read_start", "read_strand"):
            src = getattr(pb, name)
            t = torch.empty(max(src.nbytes, 1), dtype=torch.uint8, pin_memory=True)
            arr = t.numpy()[:src.nbytes].view(src.dtype).reshape(src.shape)
            arr[...] = src
            keep.append(t)
            setattr(pb, name, arr)
        groups.append(pb)
    h2d_bytes = sum(int(pb.input_bytes) for pb in groups) + sum(int(gp[2]) for gp in group_parts)
    e2e_kept = [0] * n_groups
    lanes = [ctx] + [api.Context(local_rank) for _ in range(n_lanes - 1)]
    lane_streams = [torch.cuda.ExternalStream(c.stream(), device=torch.device("cuda", local_rank)) for c in lanes]
    e2e_out = [0] * n_groups
    e2e_sus = [0] * n_groups

    trace = [] if os.environ.get("HS_E2E_TRACE") else None

    # result buffers of every lane, kept from step to step (pinned): a caller in a loop does not allocate per batch,
    # and large numpy arrays allocated and freed per call mean an mmap / munmap pair each -- with several threads in
    # the driver at the same time that showed up as stalls of every lane
    def pinned_array(n, dtype):
        tns = torch.empty(max(int(n), 1) * np.dtype(dtype).itemsize, dtype=torch.uint8, pin_memory=True)
        keep.append(tns)
        return tns.numpy().view(dtype)
    nc_max = max(int(pb.contig_len.shape[0]) for pb in groups)
    cols_max = max(int(pb.contig_len.astype(np.int64).sum()) for pb in groups)
    lane_out = []
    for _ in range(n_lanes):
        lane_out.append({
            "kept": (pinned_array(cols_max // 4 + 1024, np.int32), pinned_array(nc_max + 1, np.int64)),
            "sus": (pinned_array(cols_max // 6 + 2 * nc_max, np.int32), pinned_array(cols_max // 6 + 2 * nc_max, np.uint8),
                    pinned_array(nc_max + 1, np.int64), pinned_array(nc_max, np.int64)),
        })

    def e2e_group(lane, g):
        t0 = time.perf_counter()
        p = api.Pileup(lanes[lane], groups[g])   # H2D of the group
        t1 = time.perf_counter()
        p.build()
        t2 = time.perf_counter()
        p.column_rank()
        t3 = time.perf_counter()
        p.partitions_set(prepared=group_parts[g])   # H2D of the final partitions
        t3a = time.perf_counter()
        nc_g = int(groups[g].contig_len.shape[0])
        kept, koff = p.robust_filter_all(out=lane_out[lane]["kept"])   # loops 3+4 + D2H of snps_out
        koff = koff[: nc_g + 1]
        t3b = time.perf_counter()
        pos, au, off, ds = p.suspects_all(out=lane_out[lane]["sus"])   # D2H of the call_variants results
        off, ds = off[: nc_g + 1], ds[:nc_g]
        t4 = time.perf_counter()
        p.close()
        if trace is not None:
            trace.append((lane, g, t0, t1, t2, t3, t3a, t3b, t4, time.perf_counter()))
        e2e_out[g] = pos.nbytes + au.nbytes + off.nbytes + ds.nbytes + kept.nbytes + koff.nbytes
        e2e_sus[g] = int(off[-1])
        e2e_kept[g] = int(koff[-1])

    def e2e_steps(n_steps):
        # work items = (step, group) in order, dealt round-robin to the lanes: with one group per step and two
        # lanes this is plain double buffering -- the upload of step k+1 overlaps the kernels of step k
        items = [(st, g) for st in range(n_steps) for g in range(n_groups)]

        def work(lane):
            for i in range(lane, len(items), n_lanes):
                e2e_group(lane, items[i][1])
        ts = [threading.Thread(target=work, args=(lane,)) for lane in range(1, n_lanes)]
        for t in ts:
            t.start()
        work(0)
        for t in ts:
            t.join()
        return sum(e2e_out)

    # every lane runs at least two untimed steps: its context's memory pool and staging buffers reach their size there
    # (a lane that met its first batch inside the timed region paid the physical allocations of its pool: 10-100 ms)
    d2h_bytes = e2e_steps(max(warmup, 2 * n_lanes))
    barrier()
    t_e2e = time.perf_counter()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in lanes]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in lanes]
    for ev, st_ in zip(e0, lane_streams):
        ev.record(st_)
    d2h_bytes = e2e_steps(steps)
    for ev, st_ in zip(e1, lane_streams):
        ev.record(st_)
    for c in lanes:
        c.sync()
    barrier()
    t_e2e = (time.perf_counter() - t_e2e) * 1e3
    # device time of the slowest lane; the host wall clock around the same region is reported beside it
    e2e_ms = max(a.elapsed_time(b) for a, b in zip(e0, e1))
    for c in lanes[1:]:
        c.close()
    if trace:  # the last step, times in ms from its first call
        last = sorted(trace[-max(n_groups, 9):], key=lambda r: r[2])
        z = last[0][2]
        for r in last:
            print("  lane %d group %2d: create %.3f-%.3f build -%.3f rank -%.3f partitions_set -%.3f filter_all -%.3f suspects_all -%.3f close -%.3f" %
                  ((r[0], r[1]) + tuple((x - z) * 1e3 for x in r[2:])), file=sys.stderr)
    return {"device_ms": e2e_ms, "wall_ms": t_e2e, "h2d_bytes": h2d_bytes, "d2h_bytes": d2h_bytes,
            "suspects": sum(e2e_sus), "kept": sum(e2e_kept), "groups": n_groups, "lanes": n_lanes}


def workload_description(info, chunks):
    return f"BASELINE configs[{info['config'] - 1}]: {info['description']}; {len(chunks)} contig chunks <= 300 kb"


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path, all host threads."""
    if rank != 0:
        return
    from hairsplitter_b200 import synth
    cores = host_cores()
    n_sample = args.cpu_sample_chunks or max(1, min(cores, 17))  # one chunk per host thread, up to the whole workload
    chunks, info = synth.make_config(args.config, scale=args.scale, seed=args.config, n_chunks=n_sample)
    n_threads = min(cores, len(chunks))
    for _ in range(args.warmup if args.warmup < 1 else 1):  # one warm-up pass is enough for a CPU code path
        cpu_reference_rate(chunks[:n_threads], n_threads)
    t_total, cols = 0.0, 0
    kind = "port"
    for _ in range(args.steps):
        rate, kind, dt = cpu_reference_rate(chunks, n_threads)
        t_total += dt
        cols += sum(c.length for c in chunks)
    value = cols / WINDOW / t_total
    sample = (f"{len(chunks)} of the workload's chunks ({sum(c.length for c in chunks)} columns, "
              f"{sum(c.n_reads for c in chunks)} reads) per step, generate_msa + call_variants + keep_only_robust_variants per chunk, "
              f"one chunk per thread")
    line = {
        "impl": "reference", "metric": "pileup_windows_per_s", "value": value, "unit": "windows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_description(info, chunks) if args.scale == 1.0 else f"scaled x{args.scale}",
                   "window": "2000 columns x depth", "timing": "host wall clock (CPU code path)"},
        "cpu_baseline": {"value": value, "unit": "windows/s", "cores": n_threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from hairsplitter_b200 import api, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- workload ----
    from hairsplitter_b200 import sharding
    t_gen = time.perf_counter()
    workers = max(1, host_cores() // max(world, 1))
    lpt = None
    if args.strong:
        # ONE workload, the same on every rank, dealt out chunk by chunk: heaviest first onto the least loaded rank
        all_chunks, info = synth.make_config(args.config, scale=args.scale, seed=args.config, workers=workers)
        weights = [sharding.chunk_weight(c) for c in all_chunks]
        bins = sharding.lpt_assign(weights, world)
        loads = [sum(weights[i] for i in b) for b in bins]
        lpt = {"chunks_total": len(all_chunks), "chunks_per_rank": [len(b) for b in bins],
               "imbalance_max_over_mean": max(loads) / (sum(loads) / len(loads))}
        chunks = [all_chunks[i] for i in bins[rank]]
        del all_chunks
    else:
        # every rank gets its own genome of the configured shape (weak scaling)
        chunks, info = synth.make_config(args.config, scale=args.scale, seed=args.config + 1000 * rank, workers=workers)
    packed = api.PackedBatch(chunks)
    t_gen = time.perf_counter() - t_gen
    n_cols = int(packed.contig_len.sum())
    ctx = api.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))
    if args.only_realign:
        realign = run_realign(ctx, stream, chunks, args, verify=(rank == 0 and world == 1))
        if rank == 0:
            print(json.dumps({"only_realign": True, "bpl": os.environ.get("HSGPU_EDLIB_BPL"), "realign": realign}))
        return

    # ---- the final partitions of every chunk (loops 1-2 of keep_only_robust_variants: sequential host C++, an INPUT
    # of the C ABI's filter stage): built once from a first build + rank, outside every timed region ----
    pu = api.Pileup(ctx, packed)
    pu.build()
    pu.column_rank()
    t_parts = time.perf_counter()
    parts_per_contig = [pu.host_partitions(ci) for ci in range(len(chunks))]
    t_parts = time.perf_counter() - t_parts
    prepared = api.Pileup.prepare_partitions(parts_per_contig)
    n_parts = sum(len(x) for x in parts_per_contig)
    pu.partitions_set(prepared=prepared)
    kept_cap = n_cols

    step_trace = [] if os.environ.get("HS_STEP_TRACE") else None

    def step(pu):
        t0 = time.perf_counter()
        pu.build()                                 # generate_msa
        t1 = time.perf_counter()
        pu.column_rank()                           # call_variants
        t2 = time.perf_counter()
        out = pu.robust_filter_all(kept_cap)       # loops 3+4 of keep_only_robust_variants -> snps_out
        if step_trace is not None:
            step_trace.append((t1 - t0, t2 - t1, time.perf_counter() - t2))
        return out

    # ---- value: inputs resident in HBM ----
    for _ in range(max(args.warmup, 0)):
        step(pu)
    ctx.sync()
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    l0 = ctx.launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        kept, kept_off = step(pu)
    ev1.record(stream)
    ctx.sync()
    barrier()
    if step_trace:
        for tr in step_trace[-args.steps:]:
            print("  step host ms: build %.2f rank %.2f filter_all %.2f" % tuple(x * 1e3 for x in tr), file=sys.stderr)
    clk = clocks.stop()
    launches = ctx.launches() - l0
    ms_total = ev0.elapsed_time(ev1)
    cells, dist_sum, alen = pu.stats()
    n_sus, depth_sum = pu.column_counts()
    assert int(depth_sum.sum()) == int(cells.sum()), "depth numerator must equal the number of pileup cells"
    assert int(n_sus.sum()) > 0 or n_cols < 100000
    n_kept = int(kept_off[-1])

    # ---- per-kernel timing over a second pass of the same steps ----
    ctx.profile(True)
    for _ in range(args.steps):
        step(pu)
    prof = ctx.profile_report()
    ctx.profile(False)
    n_cells = int(cells.sum())
    finfo = pu.filter_info()
    alg_bytes = {
        # 1 B code out + 2-bit read base + 2-bit contig base per cell, ~0.1 B of CIGAR per cell (the ops
        # as the byte-sized form the kernel reads), 48 B of metadata per read
        "pileup_kernel": n_cells * 1.0 + packed.read_bases.nbytes + n_cells * 0.25 + float(pu.cigar_bytes())
        + packed.n_reads * 48,
        # 1 B code in per cell (+ tile padding not counted), 4 B per (tile, read) index entry, 19 B summary out per column
        "column_rank_kernel": n_cells * 1.0 + (n_cells / 128.0 + packed.n_reads) * 4 + n_cols * 19.0,
        # the cells of the active columns: 1 B code + 2 B read index each, one state byte per cell and
        # partition that holds the read, 1 B kept flag out per column
        "robust_filter_kernel": finfo["active_cells"] * 3.0 + finfo["active_cells"] * finfo["parts_per_cell"] + n_cols * 1.0,
    }
    # the same bytes whichever kernel instance does the ordinary depths (one lane per partition by default)
    alg_bytes["robust_filter_lanes_kernel"] = alg_bytes["robust_filter_kernel"]
    peaks, ipeaks = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    kernels = []
    for name, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        k = {"name": name, "launches_per_step": n / max(args.steps, 1), "ms_per_step": ms / max(args.steps, 1)}
        if name in alg_bytes and n:
            k["alg_bytes_per_launch"] = alg_bytes[name]
            k["achieved_gbs"] = alg_bytes[name] / (ms / n * 1e-3) / 1e9
            k["frac_of_hbm_peak"] = k["achieved_gbs"] / hbm_peak
        kernels.append(k)
    # the robust filter runs as two launches of one template (ordinary tiles, then tiles deeper than its shared-memory
    # stage); the active cells are counted over both, so the pair is one roofline entry
    fam = [k for k in kernels if k["name"].startswith("robust_filter_") and "overflow" not in k["name"]]
    if len(fam) == 2:
        ms_f = sum(k["ms_per_step"] for k in fam)
        both = {"name": fam[0]["name"].split("<")[0] + " + " + fam[1]["name"], "ms_per_step": ms_f, "launches_per_step": 2.0,
                "alg_bytes_per_launch": alg_bytes["robust_filter_kernel"],
                "achieved_gbs": alg_bytes["robust_filter_kernel"] / (ms_f * 1e-3) / 1e9}
        for k in fam:
            for key in ("alg_bytes_per_launch", "achieved_gbs", "frac_of_hbm_peak"):
                k.pop(key, None)
            k["note"] = "bytes are counted over both launches of the template: see roofline"
    else:
        both = None
    dom = kernels[0] if kernels else None
    if both and dom and (dom in fam or both["ms_per_step"] > dom["ms_per_step"]):
        dom = both
    if dom and "achieved_gbs" not in dom:  # a helper kernel on top (tiny workloads): the largest kernel with a byte model
        dom = next((k for k in kernels if "achieved_gbs" in k), None)
    roofline = None
    if dom and "achieved_gbs" in dom:
        tr = traffic.get(dom["name"])
        roofline = {"kernel": dom["name"], "bound": "hbm", "achieved": dom["achieved_gbs"], "peak": hbm_peak,
                    "unit": "GB/s", "frac": dom["achieved_gbs"] / hbm_peak, "peak_source": peak_src,
                    "traffic": tr.get("bytes") if isinstance(tr, dict) else tr,
                    "traffic_source": (tr.get("source") if isinstance(tr, dict) else
                                       "profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture)") if tr else None,
                    "share_of_step": dom["ms_per_step"] / sum(k["ms_per_step"] for k in kernels if k is not both),
                    "note": "per-launch duration from CUDA events on the launching stream, second pass of the same steps"}
    if roofline and roofline["kernel"].startswith("robust_filter"):
        # not a streaming kernel: its work is one 2-bit state per (active cell, partition that holds a read of the column)
        pairs = float(finfo["state_bytes"])
        roofline["work"] = {"cell_partition_pairs_per_launch": pairs,
                            "pairs_per_s": pairs / (dom["ms_per_step"] * 1e-3),
                            "note": "latency / issue bound (ncu: profiles/r02w_ncu_filter_lanes.txt); the HBM fraction is "
                                    "reported because the contract asks for it"}
    pu.close()

    # ---- e2e: host buffers -> C ABI -> host results, every step ----
    e2e_lanes = args.e2e_lanes if args.e2e_lanes > 0 else max(2, min(5, host_cores() // max(world, 1) - 1))
    e2e = run_e2e(chunks, parts_per_contig, local_rank, ctx, e2e_lanes, args.e2e_groups, args.steps, args.warmup, barrier)
    assert e2e["suspects"] == int(n_sus.sum()), "the e2e path must find the same suspect columns"
    assert e2e["kept"] == n_kept, "the e2e path must keep the same columns"
    e2e_ms, t_e2e, h2d_bytes, d2h_bytes = e2e["device_ms"], e2e["wall_ms"], e2e["h2d_bytes"], e2e["d2h_bytes"]
    n_groups, n_lanes = e2e["groups"], e2e["lanes"]
    # the floor of the e2e step on this box: the same bytes from pinned memory -- first with every rank copying at
    # the same moment (what the ranks of a multi-GPU run do to the host's memory system and PCIe uplinks), then alone
    src = torch.empty(h2d_bytes, dtype=torch.uint8, pin_memory=True)
    dst = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()

    def copy_floor():
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        return c0.elapsed_time(c1) / 3

    barrier()
    h2d_floor_concurrent_ms = copy_floor()
    barrier()
    h2d_floor_ms = h2d_floor_concurrent_ms
    if world > 1:
        for turn in range(world):  # one rank at a time
            if turn == rank:
                h2d_floor_ms = copy_floor()
            barrier()
    del src, dst

    # ---- realignment: measured on every rank (BASELINE metric: realign GCUPS at 1/2/4/8 GPUs) ----
    realign = run_realign(ctx, stream, chunks, args, verify=(rank == 0 and world == 1))
    # ---- the other stages of the hot path (reported beside the headline, rank 0 / 1 GPU only) ----
    stages = {"realign": realign}
    if rank == 0 and world == 1 and not args.no_stages:
        stages.update(run_stages(ctx, stream, chunks, args, hbm_peak))
        if args.wall_chunks >= 0:
            args.links = info.get("links", [])
            stages["call_variants_wall"] = run_call_variants_wall(chunks, args)

    # ---- max over ranks of the time, sum over ranks of the columns (hairsplitter_b200/sharding.py) ----
    ms_total, total_cols = sharding.reduce_step(ms_total, float(n_cols), device="cuda")
    e2e_ms, _ = sharding.reduce_step(e2e_ms, 0.0, device="cuda")
    h2d_floor_concurrent_ms, _ = sharding.reduce_step(h2d_floor_concurrent_ms, 0.0, device="cuda")
    realign_line = {}
    for tag in ("chunk_1536x1766", "pipeline_300x2300"):
        k_ms, cells_all = sharding.reduce_step(realign[tag]["kernel_ms"], realign[tag]["cells"], device="cuda")
        e_ms, _ = sharding.reduce_step(realign[tag]["e2e_ms"], 0.0, device="cuda")
        realign_line[tag] = {"kernel_gcups": cells_all / (k_ms * 1e-3) / 1e9, "e2e_gcups": cells_all / (e_ms * 1e-3) / 1e9}
        for key in ("frac_of_int32_peak", "frac_of_alu_pipe_peak"):
            if key in realign[tag]:
                realign_line[tag][key + "_rank0"] = realign[tag][key]

    if rank == 0:
        value = total_cols * args.steps / WINDOW / (ms_total * 1e-3)
        e2e_value = total_cols * args.steps / WINDOW / (e2e_ms * 1e-3)
        line = {
            "metric": "pileup_windows_per_s", "value": value, "unit": "windows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / max(args.steps, 1),
            "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {
                "workload": workload_description(info, chunks) if args.scale == 1.0 else
                            f"BASELINE configs[{args.config - 1}] scaled x{args.scale}: {info['description']}",
                "window": "2000 columns x depth", "columns_per_gpu": n_cols, "reads_per_gpu": packed.n_reads,
                "cells_per_gpu": n_cells, "partitions_per_gpu": n_parts, "suspects_per_gpu": int(n_sus.sum()),
                "snps_out_per_gpu": n_kept,
                "step": "hsgpu_pileup_build + hsgpu_column_rank + hsgpu_robust_filter_all (generate_msa + call_variants "
                        "+ loops 3-4 of keep_only_robust_variants on resident partitions)",
                "l2": "inputs + pileup per step (%.0f MB) exceed the 126 MB L2; no explicit flush" %
                      ((packed.input_bytes + n_cells) / 1e6),
                "sharding": ("one workload dealt to the ranks by longest-processing-time-first, no data-path collective"
                             if args.strong else "independent contig chunks per rank, no data-path collective"),
                "generation_s": round(t_gen, 1), "host_partitions_s": round(t_parts, 2),
            },
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "windows/s", "ms_per_step": e2e_ms / max(args.steps, 1),
                    "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                    "host_wall_ms_per_step": t_e2e / max(args.steps, 1),
                    "h2d_copy_floor_ms": h2d_floor_ms, "h2d_gbs": h2d_bytes / (h2d_floor_ms * 1e-3) / 1e9,
                    "h2d_copy_floor_all_ranks_at_once_ms": h2d_floor_concurrent_ms,
                    "h2d_gbs_per_gpu_all_ranks_at_once": h2d_bytes / (h2d_floor_concurrent_ms * 1e-3) / 1e9,
                    "path": f"steps dealt in turn to {n_lanes} hsgpu contexts (one host thread each), {n_groups} group(s) of "
                            "chunks per step: hsgpu_pileup_create(pinned host buffers, 8-bit CIGAR; one upload at a "
                            "time, so the upload of step k+1 overlaps the kernels of step k) + build + column_rank + "
                            "partitions_set + robust_filter_all + suspects_all; every step pays its own H2D and D2H"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "realign_gcups": realign_line,
            "kernels": kernels,
            "stages": stages,
        }
        if lpt is not None:
            line["strong_scaling"] = lpt
        if world == 1:
            cores = host_cores()
            n_sample = args.cpu_sample_chunks or max(1, min(cores, len(chunks)))  # every host core gets a chunk
            sample = chunks[:n_sample]
            n_threads = min(cores, len(sample))
            rate, kind, dt = cpu_reference_rate(sample, n_threads)
            line["cpu_baseline"] = {
                "value": rate, "unit": "windows/s", "cores": n_threads, "kind": kind, "seconds": round(dt, 2),
                "sample": f"first {len(sample)} chunks of the workload ({sum(c.length for c in sample)} columns), "
                          f"generate_msa + call_variants + keep_only_robust_variants, one chunk per thread",
            }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
