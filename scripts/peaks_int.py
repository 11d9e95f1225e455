#!/usr/bin/env python
"""Measures the two roofline denominators MEASURED_PEAKS.json does not have:

  * INT32 / logic-pipe throughput: chains of lop3 / add / mad.lo / funnel shifts (csrc/peaks.cu), in thread
    instructions per second over the whole chip, with the SM clock seen during the run;
  * dense int8 tensor throughput: cuBLASLt through torch._int_mm (8192^3), the library number a hand-written
    tcgen05 kind::i8 kernel is compared with -- and our own pair_umma_kernel on its dense schedule.

Run on the GPU box:  python scripts/peaks_int.py  -> gpurun_out/peaks_int.json (committed as profiles/peaks_int.json).
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sm_clock():
    try:
        out = subprocess.run(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=10).stdout.strip().split(",")
        return float(out[0]), float(out[1])
    except Exception:
        return None, None


def main():
    import torch
    from hairsplitter_b200 import api
    ctx = api.Context(0)
    lib = ctx.lib
    lib.hsgpu_debug_int_peak.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    res = {"when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "gpu": torch.cuda.get_device_name(0),
           "sm_count": torch.cuda.get_device_properties(0).multi_processor_count}
    kinds = {0: "lop3", 1: "iadd", 2: "imad", 3: "lop3+imad", 4: "shf"}
    for k, name in kinds.items():
        ops, ms = C.c_double(), C.c_double()
        ctx.check(lib.hsgpu_debug_int_peak(ctx.h, k, 20000, C.byref(ops), C.byref(ms)), "int_peak")
        clk, clk_max = sm_clock()
        res[f"int32_{name}_tops"] = ops.value / 1e12
        res[f"int32_{name}_ms"] = ms.value
        # thread instructions per SM and clock at the maximum SM clock (128 = one warp instruction per scheduler and clock)
        if clk_max:
            res[f"int32_{name}_per_sm_clk_at_max"] = ops.value / res["sm_count"] / (clk_max * 1e6)
    res["sm_mhz_after"], res["sm_max_mhz"] = sm_clock()
    # (ptxas folds pairs of dependent adds into one IADD3, so the "iadd" figure counts PTX adds, not issued instructions)
    res["int32_peak_tops"] = max(res[f"int32_{n}_tops"] for n in ("lop3", "imad", "lop3+imad", "shf"))
    res["int32_alu_pipe_tops"] = max(res["int32_lop3_tops"], res["int32_shf_tops"])

    # dense int8 GEMM through cuBLASLt
    n = 8192
    a = torch.randint(-8, 8, (n, n), dtype=torch.int8, device="cuda")
    b = torch.randint(-8, 8, (n, n), dtype=torch.int8, device="cuda")
    try:
        for _ in range(3):
            torch._int_mm(a, b)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res["int8_cublaslt_tops"] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
        res["int8_cublaslt_how"] = "torch._int_mm 8192^3 int8 -> int32, best of 10, CUDA events"
    except Exception as e:  # noqa: BLE001
        res["int8_cublaslt_tops"] = None
        res["int8_cublaslt_error"] = str(e)[:200]
    del a, b

    # our own tcgen05 kind::i8 kernel on its dense schedule (every tile pair over every SNP block)
    import numpy as np
    rng = np.random.default_rng(5)
    R, S, d = 4096, 4096, 1024
    idx = np.stack([np.sort(rng.choice(R, d, replace=False)) for _ in range(S)]).astype(np.uint32).reshape(-1)
    code = rng.integers(40, 44, idx.shape[0]).astype(np.uint8)
    off = np.arange(S + 1, dtype=np.int64) * d
    cols = [(R, off, idx, code, np.full(S, 40, np.uint8), np.full(S, 41, np.uint8))]
    P = api.Pairs(ctx, cols, api.PAIRS_DENSE)
    info = P.info()
    for _ in range(2):
        P.compute()
    ctx.sync()
    ctx.profile(True)
    for _ in range(5):
        P.compute()
    ctx.sync()
    prof = ctx.profile_report()
    ctx.profile(False)
    kms = prof["pair_umma_kernel"][1] / prof["pair_umma_kernel"][0]
    macs = info["kblocks"] * 128.0 * 2 * 128 * 256
    res["int8_pair_umma_dense_tops"] = 2 * macs / (kms * 1e-3) / 1e12
    res["int8_pair_umma_dense_how"] = f"pair_umma_kernel, {R} reads x {S} SNPs, HSGPU_PAIRS_DENSE, {kms:.3f} ms"
    P.close()
    vals = [v for v in (res.get("int8_cublaslt_tops"), res["int8_pair_umma_dense_tops"]) if v]
    res["int8_peak_tops"] = max(vals)
    ctx.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "peaks_int.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
