"""The realignment stage of bench.py alone (20000 read-chunk x window pairs, HW + PATH), for ncu captures of the
edlib kernels."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from hairsplitter_b200 import api

rng = np.random.default_rng(12345)
contig = rng.integers(0, 4, 300000).astype(np.uint8)
qs, ts = bench.make_realign_pairs(rng, contig, 20000)
cells = float(sum(len(q) * len(t) for q, t in zip(qs, ts)))
ctx = api.Context(0)
ctx.edlib_align_batch(qs[:256], ts[:256], k=-1, mode=2, task=2)
for rep in range(2):
    ctx.profile(True)
    ctx.edlib_align_batch(qs, ts, k=-1, mode=2, task=2)
    ctx.sync()
    prof = ctx.profile_report()
    ctx.profile(False)
    ms = sum(v[1] for k, v in prof.items() if k.startswith("edlib_"))
    print(f"realign: {cells / (ms * 1e-3) / 1e9:.0f} GCUPS, " + ", ".join(f"{k} {v[1]:.3f} ms" for k, v in prof.items()), flush=True)
