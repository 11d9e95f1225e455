"""One 300 kb chunk of configs[1] through pileup + ranking + hsgpu_robust_filter (for ncu captures of the
contingency kernels). Partitions come from the reference's keep_only_robust_variants when oracle/_ref is there."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hairsplitter_b200 import api, synth
from oracle import pyoracle

chunks, info = synth.make_config(2, scale=1.0, seed=2, n_chunks=1)
cb = chunks[0]
ctx = api.Context(0)
pu = api.Pileup(ctx, api.PackedBatch([cb]))
pu.build()
pu.column_rank()
pos, _ = pu.suspects(0)
filt = None
if pyoracle.ref_available():
    R = pyoracle.RefCV(cb)
    R.call_variants()
    parts, filt, merged = R.robust()
else:
    ends_ = pu.read_ends()
    parts = []
    for w0 in range(0, cb.length, 3000):
        idx = np.nonzero((cb.start < w0 + 3000) & (ends_ > w0))[0].astype(np.int32)
        st = np.where(cb.strain[idx] == 0, 1, -1).astype(np.int16)
        parts.append(dict(read_idx=idx, state=st, more=np.full(idx.size, 3, np.int32), less=np.zeros(idx.size, np.int32)))
for rep in range(3):
    ctx.profile(True)
    t0 = time.perf_counter()
    kept = pu.robust_filter(0, parts, pos)
    dt = time.perf_counter() - t0
    prof = ctx.profile_report()
    ctx.profile(False)
    print(f"robust_filter: {len(parts)} partitions, {pos.size} suspects -> {kept.size} kept, call {dt * 1e3:.3f} ms, kernels "
          + ", ".join(f"{k} {v[1]:.3f} ms" for k, v in prof.items()), flush=True)
if filt is not None:
    assert np.array_equal(kept, filt["pos"]), "snps_out differs from the reference"
    print("identical to the reference's snps_out")
