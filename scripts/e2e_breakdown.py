"""Sweep of the e2e path of bench.py over (host lanes, chunk groups) -- diagnostic, not a bench number."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from hairsplitter_b200 import api, synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
chunks, info = synth.make_config(2, scale=scale, seed=2)
cols = sum(c.length for c in chunks)
ctx = api.Context(0)
for lanes, groups in ((2, 1), (3, 1), (2, 1)):
    r = bench.run_e2e(chunks, 0, ctx, lanes, groups, 10, 3, torch.cuda.synchronize)
    print(f"lanes {r['lanes']} groups {r['groups']:2d}: device {r['device_ms'] / 10:.3f} ms/step, wall {r['wall_ms'] / 10:.3f} ms/step, "
          f"{cols / 2000 / (r['device_ms'] / 10 * 1e-3):.0f} windows/s, h2d {r['h2d_bytes'] / 1e6:.1f} MB", flush=True)
