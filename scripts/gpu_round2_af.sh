#!/bin/bash
# e2e with per-lane result buffers and two warm-up steps per lane: lanes sweep, twice
T=${1:-r02af}
mkdir -p gpurun_out
for rep in 1 2; do
for l in 3 4; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-stages --wall-chunks -1 --e2e-lanes $l > gpurun_out/${T}_bench_lanes${l}_$rep.json 2> gpurun_out/${T}_bench_lanes${l}_$rep.err; echo "lanes $l rep $rep rc=$?"; tail -1 gpurun_out/${T}_bench_lanes${l}_$rep.err
python scripts/show_bench.py gpurun_out/${T}_bench_lanes${l}_$rep.json 2>&1 | head -1
done
done
