#!/bin/bash
T=${1:-r02z}
mkdir -p gpurun_out
for sp in 0 1; do
HSGPU_SHARED_POOL=$sp HS_STEP_TRACE=1 timeout 900 python bench.py --config 3 --steps 5 --warmup 3 --no-stages --wall-chunks -1 > gpurun_out/${T}_bench_c3_sp$sp.json 2> gpurun_out/${T}_bench_c3_sp$sp.err; echo "config 3 shared_pool=$sp rc=$?"; grep "step host" gpurun_out/${T}_bench_c3_sp$sp.err | tail -5
python scripts/show_bench.py gpurun_out/${T}_bench_c3_sp$sp.json 2>&1 | head -3
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-stages --wall-chunks -1 > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "config 2 rc=$?"
python scripts/show_bench.py gpurun_out/${T}_bench_c2.json 2>&1 | head -3
