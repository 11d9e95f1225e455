#!/bin/bash
T=${1:-r02ag}
mkdir -p gpurun_out
for rep in 1 2 3; do
HS_STEP_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-stages --wall-chunks -1 > gpurun_out/${T}_bench_$rep.json 2> gpurun_out/${T}_bench_$rep.err; echo "rep $rep rc=$?"
grep "step host" gpurun_out/${T}_bench_$rep.err | tail -10
python scripts/show_bench.py gpurun_out/${T}_bench_$rep.json 2>&1 | head -1
done
