#!/bin/bash
T=${1:-r02q}
mkdir -p gpurun_out
HSGPU_TIMING=1 HS_E2E_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-stages --wall-chunks -1 --e2e-lanes 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
grep "lane \|partitions_set" gpurun_out/${T}_bench.err | tail -50
python scripts/show_bench.py gpurun_out/${T}_bench.json 2>&1 | head -1
