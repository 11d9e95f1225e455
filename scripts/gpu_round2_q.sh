#!/bin/bash
T=${1:-r02q}
mkdir -p gpurun_out
for sp in 0 1; do
export HSGPU_SHARED_POOL=$sp
HS_E2E_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-stages --wall-chunks -1 --e2e-lanes 3 > gpurun_out/${T}_bench_sp$sp.json 2> gpurun_out/${T}_bench_sp$sp.err; echo "bench shared_pool=$sp rc=$?"
grep "lane " gpurun_out/${T}_bench_sp$sp.err | tail -9
python scripts/show_bench.py gpurun_out/${T}_bench_sp$sp.json 2>&1 | head -1
done
unset HSGPU_SHARED_POOL
timeout 600 python bench.py --steps 10 --warmup 3 --no-stages --wall-chunks -1 --e2e-lanes 2 > gpurun_out/${T}_bench_l2.json 2> gpurun_out/${T}_bench_l2.err; echo "bench lanes 2 rc=$?"
python scripts/show_bench.py gpurun_out/${T}_bench_l2.json 2>&1 | head -1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_callvariants.py tests/test_gpu_sepreads.py -m gpu -x -q 2>&1 | tail -3
