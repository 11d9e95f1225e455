#!/bin/bash
# step clean-up (suspect kernels, filter_active, device-side partition rows): full GPU tests, bench with e2e trace
T=${1:-r02p}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${T}_gpu_tests.log
for l in 2 3; do
HS_E2E_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-stages --wall-chunks -1 --e2e-lanes $l > gpurun_out/${T}_bench_lanes$l.json 2> gpurun_out/${T}_bench_lanes$l.err; echo "bench rc=$?"
grep "lane" gpurun_out/${T}_bench_lanes$l.err | tail -6
python scripts/show_bench.py gpurun_out/${T}_bench_lanes$l.json 2>&1 | head -24
done
