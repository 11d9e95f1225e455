#!/bin/bash
# filter defaults (24 codes, 6 CTAs per SM): parity tests; e2e lanes sweep
T=${1:-r02ae}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_callvariants.py -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
for l in 3 4 5; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-stages --wall-chunks -1 --e2e-lanes $l > gpurun_out/${T}_bench_lanes$l.json 2> gpurun_out/${T}_bench_lanes$l.err; echo "lanes $l rc=$?"
python scripts/show_bench.py gpurun_out/${T}_bench_lanes$l.json 2>&1 | head -3
done
