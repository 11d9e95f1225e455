#!/bin/bash
# e2e lanes sweep on 1 GPU, then full GPU test suite
T=${1:-r02n}
mkdir -p gpurun_out
for l in 2 3 4; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-stages --wall-chunks -1 --e2e-lanes $l > gpurun_out/${T}_bench_lanes$l.json 2> gpurun_out/${T}_bench_lanes$l.err; echo "lanes $l rc=$?"
  python scripts/show_bench.py gpurun_out/${T}_bench_lanes$l.json 2>&1 | head -1
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
