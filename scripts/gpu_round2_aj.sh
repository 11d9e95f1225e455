#!/bin/bash
T=${1:-r02aj}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-stages --wall-chunks -1 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/${T}_bench.err
python scripts/show_bench.py gpurun_out/${T}_bench.json 2>&1 | head -24
