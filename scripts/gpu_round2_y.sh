#!/bin/bash
T=${1:-r02y}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_callvariants.py -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
for cfg in 2 3; do
timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 --no-stages --wall-chunks -1 > gpurun_out/${T}_bench_c${cfg}.json 2> gpurun_out/${T}_bench_c${cfg}.err; echo "config $cfg rc=$?"; tail -1 gpurun_out/${T}_bench_c${cfg}.err
python scripts/show_bench.py gpurun_out/${T}_bench_c${cfg}.json 2>&1 | head -5
done
timeout 1200 python scripts/full_config.py --config 4 --mode check > gpurun_out/${T}_full_4.log 2>&1; echo "full 4 rc=$?"; tail -c 400 gpurun_out/${T}_full_4.log
