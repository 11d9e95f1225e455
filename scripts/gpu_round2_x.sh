#!/bin/bash
# robust filter: lane-per-partition (second version) against lane-per-read, configs 2 and 3
T=${1:-r02x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_callvariants.py -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
for cfg in 2 3; do
for l in 1 0; do
HSGPU_FILTER_LANES=$l timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 --no-stages --wall-chunks -1 > gpurun_out/${T}_bench_c${cfg}_lanes$l.json 2> gpurun_out/${T}_bench_c${cfg}_lanes$l.err; echo "config $cfg lanes=$l rc=$?"; tail -1 gpurun_out/${T}_bench_c${cfg}_lanes$l.err
python scripts/show_bench.py gpurun_out/${T}_bench_c${cfg}_lanes$l.json 2>&1 | head -5
done
done
