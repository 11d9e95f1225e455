#!/bin/bash
# robust filter with one lane per partition: parity tests, bench (A/B against the lane-per-read kernel), config 4 at full size
T=${1:-r02v}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_callvariants.py -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-stages --wall-chunks -1 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench.err
python scripts/show_bench.py gpurun_out/${T}_bench.json 2>&1 | head -12
for c in 4 6; do
HSGPU_FILTER_CTAS=$c timeout 600 python bench.py --steps 10 --warmup 3 --no-stages --wall-chunks -1 > gpurun_out/${T}_bench_ctas$c.json 2> gpurun_out/${T}_bench_ctas$c.err; echo "bench ctas=$c rc=$?"
python scripts/show_bench.py gpurun_out/${T}_bench_ctas$c.json 2>&1 | head -4
done
timeout 1200 python scripts/full_config.py --config 4 --mode check > gpurun_out/${T}_full_4.log 2>&1; echo "full 4 rc=$?"; tail -c 600 gpurun_out/${T}_full_4.log
