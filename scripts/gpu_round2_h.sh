#!/bin/bash
# every BASELINE config at full size: the two executables against the reference's hashes (config 3 here; 1, 2, 4, 5
# in earlier calls) and one bench line per config
T=${1:-r02h}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_callvariants.py -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
timeout 1800 python scripts/full_config.py --config 3 --mode check > gpurun_out/${T}_full_3.log 2>&1; echo "full 3 rc=$?"; tail -c 900 gpurun_out/${T}_full_3.log
for k in 1 4 5 3; do
  timeout 1500 python bench.py --config $k --steps 5 --warmup 3 --no-stages --wall-chunks -1 > gpurun_out/${T}_bench_config$k.json 2> gpurun_out/${T}_bench_config$k.err; echo "bench config $k rc=$?"
  python scripts/show_bench.py gpurun_out/${T}_bench_config$k.json 2>&1 | head -6
done
