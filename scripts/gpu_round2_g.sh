#!/bin/bash
T=${1:-r02g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/${T}_gpu_tests.log
tail -15 gpurun_out/${T}_gpu_tests.log
if [ $rc -ne 0 ]; then exit 1; fi
for occ in 4 3; do
HSGPU_FILTER_OCC=$occ timeout 900 python bench.py --steps 10 --warmup 3 --no-stages > gpurun_out/${T}_bench_occ$occ.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/${T}_bench_occ$occ.json 2>&1 | head -8
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"robust_filter_kernel" -c 1 -o gpurun_out/${T}_prof -f python bench.py --steps 1 --warmup 1 --no-stages --wall-chunks -1 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
timeout 900 python scripts/full_config.py --config 4 --mode check > gpurun_out/${T}_full_4.log 2>&1; echo "full 4 rc=$?"; tail -c 700 gpurun_out/${T}_full_4.log
