#!/bin/bash
T=${1:-r02d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/${T}_gpu_tests.log
tail -15 gpurun_out/${T}_gpu_tests.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-stages > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/${T}_bench.err
python scripts/show_bench.py gpurun_out/${T}_bench.json 2>&1 | head -40
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"robust_filter_kernel" -c 1 -o gpurun_out/${T}_prof -f python bench.py --steps 1 --warmup 1 --no-stages --wall-chunks -1 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
true
