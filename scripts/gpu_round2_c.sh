#!/bin/bash
# ncu --set full of the two kernels that dominate the step (one launch each), source-level counters included
T=${1:-r02c}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pileup_kernel|robust_filter_kernel" -c 3 -o gpurun_out/${T}_prof -f python bench.py --steps 1 --warmup 1 --no-stages --wall-chunks -1 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/${T}_ncu.log
ls -la gpurun_out/${T}_prof.ncu-rep
