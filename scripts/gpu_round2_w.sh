#!/bin/bash
T=${1:-r02w}
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"^robust_filter_lanes_kernel" -s 1 -c 1 -o gpurun_out/${T}_filter --force-overwrite python bench.py --steps 2 --warmup 1 --no-stages --wall-chunks -1 --e2e-lanes 1 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/${T}_ncu.log
