#!/bin/bash
# ncu --set full of the step's three big kernels (one launch each) + the launch list of the same command
T=${1:-r02t}
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"^(robust_filter_kernel|pileup_kernel|column_rank_kernel)" -s 3 -c 3 -o gpurun_out/${T}_step --force-overwrite python bench.py --steps 2 --warmup 1 --no-stages --wall-chunks -1 --e2e-lanes 1 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/${T}_ncu.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-stages --wall-chunks -1 --e2e-lanes 1 > gpurun_out/${T}_launches.log 2>&1; echo "launch list rc=$?"
