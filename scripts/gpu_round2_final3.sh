#!/bin/bash
# last run of the round at HEAD: the whole GPU test-suite (with the glued reference executables and the sidecar chain),
# the default bench line with the driver's arguments, smoke(), then -- with what is left of the budget -- one
# ncu --set full capture of the step's three big kernels for roofline.traffic
T=${1:-r02ar}
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_gpu_tests.log
timeout 170 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/${T}_bench.json 2>&1 | head -12
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
timeout 110 ncu --set full --import-source on --clock-control none -k regex:"^(robust_filter_lanes_kernel|pileup_kernel|column_rank_kernel)" -s 3 -c 3 -o gpurun_out/${T}_step --force-overwrite python bench.py --steps 2 --warmup 1 --no-stages --wall-chunks -1 --e2e-lanes 1 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/${T}_ncu.log
