#!/bin/bash
# first GPU call of round 2: sanity of the round-1 state + the new measurement scripts
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt 2>&1
nproc > gpurun_out/r02a_nproc.txt; free -g >> gpurun_out/r02a_nproc.txt; df -h /tmp >> gpurun_out/r02a_nproc.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02a_smoke.log
timeout 300 python scripts/peaks_int.py > gpurun_out/r02a_peaks.log 2>&1; echo "peaks rc=$?" >> gpurun_out/r02a_peaks.log
for k in 1 2; do
  timeout 600 python scripts/full_config.py --config $k --mode check > gpurun_out/r02a_full_$k.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_full_$k.log
done
timeout 600 python scripts/full_config.py --config 5 --scale 0.01 --mode check > gpurun_out/r02a_full_5s.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_full_5s.log
tail -3 gpurun_out/r02a_*.log
