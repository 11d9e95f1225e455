#!/bin/bash
# GPU call: parity tests first (stop when red), then smoke, bench + launch list, full-size configs
T=${1:-r02b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/${T}_gpu_tests.log
tail -15 gpurun_out/${T}_gpu_tests.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/${T}_bench.err
python scripts/show_bench.py gpurun_out/${T}_bench.json 2>&1 | head -40
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-stages --wall-chunks -1 > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu rc=$?"
shift
for k in "$@"; do
  timeout 1500 python scripts/full_config.py --config $k --mode check > gpurun_out/${T}_full_$k.log 2>&1; echo "full $k rc=$?"; tail -c 900 gpurun_out/${T}_full_$k.log
done
