#!/bin/bash
# last run of the round: GPU tests, the bench line with stages, its launch list, smoke
T=${1:-r02zzz}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench.err
python scripts/show_bench.py gpurun_out/${T}_bench.json 2>&1 | head -24
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-stages --wall-chunks -1 --e2e-lanes 1 > gpurun_out/${T}_launches.log 2>&1; echo "launch list rc=$?"
