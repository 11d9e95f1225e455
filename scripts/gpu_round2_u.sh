#!/bin/bash
# pileup walk with the shifts / increments on the FMA pipe: parity tests, bench, then the ncu captures of the step
T=${1:-r02u}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_callvariants.py -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-stages --wall-chunks -1 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/${T}_bench.json 2>&1 | head -8
bash scripts/gpu_round2_t.sh $T
