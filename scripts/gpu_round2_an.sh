#!/bin/bash
# HEAD of the round: the default bench line (with stages) and the reference arm
T=${1:-r02an}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench.err
python scripts/show_bench.py gpurun_out/${T}_bench.json 2>&1 | head -6
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference arm rc=$?"; tail -c 200 gpurun_out/${T}_bench_reference.json
