#!/bin/bash
# e2e lanes: 5 against 4, alternating
T=${1:-r02al}
mkdir -p gpurun_out
i=0
for l in 5 4 5 4; do
i=$((i+1))
timeout 600 python bench.py --steps 10 --warmup 3 --no-stages --wall-chunks -1 --e2e-lanes $l > gpurun_out/${T}_bench_${i}_lanes$l.json 2> gpurun_out/${T}_bench_${i}_lanes$l.err; echo "lanes $l rc=$?"
python scripts/show_bench.py gpurun_out/${T}_bench_${i}_lanes$l.json 2>&1 | head -1
done
