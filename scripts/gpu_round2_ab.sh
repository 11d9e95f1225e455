#!/bin/bash
# 4 ranks: host waits spinning (default) against sleeping (HSGPU_WAIT=block)
T=${1:-r02ab}
N=${2:-4}
mkdir -p gpurun_out
for w in spin block; do
export HSGPU_WAIT=$w
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-stages --wall-chunks -1 > gpurun_out/${T}_bench_${N}gpu_$w.json 2> gpurun_out/${T}_bench_${N}gpu_$w.err; echo "$w rc=$?"
python scripts/show_bench.py gpurun_out/${T}_bench_${N}gpu_$w.json 2>&1 | head -1
done
