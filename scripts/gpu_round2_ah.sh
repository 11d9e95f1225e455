#!/bin/bash
# 2 ranks after the bench changes (adaptive lanes, per-lane result buffers): weak line + reference arm under torchrun
T=${1:-r02ah}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 10 --warmup 3 --no-stages --wall-chunks -1 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err; echo "weak rc=$?"; tail -2 gpurun_out/${T}_bench_2gpu.err
python scripts/show_bench.py gpurun_out/${T}_bench_2gpu.json 2>&1 | head -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_2gpu.json 2> gpurun_out/${T}_bench_reference_2gpu.err; echo "reference arm rc=$?"; tail -c 300 gpurun_out/${T}_bench_reference_2gpu.json
