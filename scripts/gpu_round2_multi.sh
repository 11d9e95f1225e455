#!/bin/bash
# multi-GPU: the two-GPU executables' tests, the weak-scaling bench line and the strong-scaling line (one config-3
# workload dealt to the ranks by LPT)
N=${1:-2}
T=${2:-r02s}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "two_gpus" > gpurun_out/${T}_two_gpu_tests.log 2>&1; echo "two-gpu pytest rc=$?"; tail -3 gpurun_out/${T}_two_gpu_tests.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-stages --wall-chunks -1 > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err; echo "weak rc=$?"
python scripts/show_bench.py gpurun_out/${T}_bench_${N}gpu.json 2>&1 | head -1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --strong --config 3 --steps 5 --warmup 3 --no-stages --wall-chunks -1 > gpurun_out/${T}_bench_strong_${N}gpu.json 2> gpurun_out/${T}_bench_strong_${N}gpu.err; echo "strong rc=$?"
python scripts/show_bench.py gpurun_out/${T}_bench_strong_${N}gpu.json 2>&1 | head -1
python - $T $N <<'PY'
import json, sys
for kind in ("bench", "bench_strong"):
    try:
        d=json.loads(open('gpurun_out/%s_%s_%sgpu.json'%(sys.argv[1],kind,sys.argv[2])).read().strip().splitlines()[-1])
        print(kind, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'scaling', d['scaling'], 'realign', {k: round(v['kernel_gcups']) for k,v in d.get('realign_gcups',{}).items()}, d.get('strong'))
    except Exception as e:
        print(kind, 'unreadable', e)
PY
