#!/bin/bash
# realignment work: edlib GPU tests, then the bench with stages (realign timings); optional BPL override in $2
T=${1:-r02k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edlib.py -m gpu -x -q > gpurun_out/${T}_edlib_tests.log 2>&1; echo "edlib pytest rc=$?"; tail -15 gpurun_out/${T}_edlib_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${T}_bench.err
python scripts/show_bench.py gpurun_out/${T}_bench.json 2>&1 | head -3
python - $T <<'PY'
import json, sys
d=json.loads(open('gpurun_out/'+sys.argv[1]+'_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('realign_gcups')))
for k,v in d.get('stages',{}).get('realign',{}).items():
    if isinstance(v,dict): print(k, v.get('kernels'), v.get('kernel_gcups'), v.get('e2e_gcups'))
PY
