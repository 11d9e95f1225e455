#!/bin/bash
# realign after the sweep micro-optimisations (tests + stage), e2e trace with 3 lanes
T=${1:-r02o}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edlib.py -m gpu -x -q > gpurun_out/${T}_edlib_tests.log 2>&1; echo "edlib pytest rc=$?"; tail -5 gpurun_out/${T}_edlib_tests.log
timeout 600 python bench.py --only-realign > gpurun_out/${T}_realign.json 2> gpurun_out/${T}_realign.err; echo "realign rc=$?"
tail -2 gpurun_out/${T}_realign.err
python - $T <<'PY'
import json, sys
d=json.loads(open('gpurun_out/%s_realign.json'%(sys.argv[1])).read().strip().splitlines()[-1])
for k,v in d['realign'].items():
    if isinstance(v,dict): print(k, {a:round(b['ms'],3) for a,b in v.get('kernels').items()}, round(v.get('kernel_gcups')), round(v.get('e2e_gcups')))
PY
HS_E2E_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-stages --wall-chunks -1 --e2e-lanes 3 > gpurun_out/${T}_bench_trace.json 2> gpurun_out/${T}_bench_trace.err; echo "trace rc=$?"
grep "lane" gpurun_out/${T}_bench_trace.err | tail -8
python scripts/show_bench.py gpurun_out/${T}_bench_trace.json 2>&1 | head -1
