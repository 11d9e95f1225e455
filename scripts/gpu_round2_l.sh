#!/bin/bash
# realignment: edlib GPU tests, then the realign stage alone under each blocks-per-lane choice
T=${1:-r02l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edlib.py -m gpu -x -q > gpurun_out/${T}_edlib_tests.log 2>&1; echo "edlib pytest rc=$?"; tail -15 gpurun_out/${T}_edlib_tests.log
for b in 0 1 2 3 4; do
  if [ $b = 0 ]; then unset HSGPU_EDLIB_BPL; else export HSGPU_EDLIB_BPL=$b; fi
  timeout 600 python bench.py --only-realign > gpurun_out/${T}_realign_bpl$b.json 2> gpurun_out/${T}_realign_bpl$b.err; echo "bpl $b rc=$?"
  tail -2 gpurun_out/${T}_realign_bpl$b.err
  python - $T $b <<'PY'
import json, sys
d=json.loads(open('gpurun_out/%s_realign_bpl%s.json'%(sys.argv[1],sys.argv[2])).read().strip().splitlines()[-1])
for k,v in d['realign'].items():
    if isinstance(v,dict): print(k, {a:round(b['ms'],3) for a,b in v.get('kernels').items()}, round(v.get('kernel_gcups')), round(v.get('e2e_gcups')))
PY
done
