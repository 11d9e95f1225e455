#!/bin/bash
T=${1:-r02ai}
mkdir -p gpurun_out
HSGPU_EDLIB_BAND=1 timeout 900 python -m pytest tests/test_gpu_edlib.py -m gpu -x -q > gpurun_out/${T}_edlib_band_tests.log 2>&1; echo "edlib (band) pytest rc=$?"; tail -3 gpurun_out/${T}_edlib_band_tests.log
for b in 1 0; do
  HSGPU_EDLIB_BAND=$b timeout 600 python bench.py --only-realign > gpurun_out/${T}_realign_band$b.json 2> gpurun_out/${T}_realign_band$b.err; echo "realign band=$b rc=$?"
  python - $T $b <<'PY'
import json, sys
d=json.loads(open('gpurun_out/%s_realign_band%s.json'%(sys.argv[1],sys.argv[2])).read().strip().splitlines()[-1])
for k,v in d['realign'].items():
    if isinstance(v,dict): print(k, {a:round(b['ms'],3) for a,b in v.get('kernels').items()}, round(v.get('kernel_gcups')), round(v.get('e2e_gcups')))
PY
done
