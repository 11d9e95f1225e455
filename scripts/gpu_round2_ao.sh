#!/bin/bash
# band split (HSGPU_EDLIB_BAND=2: reverse sweep banded in its own launch, path sweep + walk on all lanes of the pair):
# edlib tests, realign stage against the default. Measured slower (profiles/r02ao_*); the variant's code
# was not kept, so at HEAD the value 2 behaves like 1.
T=${1:-r02ao}
mkdir -p gpurun_out
HSGPU_EDLIB_BAND=2 timeout 900 python -m pytest tests/test_gpu_edlib.py -m gpu -x -q > gpurun_out/${T}_edlib_band2_tests.log 2>&1; echo "edlib (band=2) pytest rc=$?"; tail -3 gpurun_out/${T}_edlib_band2_tests.log
for b in 2 0; do
  HSGPU_EDLIB_BAND=$b timeout 600 python bench.py --only-realign > gpurun_out/${T}_realign_band$b.json 2> gpurun_out/${T}_realign_band$b.err; echo "realign band=$b rc=$?"; tail -1 gpurun_out/${T}_realign_band$b.err
  python - $T $b <<'PY'
import json, sys
d=json.loads(open('gpurun_out/%s_realign_band%s.json'%(sys.argv[1],sys.argv[2])).read().strip().splitlines()[-1])
for k,v in d['realign'].items():
    if isinstance(v,dict): print(k, {a:round(b['ms'],3) for a,b in v.get('kernels').items()}, round(v.get('kernel_gcups')), round(v.get('e2e_gcups')))
PY
done
