#!/bin/bash
# realignment: ncu capture of the packed phase B kernel on the chunk shape, then resident-CTA sweep
T=${1:-r02m}
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:edlib_phase_b_packed -s 1 -c 1 -o gpurun_out/${T}_edlib_b --force-overwrite python bench.py --only-realign > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/${T}_ncu.log
for c in 5 6 8; do
  export HSGPU_EDLIB_CTAS=$c
  timeout 600 python bench.py --only-realign > gpurun_out/${T}_realign_ctas$c.json 2> gpurun_out/${T}_realign_ctas$c.err; echo "ctas $c rc=$?"
  tail -2 gpurun_out/${T}_realign_ctas$c.err
  python - $T $c <<'PY'
import json, sys
d=json.loads(open('gpurun_out/%s_realign_ctas%s.json'%(sys.argv[1],sys.argv[2])).read().strip().splitlines()[-1])
for k,v in d['realign'].items():
    if isinstance(v,dict): print(k, {a:round(b['ms'],3) for a,b in v.get('kernels').items()}, round(v.get('kernel_gcups')), round(v.get('e2e_gcups')))
PY
done
