#!/bin/bash
# final validation of the round: GPU tests (also with the banded realignment sweeps), realign A/B, the bench line with
# stages, the reference arm, the launch list, and the full-size checks of the configs named in $2 (default "3 5 1 2")
T=${1:-r02zz}
CFGS=${2:-"3 5 1 2"}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
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
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench.err
python scripts/show_bench.py gpurun_out/${T}_bench.json 2>&1 | head -24
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference arm rc=$?"; tail -c 400 gpurun_out/${T}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-stages --wall-chunks -1 --e2e-lanes 1 > gpurun_out/${T}_launches.log 2>&1; echo "launch list rc=$?"
for k in $CFGS; do
  timeout 1800 python scripts/full_config.py --config $k --mode check > gpurun_out/${T}_full_$k.log 2>&1; echo "full $k rc=$?"; tail -c 500 gpurun_out/${T}_full_$k.log; echo
done
