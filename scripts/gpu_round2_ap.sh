#!/bin/bash
# sidecar + temporaries guard: the whole GPU test-suite, then both executables on config 2 at full size (the second one
# reads the first one's sidecar; hashes against the reference's)
T=${1:-r02ap}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_gpu_tests.log
timeout 200 python scripts/full_config.py --config 2 --mode check > gpurun_out/${T}_full_2.log 2>&1; echo "full_config 2 rc=$?"
cp gpurun_out/full_config_2.json gpurun_out/${T}_full_config_2.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/full_config_2.json'))
print({k:d[k] for k in ('col_identical_to_reference','gro_identical_to_pinned_reference','call_variants_s','separate_reads_s')})
for l in d['timing_separate_reads'][:3]+[x for x in d['timing_call_variants'] if 'write_outputs' in x]: print(l)
PY
