#!/bin/bash
# full-size checks of configs 2, 4 and 1 on the final kernels
T=${1:-r02ak}
mkdir -p gpurun_out
for k in 2 4 1; do
  timeout 1200 python scripts/full_config.py --config $k --mode check > gpurun_out/${T}_full_$k.log 2>&1; echo "full $k rc=$?"; tail -c 420 gpurun_out/${T}_full_$k.log; echo
  cp gpurun_out/full_config_$k.json gpurun_out/${T}_full_config_$k.json 2>/dev/null
done
