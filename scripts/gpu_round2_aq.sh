#!/bin/bash
# the reference's main() on libhsgpu through integration/glue_call_variants.cpp
T=${1:-r02aq}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_callvariants.py tests/test_gpu_sepreads.py -m gpu -q -k "reference_main_on_libhsgpu" > gpurun_out/${T}_glue_test.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/${T}_glue_test.log
