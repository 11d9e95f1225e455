#!/bin/bash
# the new .col writer inside the real executable: HS_call_variants against the reference executable (byte-identical files)
T=${1:-r02as}
mkdir -p gpurun_out
timeout 55 python -m pytest tests/test_gpu_callvariants.py -m gpu -q -x -k "col_vcf_error_rate_identical_to_reference" > gpurun_out/${T}_cv_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_cv_tests.log
