#!/bin/bash
# (queued while the pod was busy) -> the final validation script
exec bash scripts/gpu_round2_final.sh r02zz "3 5"
