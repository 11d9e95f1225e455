#!/bin/bash
# A/B of robust_filter_lanes_kernel builds (distinct-code capacity, minimum CTAs per SM) on configs 2 and 3
T=${1:-r02ad}
mkdir -p gpurun_out
cp hairsplitter_b200/libhsgpu.so /tmp/libhsgpu_default.so
run() {  # name, lib, ctas
  cp $2 hairsplitter_b200/libhsgpu.so
  for cfg in 2 3; do
    HSGPU_FILTER_CTAS=$3 timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 --no-stages --wall-chunks -1 > gpurun_out/${T}_$1_c$cfg.json 2> gpurun_out/${T}_$1_c$cfg.err
    echo "$1 config $cfg rc=$?"; python scripts/show_bench.py gpurun_out/${T}_$1_c$cfg.json 2>&1 | head -3 | tail -2
  done
}
run c24b6 build_variants/libhsgpu_c24_b6.so 6
run c16b6 build_variants/libhsgpu_c16_b6.so 6
run c32b4 build_variants/libhsgpu_c32_b4.so 4
cp /tmp/libhsgpu_default.so hairsplitter_b200/libhsgpu.so
