#!/bin/bash
# ncu --set full (with source counters) of k_em_fused2<8> for two builds of the library: the in-tree one and a variant
cd /root/repo
for tag in "$@"; do
  if [ "$tag" = "main" ]; then unset DFM_BENCH_LIB; else export DFM_BENCH_LIB=/root/repo/build/variants/libdfm_$tag.so; fi
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_em_fused2 -c 1 -f -o gpurun_out/prof_em_$tag \
    python bench.py --panels 592 --em-iters 20 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_$tag.log 2>&1
  tail -2 gpurun_out/ncu_$tag.log | cut -c1-200
done
