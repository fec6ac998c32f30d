#!/bin/bash
# round evidence without the test tier: default bench line, ncu launch list, one ncu --set full capture of the EM kernel
cd /root/repo
tag=${1:-v6}
timeout 200 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 200 gpurun_out/bench_$tag.json; echo
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_l_$tag.log 2>&1; tail -1 gpurun_out/ncu_l_$tag.log | cut -c1-100
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_em_fused2 -c 1 -f -o gpurun_out/prof_em_$tag python bench.py --panels 592 --em-iters 20 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_f_$tag.log 2>&1; tail -1 gpurun_out/ncu_f_$tag.log | cut -c1-100
