#!/bin/bash
# round-2 evidence: smoke, gpu test tier, bench lines of every config + reference arm, ncu launch list, full ncu captures
cd /root/repo
tag=${1:-r02b}
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; tail -1 gpurun_out/smoke_$tag.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -1 gpurun_out/pytest_$tag.log
timeout 300 python bench.py > gpurun_out/bench_c5_$tag.json 2> gpurun_out/bench_c5_$tag.err; tail -c 200 gpurun_out/bench_c5_$tag.json; echo
for c in c4 c3 c2-single; do
  timeout 300 python bench.py --config $c > gpurun_out/bench_${c}_$tag.json 2> gpurun_out/bench_${c}_$tag.err; tail -c 200 gpurun_out/bench_${c}_$tag.json; echo
done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; tail -c 200 gpurun_out/bench_ref_$tag.json; echo
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_l_$tag.log 2>&1; tail -1 gpurun_out/ncu_l_$tag.log | cut -c1-100
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c3_$tag.csv python bench.py --config c3 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_l3_$tag.log 2>&1; tail -1 gpurun_out/ncu_l3_$tag.log | cut -c1-100
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_em_fused2 -c 1 -f -o gpurun_out/prof_em_$tag python bench.py --panels 592 --em-iters 20 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_f_$tag.log 2>&1; tail -1 gpurun_out/ncu_f_$tag.log | cut -c1-100
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_em_filter_smooth -s 2 -c 1 -f -o gpurun_out/prof_fs_c3_$tag python bench.py --config c3 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_f3_$tag.log 2>&1; tail -1 gpurun_out/ncu_f3_$tag.log | cut -c1-100
python tools/fs_prof.py c3 3 > gpurun_out/fs_prof_c3_$tag.json 2>&1; python tools/fs_prof.py c1 3 > gpurun_out/fs_prof_c1_$tag.json 2>&1
python tools/bench_c1_em.py 1000 5 4 > gpurun_out/c1_em_$tag.json 2>&1; cat gpurun_out/c1_em_$tag.json
