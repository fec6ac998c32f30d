#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests -m gpu -x -q -k "fused or als" 2>&1 | tail -3
for v in "" tc100s6 tc116s5 tc172s3 tc132s3; do
  if [ -z "$v" ]; then unset DFM_BENCH_LIB; else export DFM_BENCH_LIB=/root/repo/tools/variants/libdfm_$v.so; fi
  for P in 1184; do
  timeout 120 python bench.py --no-cpu --steps 3 --warmup 3 --panels $P 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant [$v] panels $P value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'als', round(d['als']['value']))"
  done
done
unset DFM_BENCH_LIB
DFM_FUSED_PHASES=1 timeout 150 python bench.py --no-cpu --steps 2 --warmup 3 --panels 1184 2>&1 >/dev/null | head -16
