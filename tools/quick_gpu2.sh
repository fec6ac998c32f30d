#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --no-cpu --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e'].get('ms_per_step',0),2), 'launches', d['gpu_launches'], 'als', round(d['als']['value']))"
DFM_FUSED_PHASES=1 timeout 150 python bench.py --no-cpu --steps 2 --warmup 3 --panels 1184 2>&1 >/dev/null | head -17
