#!/bin/bash
# 1 GPU: full parity suite + bench after the row-pattern generalisation
TAG=${1:-r11}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench1.log 2>&1; grep '^{' $OUT/bench1.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['config']['iterations'], d['config']['final_rel_res'], d['config']['upload_s'])
for e in d['roofline_levels']: print(e['kernel'], round(e['ms_per_launch'],4), round(e['frac'],3))
"
HB200_TIMERS=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/t1.log 2>&1; grep -A12 "hb200 timers" $OUT/t1.log | tail -13
