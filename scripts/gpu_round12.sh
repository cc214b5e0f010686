#!/bin/bash
# 1 GPU: rows-per-thread choice for the row-pattern kernel, launch list of the solve
TAG=${1:-r12}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
for R in 4 8; do
  HB200_PAT_ROWS=$R timeout 600 python scripts/spmv_sweep.py 256 27pt one 2>&1 | sed -n 2,5p
  HB200_PAT_ROWS=$R timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_R$R.log 2>&1; grep '^{' $OUT/bench_R$R.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('R=$R', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['config']['iterations'], d['config']['final_rel_res'])
for e in d['roofline_levels']: print(e['kernel'], round(e['ms_per_launch'],4), round(e['frac'],3))
"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > $OUT/ncu_bench.log 2>&1
wc -l $OUT/launches.csv
