#!/bin/bash
# 1 GPU: parity of the new variants, per-level sweep incl. 16-bit CSR, bench, ncu of the Jacobi pattern kernel
TAG=${1:-r13}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 python scripts/level_sweep.py 27pt 256 > $OUT/levels_27pt.txt 2>&1; cut -c1-300 $OUT/levels_27pt.txt | head -24
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench1.log 2>&1; grep '^{' $OUT/bench1.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['config']['iterations'], d['config']['final_rel_res'], d['config']['upload_s'])
for e in d['roofline_levels']: print(e['kernel'], round(e['ms_per_launch'],4), round(e['frac'],3))
"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_pat -c 12 -o $OUT/prof_pat python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-graph > $OUT/ncu_pat.log 2>&1
tail -2 $OUT/ncu_pat.log | cut -c1-200
