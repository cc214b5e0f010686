#!/bin/bash
# short 1-GPU session: one kernel iteration (spmv_box) — sizes, solve, one ncu capture
TAG=${1:-r2f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
one() {
  local label=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 400 python bench.py "$@" 2>$OUT/$label.err | grep '^{' | tail -1 > $OUT/$label.json
  python - <<P
import json
try:
    d=json.load(open("$OUT/$label.json"))
    if d['metric'].startswith('parcsr'):
        k=d['config']['kernel_kinds']; print("$label", k['stored']['kernel'][:40], 'ms', round(k['stored']['ms'],4), 'GB/s', round(k['stored']['achieved_gbs']), 'csr', round(k['csr']['ms'],4), 'err', d['config']['parity_vs_reference_max_rel_err'], 'upload', round(d['config']['upload_s'],2))
    else:
        print("$label", round(d['value'],1), 'MDOF/s', round(d['ms_per_step'],2), 'ms its', d['config']['iterations'], d['config']['final_rel_res'], 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'upload', round(d['config']['upload_s'],2))
        for e in d['roofline']['levels']: print('    ', e['kernel'][:70], round(e['ms_per_launch'],4), round(e['frac'],3))
except Exception as ex: print("$label FAILED", ex)
P
}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bit_exact or kernel_variants or relax_jacobi or pcg_amg or vcycle" 2>&1 | tail -2
for n in 128 192 256 384; do one spmv_box_$n X=1 -- --spmv-only --n $n --steps 2 --warmup 2 --no-cpu-baseline; done
one spmv_box_256_nouni HB200_BOX_NO_UNI=1 -- --spmv-only --n 256 --steps 2 --warmup 2 --no-cpu-baseline
one spmv_lap7_256 X=1 -- --spmv-only --problem laplacian --n 256 --steps 2 --warmup 2 --no-cpu-baseline
one spmv_box_256_nobulk HB200_BOX_NO_BULK=1 -- --spmv-only --n 256 --steps 2 --warmup 2 --no-cpu-baseline
for z in 16 32 64; do one spmv_box_z$z HB200_BOX_ZRUN=$z -- --spmv-only --n 256 --steps 2 --warmup 2 --no-cpu-baseline; done
one bench_default X=1 -- --steps 10 --warmup 3 --no-e2e-ij --no-cpu-baseline
one bench_nobox HB200_NO_BOX=1 -- --steps 10 --warmup 3 --no-cpu-baseline --no-e2e-ij
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "bit_exact_fine_level or relax_jacobi" > $OUT/racecheck.log 2>&1; grep -m 12 "hazard\|Hazard\|at hb::\|RACECHECK" $OUT/racecheck.log; tail -2 $OUT/racecheck.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_box -s 60 -c 4 -o $OUT/ncu_box_solve \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e-ij --no-graph > $OUT/ncu_box_solve.log 2>&1
HB200_TRACE=1 timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e-ij 2>&1 | grep "transpose" > $OUT/upload_trace.log; cat $OUT/upload_trace.log
