#!/bin/bash
# round 2, second 1-GPU session: the compact-stencil kernel (spmv_box) on hardware — parity, timing
# against the generic row-pattern kernel, z-run sweep, ncu --set full of one launch, launch list.
# usage: gpurun --timeout 1500 -- 'bash scripts/r2_gpu2.sh r2b'
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
echo "#### pytest -m gpu"
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
one() {  # label, env..., then bench args after --
  local label=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 400 python bench.py "$@" 2>&1 | grep '^{' | tail -1 > $OUT/$label.json
  python - <<P
import json
d=json.load(open("$OUT/$label.json"))
if d['metric'].startswith('parcsr'):
    k=d['config']['kernel_kinds']; print("$label", k['stored']['kernel'][:40], 'ms', round(k['stored']['ms'],4), 'GB/s', round(k['stored']['achieved_gbs']), 'csr', round(k['csr']['ms'],4), 'err', d['config']['parity_vs_reference_max_rel_err'])
else:
    print("$label", round(d['value'],1), 'MDOF/s', round(d['ms_per_step'],2), 'ms its', d['config']['iterations'], d['config']['final_rel_res'], 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
    for e in d['roofline']['levels']: print('    ', e['kernel'][:70], round(e['ms_per_launch'],4), round(e['frac'],3))
P
}
echo "#### spmv-only, box vs generic pattern kernel"
for n in 128 192 256 384; do
  one spmv_box_$n X=1 -- --spmv-only --n $n --steps 2 --warmup 2 --no-cpu-baseline
  one spmv_pat_$n HB200_NO_BOX=1 -- --spmv-only --n $n --steps 2 --warmup 2 --no-cpu-baseline
done
one spmv_box_lap7_256 X=1 -- --spmv-only --problem laplacian --n 256 --steps 2 --warmup 2 --no-cpu-baseline
one spmv_pat_lap7_256 HB200_NO_BOX=1 -- --spmv-only --problem laplacian --n 256 --steps 2 --warmup 2 --no-cpu-baseline
echo "#### z-run sweep (256^3)"
for z in 6 12 24 48 96 258; do one spmv_box_z$z HB200_BOX_ZRUN=$z -- --spmv-only --n 256 --steps 2 --warmup 2 --no-cpu-baseline; done
echo "#### solve"
one bench_box X=1 -- --steps 10 --warmup 3 --no-e2e-ij
one bench_nobox HB200_NO_BOX=1 -- --steps 10 --warmup 3 --no-cpu-baseline --no-e2e-ij
one bench_lap7 X=1 -- --problem laplacian --steps 10 --warmup 3 --no-cpu-baseline --no-e2e-ij
echo "#### ncu --set full: one launch of the box kernel (SpMV and fused l1-Jacobi) and of the generic kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_box -s 4 -c 2 -o $OUT/ncu_box \
   python bench.py --spmv-only --n 256 --steps 1 --warmup 1 --nmv 4 --no-cpu-baseline > $OUT/ncu_box.log 2>&1
HB200_NO_BOX=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_pat -s 4 -c 1 -o $OUT/ncu_pat \
   python bench.py --spmv-only --n 256 --steps 1 --warmup 1 --nmv 4 --no-cpu-baseline > $OUT/ncu_pat.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_box -s 60 -c 6 -o $OUT/ncu_box_solve \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e-ij --no-graph > $OUT/ncu_box_solve.log 2>&1
echo "#### ncu launch list of one solve (graph off so that every kernel is listed)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e-ij > $OUT/launches.log 2>&1
ls -la $OUT | head -40
