#!/bin/bash
# round 2, third 1-GPU session: spmv_box with the cp.async slab ring (timing, racecheck, ncu), and the
# reference's own CUDA backend (baseline/_ref/ij_hypre_cuda: hypre 3.1.0 built with -DHYPRE_ENABLE_CUDA=ON
# for sm_100, 1 rank) on the same box: its solve and its SpMV (own kernel / cuSPARSE) beside hb200's.
# usage: gpurun --timeout 1500 -- 'bash scripts/r2_gpu3.sh r2c'
TAG=${1:-r2c}
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
echo "#### spmv-only, box (cp.async ring) vs generic pattern kernel"
for n in 128 192 256 384; do
  one spmv_box_$n X=1 -- --spmv-only --n $n --steps 2 --warmup 2 --no-cpu-baseline
done
one spmv_box_lap7_256 X=1 -- --spmv-only --problem laplacian --n 256 --steps 2 --warmup 2 --no-cpu-baseline
for z in 12 24 96; do one spmv_box_z$z HB200_BOX_ZRUN=$z -- --spmv-only --n 256 --steps 2 --warmup 2 --no-cpu-baseline; done
echo "#### solve"
one bench_box X=1 -- --steps 10 --warmup 3 --no-e2e-ij
one bench_lap7 X=1 -- --problem laplacian --steps 10 --warmup 3 --no-cpu-baseline --no-e2e-ij
one bench_pmis X=1 -- --steps 10 --warmup 3 --no-e2e-ij --coarsen-type 8
echo "#### racecheck + memcheck of the box kernel (small grid)"
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "bit_exact_fine_level or relax_jacobi" 2>&1 | tail -6 > $OUT/racecheck.log; tail -3 $OUT/racecheck.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "bit_exact_fine_level or relax_jacobi or pcg_amg" 2>&1 | tail -6 > $OUT/memcheck.log; tail -3 $OUT/memcheck.log
echo "#### the reference's CUDA backend on this box (hypre 3.1.0, -DHYPRE_ENABLE_CUDA=ON, sm_100)"
H=baseline/_ref/ij_hypre_cuda
if [ -x $H ]; then
  for v in 0 1; do
    timeout 300 $H -27pt -n 256 256 256 -solver -1 -nmv 1000 -x0rand -exec_device -memory_device -mv_vendor $v > $OUT/hypre_cuda_spmv_v$v.log 2>&1
    echo "hypre-CUDA SpMV 27pt 256^3 x1000, mv_vendor=$v:"; grep -A3 "MatVec Test" $OUT/hypre_cuda_spmv_v$v.log | grep "wall clock time"
  done
  for v in 0 1; do
    timeout 600 $H -27pt -n 256 256 256 -solver 1 -rlx 18 -exec_device -memory_device -mv_vendor $v > $OUT/hypre_cuda_solve_v$v.log 2>&1
    echo "hypre-CUDA AMG-PCG 27pt 256^3 -rlx 18 (device setup, its own defaults), mv_vendor=$v:"
    grep -A2 "PCG Setup\|PCG Solve" $OUT/hypre_cuda_solve_v$v.log | grep "wall clock"; grep "Iterations\|Final Relative\|Complexity\|num_levels\|Number of levels" $OUT/hypre_cuda_solve_v$v.log | head -6
  done
  timeout 600 $H -laplacian -n 256 256 256 -solver 1 -rlx 18 -exec_device -memory_device > $OUT/hypre_cuda_solve_lap7.log 2>&1
  echo "hypre-CUDA AMG-PCG 7pt 256^3:"; grep -A2 "PCG Setup\|PCG Solve" $OUT/hypre_cuda_solve_lap7.log | grep "wall clock"; grep "Iterations\|Final Relative" $OUT/hypre_cuda_solve_lap7.log
  echo "launch list of the hypre-CUDA solve (kernel-time share of its SpMV / relax kernels)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/hypre_cuda_launches.csv \
     $H -27pt -n 128 128 128 -solver 1 -rlx 18 -exec_device -memory_device > $OUT/hypre_cuda_ncu.log 2>&1
  python - <<P
import csv, collections
t=collections.Counter(); n=collections.Counter()
try:
    rows=[r for r in csv.reader(open("$OUT/hypre_cuda_launches.csv")) if len(r)>10]
    hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
    for r in rows[1:]:
        try: t[r[ki][:70]]+=float(r[vi].replace(',','')); n[r[ki][:70]]+=1
        except: pass
    tot=sum(t.values())
    for k,v in t.most_common(12): print(f"{v/tot*100:5.1f}% {n[k]:5d}x {k}")
except Exception as e: print("no launch list", e)
P
else
  echo "baseline/_ref/ij_hypre_cuda missing"
fi
echo "#### ncu --set full: box kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_box -s 4 -c 1 -o $OUT/ncu_box \
   python bench.py --spmv-only --n 256 --steps 1 --warmup 1 --nmv 4 --no-cpu-baseline > $OUT/ncu_box.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_box -s 60 -c 4 -o $OUT/ncu_box_solve \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e-ij --no-graph > $OUT/ncu_box_solve.log 2>&1
ls $OUT | wc -l
