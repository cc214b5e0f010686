#!/bin/bash
# round 2, first 1-GPU session: everything that ships, on hardware (tests incl. the forced run-time
# switches and the shim update case), then the bench lines of the N = 1 configs.
# usage: gpurun --timeout 1500 -- 'bash scripts/r2_gpu1.sh r2a'
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $OUT/gpu.txt
free -g | head -2 >> $OUT/gpu.txt; nproc >> $OUT/gpu.txt
echo "#### pytest -m gpu"
timeout 1200 python -m pytest tests -x -q -m gpu -rs 2>&1 | tail -45 > $OUT/pytest.log; tail -4 $OUT/pytest.log
echo "#### smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "#### bench default"
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_default.log 2>&1; grep '^{' $OUT/bench_default.log | tail -1 > $OUT/bench_default.json
python - <<P
import json
d=json.load(open("$OUT/bench_default.json"))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'its', d['config']['iterations'], d['config']['final_rel_res'], 'upload', d['config']['upload_s'])
print('ij_dropin', d['e2e']['ij_dropin'])
for e in d['roofline']['levels']: print(e['kernel'][:90], round(e['ms_per_launch'],4), round(e['frac'],3))
print('csr', d['roofline']['csr']['ms_per_launch'], d['roofline']['csr']['frac']); print(d['roofline']['whole_iteration']); print(d['cpu_baseline'])
P
echo "#### bench fused dots"
HB200_FUSED_DOTS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e-ij 2>&1 | grep '^{' | tail -1 > $OUT/bench_fused_dots.json
python -c "
import json; d=json.load(open('$OUT/bench_fused_dots.json')); print('fused dots', d['value'], d['ms_per_step'], d['config']['iterations'], d['config']['final_rel_res'], d['gpu_launches'])"
echo "#### bench all-CSR"
timeout 400 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --format csr 2>&1 | grep '^{' | tail -1 > $OUT/bench_csr.json
python -c "
import json; d=json.load(open('$OUT/bench_csr.json')); print('all-CSR', d['value'], d['ms_per_step'], d['config']['iterations'], d['config']['final_rel_res']); print(d['roofline']['whole_iteration'])
for e in d['roofline']['levels']: print(e['kernel'][:90], round(e['ms_per_launch'],4), round(e['frac'],3))"
echo "#### spmv-only sweep"
for n in 128 192 256 384; do
  timeout 400 python bench.py --spmv-only --n $n --steps 3 --warmup 2 2>&1 | grep '^{' | tail -1 > $OUT/spmv_$n.json
  python -c "
import json; d=json.load(open('$OUT/spmv_$n.json')); k=d['config']['kernel_kinds']
print($n, 'value', round(d['value']), 'GB/s  ms/spmv', round(d['config']['ms_per_spmv'],4), 'csr kernel', round(k['csr']['ms'],4), 'ms', round(k['csr']['achieved_gbs']), 'GB/s frac', round(d['roofline']['frac'],3), 'stored', round(k['stored']['ms'],4), 'err', d['config']['parity_vs_reference_max_rel_err'], 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],1))"
done
echo "#### 7pt laplacian + vardifconv gmres, N=1"
timeout 400 python bench.py --problem laplacian --steps 5 --warmup 2 --no-e2e-ij 2>&1 | grep '^{' | tail -1 > $OUT/bench_lap7.json
timeout 400 python bench.py --problem vardifconv --solver gmres --steps 5 --warmup 2 --no-e2e-ij 2>&1 | grep '^{' | tail -1 > $OUT/bench_vdc_gmres.json
python -c "
import json
for f in ('bench_lap7','bench_vdc_gmres'):
    d=json.load(open('$OUT/'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['config']['iterations'], d['config']['final_rel_res'], d['config']['parity_vs_reference'], d['cpu_baseline'] and d['cpu_baseline']['value'])
    for e in d['roofline']['levels']: print('  ', e['kernel'][:90], round(e['ms_per_launch'],4), round(e['frac'],3))"
