#!/bin/bash
# N GPUs: multi-rank parity, bench (auto halo), timers attribution
TAG=${1:-r15}
NG=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
if [ "$NG" = "2" ]; then
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -3
fi
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NG --steps 3 --warmup 3 --no-cpu-baseline > $OUT/$name.log 2>&1
  echo "== $name"; grep -A12 "hb200 timers rank 0" $OUT/$name.log | tail -13; grep '^{' $OUT/$name.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e']['value'], d['config']['iterations'], d['config']['final_rel_res'], d['config']['halo'], d['config']['cuda_graph_vcycle'])
for e in d["roofline_levels"]: print(e["kernel"][:110], round(e['ms_per_launch'],4), round(e['frac'],3))
"
}
run bench HB200_X=0
run timers HB200_TIMERS=1 HB200_BENCH_LEVELS=1
grep "\[levels\]" $OUT/timers.log | head -8
