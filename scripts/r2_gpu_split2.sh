#!/bin/bash
# short N-GPU session: split operation (opt-in) and fused kernels with the two-level flag relay
TAG=${1:-r2z}
NG=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
export HB200_HALO_TIMEOUT_S=20
run() {
  local label=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
     --master-port 29531 bench.py --gpus $NG --no-cpu-baseline --stage-timeout 120 "$@" > $OUT/$label.log 2>&1
  local rc=$?
  grep '^{' $OUT/$label.log | tail -1 > $OUT/$label.json
  python - <<P
import json
try:
    d=json.load(open("$OUT/$label.json"))
    c=d['config']
    print("$label rc=$rc", round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms its', c.get('iterations'), 'ms/it', round(c.get('ms_per_iteration',0),3), c.get('final_rel_res'), 'halo', c.get('halo'), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
    if 'levels' in d.get('roofline',{}):
        print('    ', ' | '.join(f"{e['kernel'][:9]} A_{e['level']} {e['ms_per_launch']:.4f}" for e in d['roofline']['levels']))
except Exception as ex:
    print("$label rc=$rc NO RESULT", ex)
P
  if [ $rc -ne 0 ]; then grep "no progress\|rror\|timed out" $OUT/$label.log | head -5; fi
}
S="--steps 4 --warmup 3"
run base X=1 -- $S --halo peer
run fused8M HB200_FUSED_HALO_MAX=8000000 -- $S --halo peer
run split HB200_SPLIT=1 -- $S --halo peer
run split_fused8M HB200_SPLIT=1 HB200_FUSED_HALO_MAX=8000000 -- $S --halo peer
