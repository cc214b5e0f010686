#!/bin/bash
# the 8-GPU session (charged 8x: every step under a short timeout): 8-rank parity over the peer halo, then the
# bench lines of the north-star configs at N = 8
TAG=${1:-r2n8}
NG=8
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
export HB200_HALO_TIMEOUT_S=20
echo "#### 8-rank parity, peer halo (2 x 2 x 2 bricks, 7 neighbours)"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "8ranks_peer_halo and not larger" 2>&1 | tail -3
run() {
  local label=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
     --master-port 29531 bench.py --gpus $NG --no-cpu-baseline --stage-timeout 120 "$@" > $OUT/$label.log 2>&1
  local rc=$?
  grep '^{' $OUT/$label.log | tail -1 > $OUT/$label.json
  python - <<P
import json
try:
    d=json.load(open("$OUT/$label.json"))
    c=d['config']
    print("$label rc=$rc", round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms its', c.get('iterations'), 'ms/it', round(c.get('ms_per_iteration',0),3), c.get('final_rel_res'), 'halo', c.get('halo'), 'graph', c.get('cuda_graph_vcycle'), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'upload', round(c.get('upload_s',0),2), 'setup', round(c.get('setup_s_reference_cpu',0),1))
    if 'levels' in d.get('roofline',{}):
        print('    ', ' | '.join(f"{e['kernel'][:9]} A_{e['level']} {e['ms_per_launch']:.4f}" for e in d['roofline']['levels']))
except Exception as ex:
    print("$label rc=$rc NO RESULT", ex)
P
  if [ $rc -ne 0 ]; then grep "no progress\|rror\|timed out" $OUT/$label.log | head -5; fi
}
S="--steps 3 --warmup 2"
run w27_peer_graph X=1 -- $S
run w27_nccl X=1 -- $S --halo nccl
run wlap7_peer_graph X=1 -- $S --problem laplacian
run vdc_gmres_256_strong X=1 -- $S --problem vardifconv --solver gmres --size 256 --global-size
run spmv_256 X=1 -- --spmv-only --size 256 --steps 2 --warmup 2
