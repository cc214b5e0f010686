#!/bin/bash
# N-GPU session: split operation on structured levels + fused kernels on CSR levels
TAG=${1:-r2y}
NG=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
export HB200_HALO_TIMEOUT_S=20
echo "#### multi-rank parity"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_gpu or two_ranks" 2>&1 | tail -3
run() {
  local label=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
     --master-port 29531 bench.py --gpus $NG --no-cpu-baseline --stage-timeout 150 "$@" > $OUT/$label.log 2>&1
  local rc=$?
  grep '^{' $OUT/$label.log | tail -1 > $OUT/$label.json
  python - <<P
import json
try:
    d=json.load(open("$OUT/$label.json"))
    c=d['config']
    print("$label rc=$rc", round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms its', c.get('iterations'), 'ms/it', round(c.get('ms_per_iteration',0),3), c.get('final_rel_res'), 'halo', c.get('halo'), 'graph', c.get('cuda_graph_vcycle'), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
    if 'levels' in d.get('roofline',{}):
        for e in d['roofline']['levels']: print('    ', e['kernel'][:80], round(e['ms_per_launch'],4), round(e['frac'],3))
except Exception as ex:
    print("$label rc=$rc NO RESULT", ex)
P
  if [ $rc -ne 0 ]; then grep "no progress\|rror\|timed out" $OUT/$label.log | head -5; fi
}
S="--steps 5 --warmup 3"
run split_fused8M HB200_FUSED_HALO_MAX=8000000 -- $S --halo peer
run split_fused_default X=1 -- $S --halo peer
run nosplit_fused8M HB200_NO_SPLIT=1 HB200_FUSED_HALO_MAX=8000000 -- $S --halo peer
run split_nofused HB200_FUSED_HALO=0 -- $S --halo peer
run split_nccl X=1 -- $S --halo nccl
run split_fused8M_lap7 HB200_FUSED_HALO_MAX=8000000 -- $S --halo peer --problem laplacian
run split_fused8M_vdc_gmres HB200_FUSED_HALO_MAX=8000000 -- $S --halo peer --problem vardifconv --solver gmres --size 256 --global-size
