#!/bin/bash
# round 2, N-GPU session: multi-rank parity on hardware (pytest -m gpu un-skips what fits N), then the
# bench at N with each halo transport / run-time switch under a watchdog, then timers attribution.
# usage: gpurun --gpus N --timeout 1500 -- 'bash scripts/r2_gpu_multi.sh tag N [full|quick]'
TAG=${1:-r2c}
NG=${2:-2}
MODE=${3:-full}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
export HB200_HALO_TIMEOUT_S=20
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > $OUT/gpus.txt
if [ "$MODE" = full ]; then
  echo "#### pytest -m gpu on $NG GPUs"
  timeout 1500 python -m pytest tests -q -m gpu -rs 2>&1 | tail -40 > $OUT/pytest.log; grep -v "^SKIPPED" $OUT/pytest.log | tail -5; grep -c "^SKIPPED" $OUT/pytest.log
fi
run() {  # label, env..., -- bench args
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
    print("$label rc=$rc", round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms its', c.get('iterations'), 'ms/it', round(c.get('ms_per_iteration',0),3), c.get('final_rel_res'), 'halo', c.get('halo'), 'graph', c.get('cuda_graph_vcycle'), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'upload', round(c.get('upload_s',0),2))
    if 'levels' in d.get('roofline',{}):
        for e in d['roofline']['levels']: print('    ', e['kernel'][:80], round(e['ms_per_launch'],4), round(e['frac'],3))
except Exception as ex:
    print("$label rc=$rc NO RESULT", ex)
P
  if [ $rc -ne 0 ]; then grep "no progress\|rror\|timed out" $OUT/$label.log | head -5; fi
}
S="--steps 5 --warmup 3"
echo "#### 27pt weak scaling point, N=$NG"
run peer_graph X=1 -- $S --halo peer
run peer_graph_fusewait HB200_FUSE_WAIT=1 -- $S --halo peer
run nccl_eager X=1 -- $S --halo nccl
run nccl_graph HB200_GRAPH_NCCL=1 -- $S --halo nccl
if [ "$MODE" = full ]; then
  run peer_nograph X=1 -- $S --halo peer --no-graph
  run peer_graph_nowide HB200_NO_PAT_WIDE=1 -- $S --halo peer
  run peer_graph_nobox HB200_NO_BOX=1 -- $S --halo peer
fi
echo "#### attribution (timers on => eager launches)"
HB200_TIMERS=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
   --master-port 29531 bench.py --gpus $NG --no-cpu-baseline --stage-timeout 150 --steps 1 --warmup 1 --halo peer > $OUT/timers_peer.log 2>&1
grep -A12 "hb200 timers rank 0\] hb200_pcg_solve" $OUT/timers_peer.log | tail -13
echo "#### other configs at N=$NG"
run lap7 X=1 -- $S --problem laplacian
run vdc_gmres_strong X=1 -- $S --problem vardifconv --solver gmres --size 256 --global-size
run spmv_256 X=1 -- --spmv-only --size 256 --steps 2 --warmup 2
run strong_27pt_256 X=1 -- $S --size 256 --global-size
