#!/bin/bash
# 2-GPU: timers attribution (nccl / peer / no-widen), then untimed bench lines for both halo modes
TAG=${1:-r8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
run2() { # name, extra env..., halo
  local name=$1; shift; local halo=$1; shift
  env "$@" timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 2 --halo $halo --no-cpu-baseline > $OUT/$name.log 2>&1
  echo "== $name"; grep -A12 "hb200 timers rank 0" $OUT/$name.log | tail -14; grep '^{' $OUT/$name.log | tail -1 | cut -c1-400
}
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_gpu or two_ranks" 2>&1 | tail -5
run2 t2_nccl nccl HB200_TIMERS=1 HB200_BENCH_LEVELS=1
run2 t2_peer peer HB200_TIMERS=1

run2 b2_nccl nccl HB200_X=0
run2 b2_peer peer HB200_X=0
grep "\[levels\]" $OUT/t2_nccl.log
