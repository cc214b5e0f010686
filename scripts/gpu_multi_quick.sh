#!/bin/bash
# N-GPU quick check at small size (an N-GPU call is charged N x the box time: keep it under a minute).
# usage: gpurun --gpus N --timeout 300 -- 'bash scripts/gpu_multi_quick.sh tag N [size]'
# Runs the bench at 64^3 rows per GPU with both halo transports under HB200_TRACE=1 and a short
# timeout each, so that a rank lost in a collective shows its last step instead of holding the box.
TAG=${1:-mq}
NG=${2:-8}
SZ=${3:-64}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
# the peer halo runs four ways: V-cycle graph on/off, NVLS on/off (the coarse GE gather is an NCCL
# all-reduce inside the captured cycle; NVLS-class algorithms only exist above two ranks)
for cfg in "nccl - -" "nccl graphnccl -" "peer - -" "peer nograph -" "peer - nonvls" "peer nograph nonvls"; do
  set -- $cfg; halo=$1; log=$OUT/${1}_${2}_${3}.log
  extra=""; [ "$2" = nograph ] && extra="--no-graph"
  if [ "$2" = graphnccl ]; then export HB200_GRAPH_NCCL=1; else unset HB200_GRAPH_NCCL; fi   # NCCL halo inside the V-cycle graph (opt-in)
  if [ "$3" = nonvls ]; then export NCCL_NVLS_ENABLE=0; else unset NCCL_NVLS_ENABLE; fi
  echo "#### halo=$halo graph=${2} nvls=${3}"
  HB200_TRACE=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
    --master-port 29521 bench.py --gpus $NG --size $SZ --steps 2 --warmup 2 --no-cpu-baseline --halo $halo $extra \
    --stage-timeout 60 > $log 2>&1
  echo "== $halo rc=$?"
  grep '^{' $log | tail -1 | cut -c1-330
  grep "no progress\|Error\|error flag" $log | head -5
  grep "hb200 trace rank 0\]\|bench rank 0\]" $log | tail -6
done
