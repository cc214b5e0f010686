#!/bin/bash
TAG=${1:-r7}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
export HB200_TIMERS=1
timeout 900 python bench.py --gpus 1 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/t1.log 2>&1; grep -A12 "hb200 timers" $OUT/t1.log | head -14
for mode in nccl peer; do
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 1 --warmup 1 --halo $mode > $OUT/t2_$mode.log 2>&1; echo "== $mode"; grep -A12 "hb200 timers rank 0" $OUT/t2_$mode.log | head -14
done
