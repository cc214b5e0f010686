#!/bin/bash
TAG=${1:-r8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
nvidia-smi -L > $OUT/gpus.txt; df -h /dev/shm /tmp >> $OUT/gpus.txt; nproc >> $OUT/gpus.txt; free -g >> $OUT/gpus.txt
timeout 900 python -m pytest tests -q -m gpu -k "4ranks" --timeout=800 > $OUT/pytest4.log 2>&1; echo "pytest4 exit $?"; tail -3 $OUT/pytest4.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 3 --warmup 3 > $OUT/bench8.log 2>&1; echo "bench8 exit $?"; tail -3 $OUT/bench8.log | cut -c1-1800
