#!/bin/bash
# 2-GPU session: full GPU test suite (incl. 2-rank parity), 1-GPU bench, 2-GPU bench
TAG=${1:-r2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
nvidia-smi -L > $OUT/gpus.txt
timeout 1800 python -m pytest tests -q -m gpu --maxfail=10 --timeout=1200 > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
tail -15 $OUT/pytest.log
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > $OUT/bench1.log 2>&1; echo "bench1 exit $?"; tail -2 $OUT/bench1.log | cut -c1-1500
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench2.log 2>&1; echo "bench2 exit $?"; tail -4 $OUT/bench2.log | cut -c1-1500
