#!/bin/bash
# 1 GPU: row-pattern kernel parity + sweep + bench
TAG=${1:-r9}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "matvec or format or relax_jacobi or pcg or amg" 2>&1 | tail -8
timeout 600 python scripts/spmv_sweep.py 256 27pt one > $OUT/sweep27.txt 2>&1; head -6 $OUT/sweep27.txt
timeout 600 python scripts/spmv_sweep.py 256 laplacian one > $OUT/sweep7.txt 2>&1; head -6 $OUT/sweep7.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench1.log 2>&1; grep '^{' $OUT/bench1.log | tail -1
