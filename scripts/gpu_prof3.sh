#!/bin/bash
TAG=${1:-p3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
python scripts/spmv_sweep.py 256 27pt all > $OUT/sweep.log 2>&1; cat $OUT/sweep.log
python scripts/level_sweep.py 27pt 160 > $OUT/levels_27pt.log 2>&1; cut -c1-150 $OUT/levels_27pt.log
python scripts/level_sweep.py laplacian 200 > $OUT/levels_7pt.log 2>&1; cut -c1-150 $OUT/levels_7pt.log
