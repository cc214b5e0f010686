#!/bin/bash
# SpMV sweep + ncu captures.  usage: scripts/gpu_prof.sh tag
TAG=${1:-p1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
python scripts/spmv_sweep.py 256 27pt all > $OUT/sweep.log 2>&1; tail -25 $OUT/sweep.log
echo "== ncu full (stream kernels)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_stream -s 3 -c 2 -o $OUT/prof_spmv python scripts/spmv_sweep.py 256 27pt one > $OUT/ncu_spmv.log 2>&1
tail -3 $OUT/ncu_spmv.log
ls -la $OUT
