#!/bin/bash
# the last GPU seconds of the round: ncu launch list of one default solve with the final library
OUT=gpurun_out/r2last
mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 85 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e-ij > $OUT/launches.log 2>&1
tail -c 300 $OUT/launches.log; wc -l $OUT/launches.csv
