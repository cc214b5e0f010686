#!/bin/bash
TAG=${1:-r3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 2400 python -m pytest tests -q -m gpu --maxfail=10 --timeout=1200 > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
tail -15 $OUT/pytest.log
grep -n "relax type" $OUT/pytest.log | head
echo "== ncu launch list of the bench command"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
tail -2 $OUT/bench_under_ncu.log | cut -c1-300
wc -l $OUT/launches.csv
