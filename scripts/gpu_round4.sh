#!/bin/bash
TAG=${1:-r4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 2400 python -m pytest tests -q -m gpu --maxfail=10 --timeout=1200 > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
tail -12 $OUT/pytest.log
python scripts/spmv_sweep.py 256 27pt all > $OUT/sweep27.log 2>&1; cat $OUT/sweep27.log
python scripts/spmv_sweep.py 256 laplacian all > $OUT/sweep7.log 2>&1; head -8 $OUT/sweep7.log
python scripts/spmv_sweep.py 200 vardifconv all > $OUT/sweepvdc.log 2>&1; head -8 $OUT/sweepvdc.log
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > $OUT/bench1.log 2>&1; echo "bench1 exit $?"; tail -1 $OUT/bench1.log | cut -c1-2200
