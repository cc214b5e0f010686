#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, launch list.  Outputs under gpurun_out/.
# usage: scripts/gpu_round.sh [tag] [bench args...]
TAG=${1:-r1}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
nproc > $OUT/nproc.txt; free -g >> $OUT/nproc.txt
export PYTHONPATH=$PWD
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 --timeout=900 > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
tail -30 $OUT/pytest.log
echo "== smoke"
timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log
tail -5 $OUT/smoke.log
echo "== bench"
timeout 1500 python bench.py "$@" > $OUT/bench.log 2>&1; echo "bench exit $?" | tee -a $OUT/bench.log
tail -5 $OUT/bench.log
if grep -q '"metric"' $OUT/bench.log && [ -n "$BENCH2" ]; then
  echo "== bench2: $BENCH2"
  timeout 1800 python bench.py $BENCH2 > $OUT/bench2.log 2>&1; echo "bench2 exit $?" | tee -a $OUT/bench2.log
  tail -3 $OUT/bench2.log
fi
