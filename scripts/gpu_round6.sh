#!/bin/bash
TAG=${1:-r6}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
python scripts/spmv_sweep.py 256 27pt all > $OUT/sweep27.log 2>&1; head -4 $OUT/sweep27.log
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench1.log 2>&1; echo "bench1 exit $?"; tail -1 $OUT/bench1.log | cut -c1-330
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-graph > $OUT/bench1_nograph.log 2>&1; echo "bench1 nograph exit $?"; tail -1 $OUT/bench1_nograph.log | cut -c1-330
for mode in "--halo peer" "--halo peer --no-graph" "--halo nccl"; do
tag=$(echo $mode | tr -d ' -')
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 5 --warmup 3 $mode > $OUT/bench2_$tag.log 2>&1; echo "bench2 $mode exit $?"; tail -1 $OUT/bench2_$tag.log | cut -c1-330
done
