#!/bin/bash
TAG=${1:-r5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests -q -m gpu -k "peer_halo or 2ranks" --timeout=900 > $OUT/pytest_peer.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_peer.log; grep -n "ok   \|FAIL \|rror" $OUT/pytest_peer.log | head -20
for halo in nccl peer; do
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 --halo $halo > $OUT/bench2_$halo.log 2>&1; echo "bench2 $halo exit $?"; tail -1 $OUT/bench2_$halo.log | cut -c1-900
done
echo "== ncu full spmv_sell"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_sell -s 3 -c 2 -o $OUT/prof_sell python scripts/spmv_sweep.py 256 27pt one > $OUT/ncu_sell.log 2>&1
tail -2 $OUT/ncu_sell.log
