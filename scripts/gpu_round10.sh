#!/bin/bash
# 1 GPU: timers attribution, ncu full captures of the fine-level kernels, ncu launch list of the bench
TAG=${1:-r10}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "format" 2>&1 | tail -3
HB200_TIMERS=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/t1.log 2>&1; grep -A12 "hb200 timers" $OUT/t1.log | tail -14
cat > /tmp/one.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch, hypre_b200 as hb
from hypre_b200._lib import lib, check
from oracle import refbridge as rb
hb.init(0); rb.load()
pb = rb.Problem("27pt", (256, 256, 256))
A = hb.ParCSRMatrix.from_view(pb.level_view(0, 0))
x = torch.randn(A.num_rows, dtype=torch.float64, device="cuda"); y = torch.empty_like(x); torch.cuda.synchronize()
for k, L in ((7, 0), (6, 0), (1, 0)):
    A.set_spmv_kernel(k, L)
    for _ in range(3): check(lib.hb200_parcsr_matvec(A.handle, 1.0, x.data_ptr(), 0.0, y.data_ptr(), y.data_ptr()))
    hb.sync()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_ -c 9 -o $OUT/prof_fine python /tmp/one.py > $OUT/ncu_fine.log 2>&1
tail -2 $OUT/ncu_fine.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > $OUT/ncu_bench.log 2>&1
tail -2 $OUT/ncu_bench.log; wc -l $OUT/launches.csv
