#!/bin/bash
TAG=${1:-p2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
python scripts/level_sweep.py 27pt 160 > $OUT/levels_27pt.log 2>&1; cat $OUT/levels_27pt.log | cut -c1-260
python scripts/level_sweep.py laplacian 200 > $OUT/levels_7pt.log 2>&1; cat $OUT/levels_7pt.log | cut -c1-260
echo "== ncu full (vector kernel K=2,4)"
cat > /tmp/one.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch, hypre_b200 as hb
from hypre_b200._lib import lib, check
from oracle import refbridge as rb
hb.init(0); rb.load()
pb = rb.Problem("27pt", (256, 256, 256))
A = hb.ParCSRMatrix.from_view(pb.level_view(0, 0))
x = torch.randn(A.num_rows, dtype=torch.float64, device="cuda"); y = torch.empty_like(x); torch.cuda.synchronize()
for k, L in ((1, 2), (1, 4)):
    A.set_spmv_kernel(k, L)
    for _ in range(3): check(lib.hb200_parcsr_matvec(A.handle, 1.0, x.data_ptr(), 0.0, y.data_ptr(), y.data_ptr()))
    hb.sync()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_vector -c 6 -o $OUT/prof_vec python /tmp/one.py > $OUT/ncu_vec.log 2>&1
tail -3 $OUT/ncu_vec.log
