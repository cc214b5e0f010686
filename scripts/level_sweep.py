"""Per-level kernel sweep: for every A_l, P_l, P_l^T of a hierarchy, time each SpMV kernel
variant and print the fastest (data for the per-level kernel choice in dcsr_choose_kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import hypre_b200 as hb
from hypre_b200._lib import lib, check
from oracle import refbridge as rb
kind = sys.argv[1] if len(sys.argv) > 1 else "27pt"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
hb.init(0); rb.load()
pb = rb.Problem(kind, (n, n, n)); pb.setup_amg(relax_type=18)
mats, amg = hb.amg_from_hierarchy(pb.hierarchy())
stream = torch.cuda.ExternalStream(lib.hb200_compute_stream())
def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    hb.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); hb.sync()
    return e0.elapsed_time(e1) / reps
cfgs = [(0, 0), (9, 0), (7, 0), (6, 0), (1, 1), (1, 2), (1, 4), (1, 8), (1, 16), (1, 32), (8, 1), (8, 2), (8, 4), (8, 8), (8, 16), (8, 32)]
if len(sys.argv) > 3: cfgs = [(0, 0), (1, 0), (8, 0)]
print(f"# {kind} n={n}")
for l, (A, P) in enumerate(mats):
    for name, M, T in (("A", A, False), ("P", P, False), ("PT", P, True)):
        if M is None: continue
        nr, nc = (M.num_cols, M.num_rows) if T else (M.num_rows, M.num_cols)
        x = torch.randn(nc, dtype=torch.float64, device="cuda"); y = torch.zeros(nr, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        res = []
        for k, L in cfgs:
            M.set_spmv_kernel(k, L)
            if T: fn = lambda: check(lib.hb200_parcsr_matvecT(M.handle, 1.0, x.data_ptr(), 0.0, y.data_ptr()))
            else: fn = lambda: check(lib.hb200_parcsr_matvec(M.handle, 1.0, x.data_ptr(), 0.0, y.data_ptr(), y.data_ptr()))
            res.append((timeit(fn), k, L))
        M.set_spmv_kernel(0, 0)
        nnz = M.num_nonzeros
        byt = 12.0 * nnz + 4.0 * nr + 8.0 * nc + 8.0 * nr
        best = min(res)
        line = " ".join(f"{k}/{L}:{ms*1e3:.1f}" for ms, k, L in res)
        print(f"L{l} {name:2s} rows {nr:9d} nnz/row {nnz/max(nr,1):6.1f} best {best[1]}/{best[2]} {best[0]*1e3:8.1f} us {byt/best[0]/1e6:7.0f} GB/s | {line}")
