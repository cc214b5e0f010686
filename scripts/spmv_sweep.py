"""SpMV / vector-kernel sweep on one GPU: times every kernel variant on the fine-level matrix of
`ij -27pt -n N N N` (no AMG setup needed) and prints achieved GB/s against the byte model."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import hypre_b200 as hb
from hypre_b200._lib import lib, check
from oracle import refbridge as rb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kind = sys.argv[2] if len(sys.argv) > 2 else "27pt"
variants = sys.argv[3] if len(sys.argv) > 3 else "all"
hb.init(0)
rb.load()
pb = rb.Problem(kind, (n, n, n))
A = hb.ParCSRMatrix.from_view(pb.level_view(0, 0))
N, nnz = A.num_rows, A.num_nonzeros
stream = torch.cuda.ExternalStream(lib.hb200_compute_stream())
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
x = torch.randn(N, dtype=torch.float64, device="cuda")
y = torch.empty(N, dtype=torch.float64, device="cuda")
f = torch.randn(N, dtype=torch.float64, device="cuda")
l1 = torch.rand(N, dtype=torch.float64, device="cuda") + 26.0
torch.cuda.synchronize()

def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    hb.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream)
    hb.sync()
    return e0.elapsed_time(e1) / reps

rows = []
spmv_bytes = 12.0 * nnz + 4.0 * N + 8.0 * N + 8.0 * N
def spmv(): check(lib.hb200_parcsr_matvec(A.handle, 1.0, x.data_ptr(), 0.0, y.data_ptr(), y.data_ptr()))
if variants == "one":
    cfgs = [(9, 0), (7, 0), (6, 0)]
else:
    cfgs = [(9, 0), (7, 0), (6, 0), (1, 2), (1, 4), (2, 0)]
for k, L in cfgs:
    A.set_spmv_kernel(k, L)
    ms = timeit(spmv)
    rows.append((f"spmv kind={k} lanes={L}", ms, spmv_bytes / ms / 1e6))
A.set_spmv_kernel(0, 0)
print("# format:", A.format_info())
vt = torch.empty(N, dtype=torch.float64, device="cuda")
def jac(): check(lib.hb200_relax(A.handle, f.data_ptr(), None, 18, 0, 1.0, 1.0, l1.data_ptr(), x.data_ptr(), 0, vt.data_ptr()))
ms = timeit(jac)
rows.append(("l1-jacobi fused sweep (+copy back)", ms, (12.0 * nnz + 36.0 * N + 16.0 * N) / ms / 1e6))
def axpy(): check(lib.hb200_vec_axpy(0.5, x.data_ptr(), y.data_ptr(), N))
ms = timeit(axpy); rows.append(("axpy", ms, 24.0 * N / ms / 1e6))
def copy(): check(lib.hb200_vec_copy(x.data_ptr(), y.data_ptr(), N))
ms = timeit(copy); rows.append(("copy (cudaMemcpyAsync d2d)", ms, 16.0 * N / ms / 1e6))
from hypre_b200.solver import inner_prod
import ctypes as C
out = C.c_double()
def dot(): check(lib.hb200_vec_inner_prod(x.data_ptr(), y.data_ptr(), N, C.byref(out)))
ms = timeit(dot); rows.append(("inner_prod (+host fetch)", ms, 16.0 * N / ms / 1e6))
big = torch.empty(1 << 28, dtype=torch.float64, device="cuda"); big2 = torch.empty_like(big)
def bigcopy(): check(lib.hb200_vec_copy(big.data_ptr(), big2.data_ptr(), big.numel()))
ms = timeit(bigcopy, reps=10); rows.append(("copy 2 GiB", ms, 16.0 * big.numel() / ms / 1e6))
def bigaxpy(): check(lib.hb200_vec_axpy(0.5, big.data_ptr(), big2.data_ptr(), big.numel()))
ms = timeit(bigaxpy, reps=10); rows.append(("axpy 2 GiB vectors", ms, 24.0 * big.numel() / ms / 1e6))
print(f"# {kind} n={n}: rows {N} nnz {nnz}; peak {peak} GB/s")
for name, ms, gbs in rows:
    print(f"{name:40s} {ms:9.4f} ms {gbs:9.1f} GB/s  {gbs / peak:6.3f} of measured peak")
