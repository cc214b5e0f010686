// parcsr.cu — device-resident hypre_ParCSRMatrix (diag + offd CSR, col_map_offd, CommPkg),
// the halo exchange, and the ParCSR matvec / transposed matvec.
//
// Reference: hypre_ParCSRMatrixMatvecOutOfPlaceHost (src/parcsr_mv/par_csr_matvec.c:21-232),
// hypre_ParCSRMatrixMatvecTHost (:288-514), hypre_ParCSRCommHandleCreate_v2 / Destroy
// (src/parcsr_mv/par_csr_communication.c:358-723).
//
// Data flow of y = alpha*A*x + beta*b on one rank (same ordering as the reference, :170-217):
//   comm stream : pack x[send_map_elmts] -> send_buf ; grouped ncclSend/ncclRecv -> x_ext
//   comp stream : diag block SpMV (all rows, epilogue y = beta*b + alpha*sum)   [overlapped]
//   comp stream : wait(halo) ; offd block SpMV over the non-empty-row list, y += alpha*sum
#include "hb_internal.cuh"
#ifndef HB200_EMU
#include <cub/device/device_radix_sort.cuh>
#endif
#include <chrono>
#include <algorithm>
#include <numeric>
#include <string.h>
#include <stdlib.h>

namespace hb {

bool parcsr_main_skips_boundary(const hb200_parcsr *A)
{
   const DCsr &D = A->diag;
   return (D.kind == SPMV_PAT || D.kind == SPMV_BOX) && D.has_pat && D.pat_skips_boundary;
}

// ---------------------------------------------------------------------------------------
// halo kernels
// ---------------------------------------------------------------------------------------
__global__ void pack_kernel(int n, const int *__restrict__ map, const double *__restrict__ x,
                            double *__restrict__ buf)
{
   const int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k < n) buf[k] = x[map[k]];
}

// y[rows[t]] += buf[idx[ptr[t]]] + buf[idx[ptr[t]+1]] + ...  in ascending send-entry order:
// the order of the reference's sequential loop (par_csr_matvec.c:491-496)
__global__ void unpack_add_kernel(int nrows, const int *__restrict__ rows,
                                  const int *__restrict__ ptr, const int *__restrict__ idx,
                                  const double *__restrict__ buf, double *__restrict__ y)
{
   const int t = blockIdx.x * blockDim.x + threadIdx.x;
   if (t < nrows) {
      double v = y[rows[t]];
      for (int q = ptr[t]; q < ptr[t + 1]; q++) v = __dadd_rn(v, buf[idx[q]]);
      y[rows[t]] = v;
   }
}

static int exchange(hb200_parcsr *A, bool forward, cudaStream_t st)
{
   // forward (job 1): send send_buf segments to send_procs, receive x_ext segments from
   // recv_procs.  reverse (job 2): roles swapped (par_csr_communication.c:381-410).
   Ctx &c = ctx();
   CommPkgD &pk = A->pkg;
   if (pk.num_sends == 0 && pk.num_recvs == 0) return 0;
#ifdef HB200_WITH_NCCL
   HB_REQUIRE(c.nccl != nullptr, HB200_ERROR_GENERIC,
              "matrix has off-processor couplings but hb200_comm_init was not called");
   HB_NCCL(nccl_api().GroupStart());
   for (int i = 0; i < pk.num_recvs; i++) {
      const int cnt = pk.recv_vec_starts[i + 1] - pk.recv_vec_starts[i];
      double *p = (forward ? pk.d_recv_buf : A->d_ytmp) + pk.recv_vec_starts[i];
      if (forward) HB_NCCL(nccl_api().Recv(p, cnt, ncclDouble, pk.recv_procs[i], c.nccl, st));
      else         HB_NCCL(nccl_api().Send(p, cnt, ncclDouble, pk.recv_procs[i], c.nccl, st));
   }
   for (int i = 0; i < pk.num_sends; i++) {
      const int cnt = pk.send_map_starts[i + 1] - pk.send_map_starts[i];
      double *p = pk.d_send_buf + pk.send_map_starts[i];
      if (forward) HB_NCCL(nccl_api().Send(p, cnt, ncclDouble, pk.send_procs[i], c.nccl, st));
      else         HB_NCCL(nccl_api().Recv(p, cnt, ncclDouble, pk.send_procs[i], c.nccl, st));
   }
   HB_NCCL(nccl_api().GroupEnd());
   return 0;
#else
   (void) forward; (void) st; (void) c;
   return set_error(HB200_ERROR_GENERIC, "libhb200 built without NCCL");
#endif
}

int parcsr_halo_begin(hb200_parcsr *A, const double *x, cudaStream_t st_comp)
{
   Ctx &c = ctx();
   CommPkgD &pk = A->pkg;
   if (c.halo_mode == 1 && c.nranks > 1) {
      // NVLink peer puts, everything on the compute stream (parcsr_peer.cu); the plan build is
      // collective, so ranks without neighbours on this matrix take part too
      HB_CHECK(peer_plans_ensure(A, false));
      if (!pk.peer_off) {
         if (!peer_has_out(pk.fwd)) return 0;
         // the put (gather + NVLink stores + system fences, ~10 us of latency) runs beside the diag pass
         HB_CUDA(cudaEventRecord(c.ev_a, st_comp));
         HB_CUDA(cudaStreamWaitEvent(c.s_comm, c.ev_a, 0));
         HB_CHECK(peer_put(pk.fwd, x, c.s_comm));
         HB_CUDA(cudaEventRecord(c.ev_b, c.s_comm));
         return 0;
      }
      // (the ranks agreed that this matrix has no peer plan: NCCL below)
   }
   if (pk.num_sends == 0 && pk.num_recvs == 0) return 0;
   HB_CUDA(cudaEventRecord(c.ev_a, st_comp));
   HB_CUDA(cudaStreamWaitEvent(c.s_comm, c.ev_a, 0));
   const int nsend = pk.num_sends ? pk.send_map_starts[pk.num_sends] : 0;
   if (nsend > 0) {
      HB_LAUNCH(pack_kernel, (nsend + 255) / 256, 256, 0, c.s_comm, nsend, pk.d_send_map_elmts, x,
                pk.d_send_buf);
      HB_LAUNCH_CHECK();
   }
   return exchange(A, true, c.s_comm);
}

int parcsr_halo_end(hb200_parcsr *A, cudaStream_t st_comp)
{
   Ctx &c = ctx();
   CommPkgD &pk = A->pkg;
   if (c.halo_mode == 1 && c.nranks > 1 && !pk.peer_off) {
      if (peer_has_out(pk.fwd)) HB_CUDA(cudaStreamWaitEvent(st_comp, c.ev_b, 0));
      return pk.fwd ? peer_wait(pk.fwd, st_comp) : 0;
   }
   if (pk.num_sends == 0 && pk.num_recvs == 0) return 0;
   HB_CUDA(cudaEventRecord(c.ev_b, c.s_comm));
   HB_CUDA(cudaStreamWaitEvent(st_comp, c.ev_b, 0));
   return 0;
}

static bool fuse_wait_enabled()
{
   static const bool on = env_flag("HB200_FUSE_WAIT", false);   // opt-in until measured on hardware
   return on;
}

// the second half of a ParCSR operation: wait for the halo, then the offd block over the boundary
// rows with the accumulate epilogue `epi_kind`.  Peer-put halo + HB200_FUSE_WAIT=1: one kernel does
// both and reads the receive buffer in place (kernels_offd.cu).
int parcsr_offd_pass(hb200_parcsr *A, int epi_kind, const EpiArgs &ea)
{
   Ctx &c = ctx();
   CommPkgD &pk = A->pkg;
   PeerWaitArgs w;
   if (c.halo_mode == 1 && c.nranks > 1 && !pk.peer_off && fuse_wait_enabled() && A->num_cols_offd > 0 && A->offd.num_rownnz > 0 &&
       peer_wait_args(pk.fwd, &w)) {
      timer_tick(T_HALO_WAIT);
      if (peer_has_out(pk.fwd)) HB_CUDA(cudaStreamWaitEvent(c.s_comp, c.ev_b, 0));   // the local put is done
      timer_tick(T_MATVEC_OFFD);
      return spmv_offd_wait_launch(A->offd, w, epi_kind, ea, c.s_comp);
   }
   timer_tick(T_HALO_WAIT);
   HB_CHECK(parcsr_halo_end(A, c.s_comp));
   timer_tick(T_MATVEC_OFFD);
   if (A->num_cols_offd > 0) {
      HB_CHECK(spmv_launch(A->offd, A->pkg.d_recv_buf, epi_kind, ea, true, c.s_comp));
   }
   return 0;
}

int parcsr_matvec(hb200_parcsr *A, double alpha, const double *x, double beta, const double *b,
                  double *y)
{
   Ctx &c = ctx();
   if (alpha == 0.0) {
      c.dot_req_armed = false; c.last_dot_fused = false;
      // csr_matvec.c:92-105: y = beta*b
      if (beta == 0.0) return vec_set(y, 0.0, A->num_rows, c.s_comp);
      return vec_axpby_out(beta, b, 0.0, b, y, A->num_rows, c.s_comp);
   }
   EpiArgs ea;
   ea.alpha = alpha; ea.beta = beta; ea.b = b; ea.y = y;
   if (!c.dot_req_armed) {
      // latency-bound level over the peer halo: the whole operation is one kernel (kernels_offd.cu)
      bool done = false;
      HB_CHECK(parcsr_fused_try(A, x, EPI_AXPBY, ea, &done));
      if (done) { c.last_dot_fused = false; return 0; }
   }
   if (parcsr_main_skips_boundary(A)) {
      // split operation: the main kernel computes the rows without offd entries, the boundary kernel the others
      c.dot_req_armed = false; c.last_dot_fused = false;
      const bool peer = (c.halo_mode == 1 && c.nranks > 1);
      if (peer) HB_CHECK(peer_plans_ensure(A, false));
      if (peer && !A->pkg.peer_off) {
         timer_tick(T_HALO_START);
         HB_CUDA(cudaEventRecord(c.ev_a, c.s_comp));
         HB_CUDA(cudaStreamWaitEvent(c.s_comm, c.ev_a, 0));
         HB_CHECK(parcsr_boundary_launch(A, x, EPI_AXPBY, ea, true, c.s_comm));   // put + flags + boundary rows, side stream
         HB_CUDA(cudaEventRecord(c.ev_b, c.s_comm));
         timer_tick(T_MATVEC_DIAG);
         HB_CHECK(spmv_launch(A->diag, x, EPI_AXPBY, ea, false, c.s_comp));
         HB_CUDA(cudaStreamWaitEvent(c.s_comp, c.ev_b, 0));
      } else {
         timer_tick(T_HALO_START);
         HB_CHECK(parcsr_halo_begin(A, x, c.s_comp));
         timer_tick(T_MATVEC_DIAG);
         HB_CHECK(spmv_launch(A->diag, x, EPI_AXPBY, ea, false, c.s_comp));
         timer_tick(T_HALO_WAIT);
         HB_CHECK(parcsr_halo_end(A, c.s_comp));
         timer_tick(T_MATVEC_OFFD);
         HB_CHECK(parcsr_boundary_launch(A, x, EPI_AXPBY, ea, false, c.s_comp));
      }
      timer_tick(T_OTHER);
      return 0;
   }
   timer_tick(T_HALO_START);
   HB_CHECK(parcsr_halo_begin(A, x, c.s_comp));
   timer_tick(T_MATVEC_DIAG);
   // an armed fused-dot request (<y, w>) is honoured when y is final after the diag pass
   const bool want_dot = c.dot_req_armed;
   c.dot_req_armed = false;
   c.last_dot_fused = false;
   if (want_dot && c.nranks == 1 && A->num_cols_offd == 0 && spmv_can_fuse_dot(A->diag, EPI_AXPBY)) {
      ea.dotw = c.dot_req_w; ea.dot_slot = c.dot_req_slot;
      c.last_dot_fused = true;
   }
   HB_CHECK(spmv_launch(A->diag, x, EPI_AXPBY, ea, false, c.s_comp));
   EpiArgs eo;
   eo.alpha = alpha; eo.y = y;
   HB_CHECK(parcsr_offd_pass(A, EPI_ACC, eo));
   timer_tick(T_OTHER);
   return 0;
}

__global__ void diag_extract_kernel(int n, const int *__restrict__ di, const double *__restrict__ da,
                                    double *__restrict__ out)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) out[i] = da[di[i]];   // first entry of each diag row is the diagonal (par_relax.c:274)
}

// diagonal of A as a device vector (built on first use)
int parcsr_diag(hb200_parcsr *A, const double **out)
{
   Ctx &c = ctx();
   if (!A->d_diaginv) {
      const size_t n = (size_t) A->num_rows;
      HB_CUDA(cudaMalloc(&A->d_diaginv, sizeof(double) * (n ? n : 1)));
      if (n) {
         HB_LAUNCH(diag_extract_kernel, (int) ((n + 255) / 256), 256, 0, c.s_comp, (int) n, A->diag.i,
                   A->diag.a, A->d_diaginv);
         HB_LAUNCH_CHECK();
      }
   }
   *out = A->d_diaginv;
   return 0;
}

// download a device CSR block to the host (used to build transposes / schedules lazily)
static int dcsr_download(const DCsr &M, std::vector<int> &hi, std::vector<int> &hj,
                         std::vector<double> &ha)
{
   hi.resize((size_t) M.nrows + 1);
   hj.resize((size_t) M.nnz);
   ha.resize((size_t) M.nnz);
   HB_CUDA(cudaMemcpy(hi.data(), M.i, sizeof(int) * hi.size(), cudaMemcpyDeviceToHost));
   if (M.nnz) {
      HB_CUDA(cudaMemcpy(hj.data(), M.j, sizeof(int) * hj.size(), cudaMemcpyDeviceToHost));
      HB_CUDA(cudaMemcpy(ha.data(), M.a, sizeof(double) * ha.size(), cudaMemcpyDeviceToHost));
   }
   return 0;
}

#ifndef HB200_EMU
// ---- stored transpose on the device (hypre_CSRMatrixTranspose / hypre_ParCSRMatrixLocalTranspose,
// src/parcsr_mv/par_csr_matop.c:2200): a STABLE sort of the entries by column — entries of one output
// row keep their ascending source-row order, the order hypre_CSRMatrixMatvecTHost accumulates in
// (csr_matvec.c:1095-1110) — so the result is the same array the host transpose produces.
__global__ void tr_entry_rows_kernel(int nrows, const int *__restrict__ ai, int *__restrict__ rows, int *__restrict__ idx)
{
   const int r = blockIdx.x * blockDim.x + threadIdx.x;
   if (r < nrows) {
      for (int p = ai[r]; p < ai[r + 1]; p++) { rows[p] = r; idx[p] = p; }
   }
}
__global__ void tr_gather_kernel(long long nnz, const int *__restrict__ perm, const int *__restrict__ rows,
                                 const double *__restrict__ a, int *__restrict__ tj, double *__restrict__ ta)
{
   const long long q = (long long) blockIdx.x * blockDim.x + threadIdx.x;
   if (q < nnz) { const int p = perm[q]; tj[q] = rows[p]; ta[q] = a[p]; }
}
// row pointer of the transpose from the sorted column keys: ti[c] = first position whose key >= c
__global__ void tr_rowptr_kernel(long long nnz, int ncols, const int *__restrict__ keys, int *__restrict__ ti)
{
   const long long q = (long long) blockIdx.x * blockDim.x + threadIdx.x;
   if (q > nnz) return;
   const int lo = q == 0 ? -1 : keys[q - 1];
   const int hi = q == nnz ? ncols : keys[q];
   for (int c = lo + 1; c <= hi; c++) ti[c] = (int) q;
}

static int dcsr_transpose_device(const DCsr &M, int **ti_out, int **tj_out, double **ta_out)
{
   Ctx &c = ctx();
   const long long nnz = M.nnz;
   const int n = M.nrows, m = M.ncols;
   const auto t0 = std::chrono::steady_clock::now();
   int *rows = nullptr, *idx = nullptr, *keys = nullptr, *perm = nullptr, *ti = nullptr, *tj = nullptr;
   double *ta = nullptr;
   void *tmp = nullptr;
   size_t tmp_bytes = 0;
   auto cleanup = [&]() { cudaFree(rows); cudaFree(idx); cudaFree(keys); cudaFree(perm); cudaFree(tmp); };
   HB_CUDA(cudaMalloc(&rows, sizeof(int) * (size_t) nnz));
   HB_CUDA(cudaMalloc(&idx, sizeof(int) * (size_t) nnz));
   HB_CUDA(cudaMalloc(&keys, sizeof(int) * (size_t) nnz));
   HB_CUDA(cudaMalloc(&perm, sizeof(int) * (size_t) nnz));
   HB_CUDA(cudaMalloc(&ti, sizeof(int) * ((size_t) m + 1)));
   HB_CUDA(cudaMalloc(&tj, sizeof(int) * ((size_t) nnz + 8)));
   HB_CUDA(cudaMalloc(&ta, sizeof(double) * ((size_t) nnz + 8)));
   HB_CUDA(cudaMemsetAsync(tj + nnz, 0, sizeof(int) * 8, c.s_comp));
   HB_CUDA(cudaMemsetAsync(ta + nnz, 0, sizeof(double) * 8, c.s_comp));
   const auto t1 = std::chrono::steady_clock::now();
   HB_LAUNCH(tr_entry_rows_kernel, (n + 255) / 256, 256, 0, c.s_comp, n, M.i, rows, idx);
   int bits = 1;
   while (bits < 31 && (1LL << bits) < (long long) m) bits++;
   cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, M.j, keys, idx, perm, (int) nnz, 0, bits, c.s_comp);
   if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 8);
   if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, M.j, keys, idx, perm, (int) nnz, 0, bits, c.s_comp);
   if (e != cudaSuccess) {
      cleanup(); cudaFree(ti); cudaFree(tj); cudaFree(ta);
      return set_error(HB200_ERROR_GENERIC, "device transpose: radix sort failed: %s", cudaGetErrorString(e));
   }
   HB_LAUNCH(tr_gather_kernel, (int) ((nnz + 255) / 256), 256, 0, c.s_comp, nnz, perm, rows, M.a, tj, ta);
   HB_LAUNCH(tr_rowptr_kernel, (int) ((nnz + 1 + 255) / 256), 256, 0, c.s_comp, nnz, m, keys, ti);
   HB_LAUNCH_CHECK();
   HB_CUDA(cudaStreamSynchronize(c.s_comp));
   const auto t2 = std::chrono::steady_clock::now();
   cleanup();
   if (nnz >= 1000000) {
      HB_TRACE("device transpose of %lld entries: allocations %.3f s, kernels + sort %.3f s, frees %.3f s", nnz,
               std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count(),
               std::chrono::duration<double>(std::chrono::steady_clock::now() - t2).count());
   }
   *ti_out = ti; *tj_out = tj; *ta_out = ta;
   return 0;
}
#endif

// stored transpose of one block: on the device when the block is big enough to matter (the host
// transpose of P_0 at 256^3 took ~2 s: a random scatter over 60 M entries), on the host otherwise
static int dcsr_build_transpose(const DCsr &M, DCsr &T)
{
   std::vector<int> hi, hj, ti, tj;
   std::vector<double> ha, ta;
   const auto t0 = std::chrono::steady_clock::now();
#ifndef HB200_EMU
   const char *thr = getenv("HB200_DEVICE_TRANSPOSE_MIN");   // (tests lower it to run small blocks through the device path)
   const long long min_nnz = thr ? atoll(thr) : 65536;
   if (M.nnz >= min_nnz && M.nnz > 0 && M.nnz < 0x7fffffffLL && !env_flag("HB200_HOST_TRANSPOSE", false)) {
      int *d_ti = nullptr, *d_tj = nullptr;
      double *d_ta = nullptr;
      HB_CHECK(dcsr_transpose_device(M, &d_ti, &d_tj, &d_ta));
      const auto t1 = std::chrono::steady_clock::now();
      T.nrows = M.ncols; T.ncols = M.nrows; T.nnz = M.nnz;
      T.i = d_ti; T.j = d_tj; T.a = d_ta;
      // the format analysis reads the arrays on the host
      ti.resize((size_t) T.nrows + 1); tj.resize((size_t) T.nnz); ta.resize((size_t) T.nnz);
      HB_CUDA(cudaMemcpy(ti.data(), T.i, sizeof(int) * ti.size(), cudaMemcpyDeviceToHost));
      HB_CUDA(cudaMemcpy(tj.data(), T.j, sizeof(int) * tj.size(), cudaMemcpyDeviceToHost));
      HB_CUDA(cudaMemcpy(ta.data(), T.a, sizeof(double) * ta.size(), cudaMemcpyDeviceToHost));
      const auto t2 = std::chrono::steady_clock::now();
      HB_CHECK(dcsr_analyze(T, ti.data(), tj.data(), ta.data()));
      HB_TRACE("stored transpose of a %d x %d block (%lld nnz): device sort %.3f s, download %.3f s, analysis %.3f s", M.nrows,
               M.ncols, M.nnz, std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count(),
               std::chrono::duration<double>(std::chrono::steady_clock::now() - t2).count());
      return 0;
   }
#endif
   HB_CHECK(dcsr_download(M, hi, hj, ha));
   const auto t1 = std::chrono::steady_clock::now();
   host_csr_transpose(M.nrows, M.ncols, hi.data(), hj.data(), ha.data(), ti, tj, ta);
   const auto t2 = std::chrono::steady_clock::now();
   HB_CHECK(dcsr_upload(T, M.ncols, M.nrows, ti.data(), tj.data(), ta.data()));
   if (M.nrows >= 1024) {
      HB_TRACE("stored transpose of a %d x %d block: download %.3f s, host transpose %.3f s, upload %.3f s", M.nrows,
               M.ncols, std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count(),
               std::chrono::duration<double>(std::chrono::steady_clock::now() - t2).count());
   }
   return 0;
}

int parcsr_ensure_T(hb200_parcsr *A)
{
   if (A->has_T) return 0;
   // stored transposes, as the reference does under keepTranspose (par_csr_matvec.c:298-299,
   // 430-468): restriction stays a row-parallel, atomics-free, deterministic SpMV
   HB_CHECK(dcsr_build_transpose(A->diag, A->diagT));
   if (A->num_cols_offd > 0) {
      HB_CHECK(dcsr_build_transpose(A->offd, A->offdT));
      HB_CUDA(cudaMalloc(&A->d_ytmp, sizeof(double) * (size_t) A->num_cols_offd));
   }
   A->has_T = true;
   return 0;
}

int parcsr_matvecT(hb200_parcsr *A, double alpha, const double *x, double beta, double *y)
{
   Ctx &c = ctx();
   HB_CHECK(parcsr_ensure_T(A));
   CommPkgD &pk = A->pkg;
   if (alpha != 0.0) {
      bool done = false;
      HB_CHECK(parcsr_fusedT_try(A, alpha, x, beta, y, &done));   // latency-bound level: one kernel
      if (done) return 0;
   }
   bool peer = (c.halo_mode == 1 && c.nranks > 1);
   const bool comm = (pk.num_sends || pk.num_recvs);
   if (peer) {
      HB_CHECK(peer_plans_ensure(A, true));
      if (pk.peer_off_rev) peer = false;   // agreed on every rank: NCCL for this matrix
   }
   timer_tick(T_MATVEC_OFFD);
   if (A->num_cols_offd > 0) {
      // y_tmp = alpha * offd^T x  (par_csr_matvec.c:402-420)
      EpiArgs eo;
      eo.alpha = alpha; eo.beta = 0.0; eo.y = A->d_ytmp;
      HB_CHECK(spmv_launch(A->offdT, x, EPI_AXPBY, eo, false, c.s_comp));
   }
   timer_tick(T_HALO_START);
   if (peer) {
      if (peer_has_out(pk.rev)) {
         HB_CUDA(cudaEventRecord(c.ev_a, c.s_comp));
         HB_CUDA(cudaStreamWaitEvent(c.s_comm, c.ev_a, 0));
         HB_CHECK(peer_put(pk.rev, A->d_ytmp, c.s_comm));
         HB_CUDA(cudaEventRecord(c.ev_b, c.s_comm));
      }
   } else if (comm) {
      HB_CUDA(cudaEventRecord(c.ev_a, c.s_comp));
      HB_CUDA(cudaStreamWaitEvent(c.s_comm, c.ev_a, 0));
      HB_CHECK(exchange(A, false, c.s_comm));
   }
   // y = alpha * diag^T x + beta * y   (overlapped with the reverse exchange)
   timer_tick(T_MATVEC_DIAG);
   EpiArgs ed;
   ed.alpha = alpha; ed.beta = beta; ed.b = y; ed.y = y;
   HB_CHECK(spmv_launch(A->diagT, x, EPI_AXPBY, ed, false, c.s_comp));
   timer_tick(T_HALO_WAIT);
   if (peer) {
      if (peer_has_out(pk.rev)) HB_CUDA(cudaStreamWaitEvent(c.s_comp, c.ev_b, 0));
      if (pk.rev) HB_CHECK(peer_wait(pk.rev, c.s_comp));
   } else if (comm) {
      HB_CUDA(cudaEventRecord(c.ev_b, c.s_comm));
      HB_CUDA(cudaStreamWaitEvent(c.s_comp, c.ev_b, 0));
   }
   if ((peer || comm) && pk.n_unpack_rows > 0) {
      HB_LAUNCH(unpack_add_kernel, (pk.n_unpack_rows + 255) / 256, 256, 0, c.s_comp,
                pk.n_unpack_rows, pk.d_unpack_rows, pk.d_unpack_ptr, pk.d_unpack_idx,
                pk.d_send_buf, y);
      HB_LAUNCH_CHECK();
   }
   timer_tick(T_OTHER);
   return 0;
}

}  // namespace hb

using namespace hb;

extern "C" {

int hb200_parcsr_create(hb200_parcsr **Aout, int num_rows, int num_cols, int num_cols_offd,
                        const int *diag_i, const int *diag_j, const double *diag_data,
                        const int *offd_i, const int *offd_j, const double *offd_data,
                        const int64_t *col_map_offd, int64_t first_row_index,
                        int64_t first_col_diag, int64_t global_num_rows, int64_t global_num_cols,
                        int num_sends, const int *send_procs, const int *send_map_starts,
                        const int *send_map_elmts, int num_recvs, const int *recv_procs,
                        const int *recv_vec_starts)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(Aout != nullptr, HB200_ERROR_ARG, "null output handle");
   HB_REQUIRE(num_rows >= 0 && num_cols >= 0 && num_cols_offd >= 0, HB200_ERROR_ARG, "negative size");
   HB_REQUIRE(num_rows == 0 || diag_i != nullptr, HB200_ERROR_ARG, "diag_i is NULL");
   HB_REQUIRE(num_cols_offd == 0 || (offd_i && col_map_offd), HB200_ERROR_ARG,
              "offd block given without offd_i / col_map_offd");
   hb200_parcsr *A = new hb200_parcsr();
   A->num_rows = num_rows;
   A->num_cols = num_cols;
   A->num_cols_offd = num_cols_offd;
   A->first_row = first_row_index;
   A->first_col = first_col_diag;
   A->global_rows = global_num_rows;
   A->global_cols = global_num_cols;
   int zero = 0;
   // N > 1: the structured formats of the diag block leave the rows with offd entries to the boundary kernel
   // (the split operation: main kernel beside put + boundary rows; opt-in with HB200_SPLIT=1 until it wins on hardware)
   if (ctx().nranks > 1 && num_cols_offd > 0 && num_rows > 0 && env_flag("HB200_SPLIT", false)) g_pat_boundary_offd_i = offd_i;
   int f = dcsr_upload(A->diag, num_rows, num_cols, num_rows ? diag_i : &zero, diag_j, diag_data);
   g_pat_boundary_offd_i = nullptr;
   if (f) { hb200_parcsr_destroy(A); return f; }   // frees what the failed upload had already allocated
   if (num_cols_offd > 0) {
      f = dcsr_upload(A->offd, num_rows, num_cols_offd, offd_i, offd_j, offd_data);
      if (f) { hb200_parcsr_destroy(A); return f; }
      // offd blocks always go through the vector kernel over the non-empty-row list
      dcsr_choose_kernel(A->offd, SPMV_VECTOR, 0);
      A->col_map_offd.assign(col_map_offd, col_map_offd + num_cols_offd);
      HB_CUDA(cudaMalloc(&A->d_col_map_offd, sizeof(int64_t) * (size_t) num_cols_offd));
      HB_CUDA(cudaMemcpy(A->d_col_map_offd, col_map_offd, sizeof(int64_t) * (size_t) num_cols_offd,
                         cudaMemcpyHostToDevice));
   }
   CommPkgD &pk = A->pkg;
   pk.num_sends = num_sends;
   pk.num_recvs = num_recvs;
   if (num_sends > 0) {
      pk.send_procs.assign(send_procs, send_procs + num_sends);
      pk.send_map_starts.assign(send_map_starts, send_map_starts + num_sends + 1);
      const int ns = pk.send_map_starts[num_sends];
      pk.send_map_elmts.assign(send_map_elmts, send_map_elmts + ns);
      if (ns > 0) {
         HB_CUDA(cudaMalloc(&pk.d_send_map_elmts, sizeof(int) * (size_t) ns));
         HB_CUDA(cudaMemcpy(pk.d_send_map_elmts, send_map_elmts, sizeof(int) * (size_t) ns,
                            cudaMemcpyHostToDevice));
         HB_CUDA(cudaMalloc(&pk.d_send_buf, sizeof(double) * (size_t) ns));
         // deterministic MatvecT unpack plan: group send entries by target row, stable
         std::vector<int> order(ns);
         std::iota(order.begin(), order.end(), 0);
         std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            return pk.send_map_elmts[a] < pk.send_map_elmts[b];
         });
         std::vector<int> rows, ptr;
         for (int q = 0; q < ns; q++) {
            const int r = pk.send_map_elmts[order[q]];
            if (rows.empty() || rows.back() != r) { rows.push_back(r); ptr.push_back(q); }
         }
         ptr.push_back(ns);
         pk.n_unpack_rows = (int) rows.size();
         HB_CUDA(cudaMalloc(&pk.d_unpack_rows, sizeof(int) * rows.size()));
         HB_CUDA(cudaMalloc(&pk.d_unpack_ptr, sizeof(int) * ptr.size()));
         HB_CUDA(cudaMalloc(&pk.d_unpack_idx, sizeof(int) * order.size()));
         HB_CUDA(cudaMemcpy(pk.d_unpack_rows, rows.data(), sizeof(int) * rows.size(), cudaMemcpyHostToDevice));
         HB_CUDA(cudaMemcpy(pk.d_unpack_ptr, ptr.data(), sizeof(int) * ptr.size(), cudaMemcpyHostToDevice));
         HB_CUDA(cudaMemcpy(pk.d_unpack_idx, order.data(), sizeof(int) * order.size(), cudaMemcpyHostToDevice));
      }
   } else {
      pk.send_map_starts.assign(1, 0);
   }
   if (num_recvs > 0) {
      pk.recv_procs.assign(recv_procs, recv_procs + num_recvs);
      pk.recv_vec_starts.assign(recv_vec_starts, recv_vec_starts + num_recvs + 1);
   } else {
      pk.recv_vec_starts.assign(1, 0);
   }
   if (num_cols_offd > 0) {
      HB_CUDA(cudaMalloc(&pk.d_recv_buf, sizeof(double) * (size_t) num_cols_offd));
      HB_CUDA(cudaMemset(pk.d_recv_buf, 0, sizeof(double) * (size_t) num_cols_offd));
   }
   *Aout = A;
   return 0;
}

int hb200_parcsr_destroy(hb200_parcsr *A)
{
   if (!A) return 0;
   cudaDeviceSynchronize();
   dcsr_free(A->diag);
   dcsr_free(A->offd);
   dcsr_free(A->diagT);
   dcsr_free(A->offdT);
   CommPkgD &pk = A->pkg;
   if (A->d_col_map_offd) cudaFree(A->d_col_map_offd);
   if (pk.d_send_map_elmts) cudaFree(pk.d_send_map_elmts);
   if (pk.d_send_buf) cudaFree(pk.d_send_buf);
   if (pk.d_recv_buf) cudaFree(pk.d_recv_buf);
   if (pk.d_unpack_rows) cudaFree(pk.d_unpack_rows);
   if (pk.d_unpack_ptr) cudaFree(pk.d_unpack_ptr);
   if (pk.d_unpack_idx) cudaFree(pk.d_unpack_idx);
   peer_plan_free(pk.fwd);
   peer_plan_free(pk.rev);
   if (A->d_ytmp) cudaFree(A->d_ytmp);
   if (A->d_unpack_slot) cudaFree(A->d_unpack_slot);
   if (A->d_diaginv) cudaFree(A->d_diaginv);
   if (A->gs_sched) gs_sched_free(A->gs_sched);
   delete A;
   return 0;
}

int hb200_parcsr_num_rows(const hb200_parcsr *A) { return A ? A->num_rows : -1; }
int hb200_parcsr_num_cols(const hb200_parcsr *A) { return A ? A->num_cols : -1; }
long long hb200_parcsr_num_nonzeros(const hb200_parcsr *A)
{
   return A ? A->diag.nnz + A->offd.nnz : -1;
}

int hb200_parcsr_download_maps(const hb200_parcsr *A, int *diag_i, int *diag_j, int *offd_i,
                               int *offd_j, int64_t *col_map_offd, int *send_map_starts,
                               int *send_map_elmts, int *recv_vec_starts, int *send_procs,
                               int *recv_procs)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A != nullptr, HB200_ERROR_ARG, "null matrix");
   HB_CUDA(cudaDeviceSynchronize());
   if (diag_i) HB_CUDA(cudaMemcpy(diag_i, A->diag.i, sizeof(int) * ((size_t) A->num_rows + 1), cudaMemcpyDeviceToHost));
   if (diag_j && A->diag.nnz) HB_CUDA(cudaMemcpy(diag_j, A->diag.j, sizeof(int) * (size_t) A->diag.nnz, cudaMemcpyDeviceToHost));
   if (A->num_cols_offd > 0) {
      if (offd_i) HB_CUDA(cudaMemcpy(offd_i, A->offd.i, sizeof(int) * ((size_t) A->num_rows + 1), cudaMemcpyDeviceToHost));
      if (offd_j && A->offd.nnz) HB_CUDA(cudaMemcpy(offd_j, A->offd.j, sizeof(int) * (size_t) A->offd.nnz, cudaMemcpyDeviceToHost));
      if (col_map_offd) HB_CUDA(cudaMemcpy(col_map_offd, A->d_col_map_offd, sizeof(int64_t) * (size_t) A->num_cols_offd, cudaMemcpyDeviceToHost));
   }
   const CommPkgD &pk = A->pkg;
   if (send_map_starts) memcpy(send_map_starts, pk.send_map_starts.data(), sizeof(int) * pk.send_map_starts.size());
   if (recv_vec_starts) memcpy(recv_vec_starts, pk.recv_vec_starts.data(), sizeof(int) * pk.recv_vec_starts.size());
   if (send_procs && pk.num_sends) memcpy(send_procs, pk.send_procs.data(), sizeof(int) * pk.num_sends);
   if (recv_procs && pk.num_recvs) memcpy(recv_procs, pk.recv_procs.data(), sizeof(int) * pk.num_recvs);
   if (send_map_elmts && pk.d_send_map_elmts) {
      HB_CUDA(cudaMemcpy(send_map_elmts, pk.d_send_map_elmts, sizeof(int) * pk.send_map_elmts.size(), cudaMemcpyDeviceToHost));
   }
   return 0;
}

int hb200_parcsr_format_info(const hb200_parcsr *A, long long *info10)
{
   HB_REQUIRE(A && info10, HB200_ERROR_ARG, "null argument");
   for (int k = 0; k < 10; k++) info10[k] = 0;
   info10[0] = A->diag.has_sell ? 1 : 0;
   if (A->diag.has_sell) {
      long long total = 0;
      HB_CUDA(cudaMemcpy(&total, A->diag.sell_ptr + A->diag.sell_nslices, sizeof(long long), cudaMemcpyDeviceToHost));
      info10[1] = total;
      info10[2] = A->diag.sell_vidx ? 2 : 9;
      info10[3] = A->diag.sell_nv;
   }
   info10[4] = (A->diag.has_pat ? 1 : 0) | (A->diag.has_box ? 2 : 0) | (A->diag.box_uniform ? 4 : 0) | (A->diag.box_geo ? 8 : 0);
   info10[5] = A->diag.pat_npat;
   info10[6] = A->diag.pat_nent;
   info10[7] = A->diag.kind;
   info10[8] = A->diag.pat_nirr;
   info10[9] = A->diag.pat_irr_nnz;
   return 0;
}

int hb200_parcsr_set_spmv_kernel(hb200_parcsr *A, int kind, int lanes_per_row)
{
   HB_REQUIRE(A != nullptr, HB200_ERROR_ARG, "null matrix");
   HB_REQUIRE(kind >= 0 && kind <= 9 && !(kind >= 3 && kind <= 5), HB200_ERROR_ARG, "kind must be 0, 1, 2, 6, 7, 8 or 9");
   HB_REQUIRE(lanes_per_row == 0 || (lanes_per_row <= 32 && (lanes_per_row & (lanes_per_row - 1)) == 0),
              HB200_ERROR_ARG, "lanes_per_row must be 0 or a power of two <= 32");
   HB_CHECK(dcsr_ensure_formats(A->diag, kind));
   dcsr_choose_kernel(A->diag, kind, lanes_per_row);
   if (A->has_T) {
      HB_CHECK(dcsr_ensure_formats(A->diagT, kind));
      dcsr_choose_kernel(A->diagT, kind, lanes_per_row);
   }
   return 0;
}

int hb200_parcsr_matvec(hb200_parcsr *A, double alpha, const double *x, double beta,
                        const double *b, double *y)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A != nullptr, HB200_ERROR_ARG, "null matrix");
   // ranks may own zero rows / columns of a coarse level: only non-empty vectors must exist
   HB_REQUIRE((x || A->num_cols == 0) && (y || A->num_rows == 0) && (b || beta == 0.0 || A->num_rows == 0),
              HB200_ERROR_ARG, "null vector argument");
   HB_REQUIRE(x != y || x == nullptr, HB200_ERROR_ARG, "x must not alias y");
   return parcsr_matvec(A, alpha, x, beta, b, y);
}

int hb200_parcsr_matvecT(hb200_parcsr *A, double alpha, const double *x, double beta, double *y)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A != nullptr, HB200_ERROR_ARG, "null matrix");
   HB_REQUIRE((x || A->num_rows == 0) && (y || A->num_cols == 0), HB200_ERROR_ARG, "null vector argument");
   return parcsr_matvecT(A, alpha, x, beta, y);
}

int hb200_parcsr_matvec_host(hb200_parcsr *A, double alpha, const double *x_host, double beta,
                             double *y_host)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A && (x_host || A->num_cols == 0) && (y_host || A->num_rows == 0), HB200_ERROR_ARG, "null argument");
   Ctx &c = ctx();
   double *dx = nullptr, *dy = nullptr;
   HB_CUDA(cudaMalloc(&dx, sizeof(double) * (size_t) (A->num_cols ? A->num_cols : 1)));
   HB_CUDA(cudaMalloc(&dy, sizeof(double) * (size_t) (A->num_rows ? A->num_rows : 1)));
   HB_CUDA(cudaMemcpyAsync(dx, x_host, sizeof(double) * (size_t) A->num_cols, cudaMemcpyHostToDevice, c.s_comp));
   if (beta != 0.0) {
      HB_CUDA(cudaMemcpyAsync(dy, y_host, sizeof(double) * (size_t) A->num_rows, cudaMemcpyHostToDevice, c.s_comp));
   }
   int f = parcsr_matvec(A, alpha, dx, beta, dy, dy);
   if (!f) {
      cudaError_t e = cudaMemcpyAsync(y_host, dy, sizeof(double) * (size_t) A->num_rows, cudaMemcpyDeviceToHost, c.s_comp);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c.s_comp);
      if (e != cudaSuccess) f = set_error(HB200_ERROR_GENERIC, "matvec_host copy back: %s", cudaGetErrorString(e));
   }
   cudaFree(dx);
   cudaFree(dy);
   return f;
}

}  // extern "C"
