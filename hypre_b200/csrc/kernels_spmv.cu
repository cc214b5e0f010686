// kernels_spmv.cu — CSR SpMV kernels for sm_100a with fused epilogues.
//
// Replaces the hot loop of hypre_CSRMatrixMatvecOutOfPlaceHost (src/seq_mv/csr_matvec.c:683-
// 845) and its device twins (cuSPARSE CSR_ALG2 / hypreGPUKernel_CSRMatvecShuffle,
// src/seq_mv/csr_spmv_device.c:149-262).  Two kernels, chosen per matrix (per AMG level):
//
//  * spmv_stream  — nnz-balanced ("merge-style"): a CTA owns a contiguous run of rows
//    holding <= CAP nonzeros (partition built once at upload).  Phase 1 streams col_ind /
//    values with 128-bit loads (int4 + 2x double2 per thread, perfectly coalesced, aligned
//    down to a 4-entry boundary), gathers x through the read-only L1 path and parks the
//    products in shared memory.  Phase 2 sums each row from shared memory with L lanes per
//    row (L = 1: one thread adds the products in CSR order with separate mul/add — exactly
//    the operation order of the reference's sequential loop, so the result is bit-identical
//    to the 1-thread CPU reference) and applies the epilogue.
//  * spmv_vector  — K lanes per row, warp-shuffle reduction; used for offd blocks through the
//    compressed non-empty-row list (hypre_CSRMatrixRownnz) and as a cross-check kernel.
//
// HBM traffic model: 12 B/nnz (4 B index + 8 B value) + 4 B/row row pointer +
// the epilogue's vectors; x is gathered through L1/L2 (a 27-point row block touches 9 short
// x segments, reuse factor ~27).
#include "hb_internal.cuh"
#include <thread>
#include <chrono>
#include "hb_epilogue.cuh"

namespace hb {

// ---------------------------------------------------------------------------------------
// stream kernel (nnz-balanced CTA windows, products parked in shared memory, then ONE thread per
// row adds them in CSR order with separate multiply / add: bit-identical to the 1-thread CPU
// reference).  It lost to the sub-warp vector kernel on every level measured (0.59 against 0.87 of
// the HBM peak on A_0: the shared-memory round trip is L1-wavefront bound, profiles/r1_level_sweep_*.txt)
// and is kept, in this one variant, as the order-preserving cross-check of the parity tests.
// Lane l of a warp takes nonzero p+l, so the x gathers of one load instruction fall on neighbouring
// columns; col_ind / values are read with scalar, perfectly coalesced loads, 4 in flight per thread.
// ---------------------------------------------------------------------------------------
constexpr int kStreamThreads = 256;
constexpr int kStreamCap = 2048;

template <int EPI, int L, int CAP, int NT>
__global__ void __launch_bounds__(NT)
spmv_stream_lc(const int *__restrict__ rowptr, const int *__restrict__ colind,
               const double *__restrict__ val, const double *__restrict__ x,
               const int *__restrict__ blk_row, EpiArgs ea)
{
   __shared__ double prod[CAP];
   const int tid = threadIdx.x;
   const int r0  = blk_row[blockIdx.x];
   const int r1  = blk_row[blockIdx.x + 1];
   const int p0  = rowptr[r0];
   const int p1  = rowptr[r1];
   const int skip = (EPI == EPI_JACOBI_CORE) ? ea.skip_diag : 0;
   const int nnzb = p1 - p0;

   if (nnzb > CAP) {
      double s = 0.0;
      for (int p = p0 + skip + tid; p < p1; p += NT) s += val[p] * __ldg(x + colind[p]);
      __shared__ double red[NT / 32];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if ((tid & 31) == 0) red[tid >> 5] = s;
      __syncthreads();
      if (tid == 0) {
         double t = 0.0;
#pragma unroll
         for (int w = 0; w < NT / 32; w++) t += red[w];
         epi_apply<EPI>(ea, r0, t, epi_needs_diag<EPI>() ? val[p0] : 0.0);
      }
      return;
   }

   // ---- phase 1
   constexpr int U = 4;
   int k = tid;
   for (; k + (U - 1) * NT < nnzb; k += U * NT) {
      int c[U];
      double v[U], xv[U];
#pragma unroll
      for (int u = 0; u < U; u++) c[u] = colind[p0 + k + u * NT];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = val[p0 + k + u * NT];
#pragma unroll
      for (int u = 0; u < U; u++) xv[u] = __ldg(x + c[u]);
#pragma unroll
      for (int u = 0; u < U; u++) prod[k + u * NT] = __dmul_rn(v[u], xv[u]);
   }
   for (; k < nnzb; k += NT) prod[k] = __dmul_rn(val[p0 + k], __ldg(x + colind[p0 + k]));
   __syncthreads();

   // ---- phase 2
   const int nrows = r1 - r0;
   if (L == 1) {
      for (int r = tid; r < nrows; r += NT) {
         const int row = r0 + r;
         const int a = rowptr[row] - p0, b = rowptr[row + 1] - p0;
         double s = 0.0;
         for (int q = a + skip; q < b; q++) s = __dadd_rn(s, prod[q]);
         epi_apply<EPI>(ea, row, s, epi_needs_diag<EPI>() ? val[rowptr[row]] : 0.0);
      }
   } else {
      const int lane = tid % L;
      const int grp  = tid / L;
      constexpr int NG = NT / L;
      const int trips = (nrows + NG - 1) / NG;
      for (int t = 0; t < trips; t++) {
         const int r = grp + t * NG;
         double s = 0.0;
         const int row = r0 + r;
         if (r < nrows) {
            const int a = rowptr[row] - p0, b = rowptr[row + 1] - p0;
            for (int q = a + skip + lane; q < b; q += L) s += prod[q];
         }
#pragma unroll
         for (int o = L / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, L);
         if (r < nrows && lane == 0) {
            epi_apply<EPI>(ea, row, s, epi_needs_diag<EPI>() ? val[rowptr[row]] : 0.0);
         }
      }
   }
}

// ---------------------------------------------------------------------------------------
// vector kernel: K lanes per row, optional compressed row list
// ---------------------------------------------------------------------------------------
constexpr int kVecThreads = 256;

// I16: column indices stored as 16-bit offsets from the row (square blocks whose rows stay within
// +-32767 of the diagonal: every AMG level of a grid problem) -> 10 B per nonzero instead of 12
template <int EPI, int K, int U, bool I16 = false>
__global__ void __launch_bounds__(kVecThreads)
spmv_vector(int nlist, const int *__restrict__ rowlist, const int *__restrict__ rowptr,
            const int *__restrict__ colind, const double *__restrict__ val,
            const double *__restrict__ x, EpiArgs ea)
{
   const short *__restrict__ col16 = reinterpret_cast<const short *>(colind);
   const int gtid = blockIdx.x * kVecThreads + threadIdx.x;
   const int idx  = gtid / K;
   const int lane = threadIdx.x % K;
   const int skip = (EPI == EPI_JACOBI_CORE) ? ea.skip_diag : 0;
   double s = 0.0;
   int row = 0, p0 = 0;
   const bool active = idx < nlist;
   if (active) {
      row = rowlist ? rowlist[idx] : idx;
      p0 = rowptr[row];
      const int p1 = rowptr[row + 1];
      int p = p0 + skip + lane;
      if (I16) {
         const double *xr = x + row;
         for (; p < p1; p += K) s += val[p] * __ldg(xr + col16[p]);
      } else {
         for (; p < p1; p += K) s += val[p] * __ldg(x + colind[p]);
      }
   }
#pragma unroll
   for (int o = K / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, K);
   if (active && lane == 0) {
      epi_apply<EPI>(ea, row, s, epi_needs_diag<EPI>() ? val[p0] : 0.0);
   }
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------

template <int EPI>
static int launch_stream(const DCsr &M, const double *x, const EpiArgs &ea, cudaStream_t st)
{
   HB_LAUNCH((spmv_stream_lc<EPI, 1, kStreamCap, kStreamThreads>), M.nblks, kStreamThreads, 0, st,
             M.i, M.j, M.a, x, M.blk_row, ea);
   HB_LAUNCH_CHECK();
   return 0;
}

template <int EPI, int K>
static int launch_vector_K(const DCsr &M, const double *x, const EpiArgs &ea, const int *rowlist, int nlist,
                           cudaStream_t st)
{
   const long long threads = (long long) nlist * K;
   const int grid = (int) ((threads + kVecThreads - 1) / kVecThreads);
   const int *rl = rowlist;
   if (!rowlist && M.kind == SPMV_VECTOR16 && M.j16) {
      HB_LAUNCH((spmv_vector<EPI, K, 1, true>), grid, kVecThreads, 0, st, nlist, rl, M.i,
                reinterpret_cast<const int *>(M.j16), M.a, x, ea);
      HB_LAUNCH_CHECK();
      return 0;
   }
   HB_LAUNCH((spmv_vector<EPI, K, 1>), grid, kVecThreads, 0, st, nlist, rl, M.i, M.j, M.a, x, ea);
   HB_LAUNCH_CHECK();
   return 0;
}

template <int EPI>
static int launch_vector(const DCsr &M, const double *x, const EpiArgs &ea, const int *rowlist, int nlist,
                         int lanes, cudaStream_t st)
{
   switch (lanes) {
      case 1:  return launch_vector_K<EPI, 1>(M, x, ea, rowlist, nlist, st);
      case 2:  return launch_vector_K<EPI, 2>(M, x, ea, rowlist, nlist, st);
      case 4:  return launch_vector_K<EPI, 4>(M, x, ea, rowlist, nlist, st);
      case 8:  return launch_vector_K<EPI, 8>(M, x, ea, rowlist, nlist, st);
      case 16: return launch_vector_K<EPI, 16>(M, x, ea, rowlist, nlist, st);
      default: return launch_vector_K<EPI, 32>(M, x, ea, rowlist, nlist, st);
   }
}

static int vector_lanes_for(double avg, long long nrows = 0)
{
   // measured on B200 over the levels of 27-pt and 7-pt hierarchies (profiles/r1_level_sweep_*.txt):
   // the fastest lane count keeps ~6-14 nonzeros per lane; with many rows 16 lanes beat a full
   // warp even on 200-entry rows
   if (avg >= 150) return nrows >= 100000 ? 16 : 32;
   if (avg >= 80)  return 16;
   if (avg >= 36)  return 8;
   if (avg >= 10)  return 2;
   return 1;
}

// A launch with fewer threads than the GPU holds is latency-bound on the per-lane serial chain of
// dependent gathers (offd blocks, coarse levels): spread each row over more lanes until the grid
// covers the SMs, as long as every lane still has an entry to work on.  HB200_NO_WIDEN=1 disables.
static int widen_lanes(int lanes, long long nlist, double avg)
{
   static const bool off = env_flag("HB200_NO_WIDEN", false);
   if (off) return lanes;
   while (lanes < 32 && nlist * lanes < 148LL * 1024 && lanes < avg) lanes *= 2;
   return lanes;
}

template <int EPI>
static int spmv_dispatch(const DCsr &M, const double *x, const EpiArgs &ea, bool use_rownnz,
                         cudaStream_t st)
{
   const int nlist = use_rownnz ? M.num_rownnz : M.nrows;
   if (nlist == 0) return 0;
   if (!use_rownnz && (M.kind == SPMV_PAT || M.kind == SPMV_BOX) && M.has_pat) {
      // the fused l1-Jacobi sweep: the generic kernel's 4 rows per thread sit 256 rows apart, which makes them
      // y-neighbours when the plane stride divides 256 — there it is the faster of the two (B200, 256^3: 0.238
      // against 0.276 ms; the SpMV itself: 0.184 against 0.150 ms; profiles/r2_session_log.md)
      static const bool box_jacobi = env_flag("HB200_BOX_JACOBI", false);
      const bool pat_aligned = (EPI == EPI_JACOBI7) && M.box_sy > 0 && (256 % M.box_sy) == 0 && !box_jacobi;
      if (M.kind == SPMV_BOX && M.has_box && spmv_box_supports(EPI) && !pat_aligned) HB_CHECK(spmv_box_launch(M, x, EPI, ea, st));
      else HB_CHECK(spmv_pat_launch(M, x, EPI, ea, st));
      if (M.pat_nirr == 0) return 0;
      // rows outside the pattern table: CSR sweep over the row list (disjoint rows, same epilogue)
      const double avg = (double) M.pat_irr_nnz / (double) M.pat_nirr;
      return launch_vector<EPI>(M, x, ea, M.pat_irr, M.pat_nirr, widen_lanes(vector_lanes_for(avg), M.pat_nirr, avg), st);
   }
   if (!use_rownnz && M.kind == SPMV_SELL && M.has_sell) return spmv_sell_launch(M, x, EPI, ea, st);
   if (!use_rownnz && M.kind == SPMV_STREAM && M.nblks > 0) return launch_stream<EPI>(M, x, ea, st);
   const bool vec = (M.kind == SPMV_VECTOR || M.kind == SPMV_VECTOR16);
   int lanes = (vec && M.lanes > 0 && !use_rownnz) ? M.lanes : 0;
   if (lanes == 0) {
      const double avg = use_rownnz ? (double) M.nnz / (double) (M.num_rownnz ? M.num_rownnz : 1)
                                    : M.avg_row_nnz;
      lanes = widen_lanes(vector_lanes_for(avg), nlist, avg);
   }
   return launch_vector<EPI>(M, x, ea, use_rownnz ? M.rownnz : (const int *) nullptr, nlist, lanes, st);
}

int spmv_launch(const DCsr &M, const double *x, int epi_kind, const EpiArgs &ea, bool use_rownnz,
                cudaStream_t st)
{
   switch (epi_kind) {
      case EPI_AXPBY:           return spmv_dispatch<EPI_AXPBY>(M, x, ea, use_rownnz, st);
      case EPI_ACC:             return spmv_dispatch<EPI_ACC>(M, x, ea, use_rownnz, st);
      case EPI_JACOBI7:         return spmv_dispatch<EPI_JACOBI7>(M, x, ea, use_rownnz, st);
      case EPI_JACOBI7_ACC:     return spmv_dispatch<EPI_JACOBI7_ACC>(M, x, ea, use_rownnz, st);
      case EPI_JACOBI_CORE:     return spmv_dispatch<EPI_JACOBI_CORE>(M, x, ea, use_rownnz, st);
      case EPI_JACOBI_CORE_ACC: return spmv_dispatch<EPI_JACOBI_CORE_ACC>(M, x, ea, use_rownnz, st);
      case EPI_CHEBY_FIRST:     return spmv_dispatch<EPI_CHEBY_FIRST>(M, x, ea, use_rownnz, st);
      case EPI_CHEBY_STEP:      return spmv_dispatch<EPI_CHEBY_STEP>(M, x, ea, use_rownnz, st);
      default: return set_error(HB200_ERROR_ARG, "spmv_launch: unknown epilogue %d", epi_kind);
   }
}

// ---------------------------------------------------------------------------------------
// DCsr management
// ---------------------------------------------------------------------------------------
int dcsr_build_partition(DCsr &M, const int *hi)
{
   // greedy nnz-balanced partition: consecutive rows with (nnz + 3) <= kStreamCap and at most
   // kStreamThreads * 4 rows; a row that alone exceeds the window gets its own block.
   std::vector<int> blk;
   blk.reserve((size_t) (M.nnz / (kStreamCap / 2)) + 16);
   const int n = M.nrows;
   const int max_rows = kStreamThreads * 4;
   int r = 0;
   blk.push_back(0);
   while (r < n) {
      const int p0 = hi[r];
      int e = r + 1;
      // grow while the window fits
      while (e < n && (hi[e + 1] - p0 + 3) <= kStreamCap && (e - r) < max_rows) e++;
      blk.push_back(e);
      r = e;
   }
   M.nblks = (int) blk.size() - 1;
   if (M.blk_row) { cudaFree(M.blk_row); M.blk_row = nullptr; }
   if (M.nblks > 0) {
      HB_CUDA(cudaMalloc(&M.blk_row, sizeof(int) * blk.size()));
      HB_CUDA(cudaMemcpy(M.blk_row, blk.data(), sizeof(int) * blk.size(), cudaMemcpyHostToDevice));
   }
   return 0;
}

void dcsr_choose_kernel(DCsr &M, int kind, int lanes)
{
   // the sub-warp vector kernel beat the shared-memory stream kernel on every level measured
   // (profiles/r1_level_sweep.md): the stream kernel is L1-wavefront bound by its smem round trip
   const int csr = M.j16 ? SPMV_VECTOR16 : SPMV_VECTOR;
   if (kind == SPMV_AUTO) kind = M.has_box ? SPMV_BOX : M.has_pat ? SPMV_PAT : M.has_sell ? SPMV_SELL : csr;
   if (kind == SPMV_BOX && !M.has_box) kind = SPMV_PAT;
   if (kind == SPMV_PAT && !M.has_pat) kind = M.has_sell ? SPMV_SELL : csr;
   if (kind == SPMV_SELL && !M.has_sell) kind = csr;
   if (kind == SPMV_VECTOR16 && !M.j16) kind = SPMV_VECTOR;
   M.kind = kind;
   if (lanes > 0) { M.lanes = lanes; return; }
   if (kind == SPMV_STREAM) {
      M.lanes = 1;   // the one variant kept: one thread per row, CSR order
   } else {
      M.lanes = widen_lanes(vector_lanes_for(M.avg_row_nnz, M.nrows), M.nrows, M.avg_row_nnz);
   }
}

// 16-bit column offsets from the row, when every entry of a square block stays within +-32767 of the
// diagonal (the SPMV_VECTOR16 kernel: 10 instead of 12 bytes per nonzero)
static int dcsr_build_j16(DCsr &M, const int *hi, const int *hj)
{
   const int nrows = M.nrows;
   if (M.j16 || !(nrows >= 1024 && nrows == M.ncols && M.nnz >= 2LL * nrows) || env_flag("HB200_NO_CSR16", false)) return 0;
   bool fits = true;
   for (int r = 0; r < nrows && fits; r++) {
      for (int q = hi[r]; q < hi[r + 1]; q++) {
         const int d = hj[q] - r;
         if (d < -32767 || d > 32767) { fits = false; break; }
      }
   }
   if (!fits) return 0;
   std::vector<short> j16((size_t) M.nnz + 8, 0);
   for (int r = 0; r < nrows; r++) {
      for (int q = hi[r]; q < hi[r + 1]; q++) j16[(size_t) q] = (short) (hj[q] - r);
   }
   HB_CUDA(cudaMalloc(&M.j16, sizeof(short) * j16.size()));
   HB_CUDA(cudaMemcpy(M.j16, j16.data(), sizeof(short) * j16.size(), cudaMemcpyHostToDevice));
   return 0;
}

// The formats the automatic choice cannot pick are not built at upload: a block stored in the
// row-pattern format never runs the packed-SELL kernel, and one whose every row is in the pattern
// table never runs the CSR kernels either.  A caller that forces such a kernel
// (hb200_parcsr_set_spmv_kernel: the level sweeps, the parity tests) gets the format built then,
// from the device copy of the CSR arrays.  HB200_EAGER_FORMATS=1 builds everything at upload.
int dcsr_ensure_formats(DCsr &M, int kind)
{
   const bool want_j16 = (kind == SPMV_VECTOR16) && M.defer_j16;
   const bool want_sell = (kind == SPMV_SELL) && M.defer_sell;
   const bool want_part = (kind == SPMV_STREAM) && M.nblks == 0 && M.nrows > 0;
   if (!want_j16 && !want_sell && !want_part) return 0;
   std::vector<int> hi((size_t) M.nrows + 1), hj;
   std::vector<double> ha;
   HB_CUDA(cudaMemcpy(hi.data(), M.i, sizeof(int) * hi.size(), cudaMemcpyDeviceToHost));
   if (want_part) HB_CHECK(dcsr_build_partition(M, hi.data()));
   if (!want_j16 && !want_sell) { HB_CUDA(cudaDeviceSynchronize()); return 0; }
   hj.resize((size_t) M.nnz);
   HB_CUDA(cudaMemcpy(hj.data(), M.j, sizeof(int) * hj.size(), cudaMemcpyDeviceToHost));
   if (want_j16) {
      M.defer_j16 = false;
      HB_CHECK(dcsr_build_j16(M, hi.data(), hj.data()));
   }
   if (want_sell) {
      M.defer_sell = false;
      ha.resize((size_t) M.nnz);
      HB_CUDA(cudaMemcpy(ha.data(), M.a, sizeof(double) * ha.size(), cudaMemcpyDeviceToHost));
      HB_CHECK(dcsr_build_sell(M, hi.data(), hj.data(), ha.data()));
   }
   // the copies above went through the legacy default stream, the kernels run on non-blocking streams:
   // small pageable H2D copies may still be in flight when cudaMemcpy returns
   HB_CUDA(cudaDeviceSynchronize());
   return 0;
}

static double upload_now()
{
   return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// second half of an upload: everything derived from the CSR arrays (non-empty row list, row statistics,
// the structured formats, the kernel choice).  M.i / M.j / M.a are already on the device; hi / hj / ha
// are the same arrays in host memory.
int dcsr_analyze(DCsr &M, const int *hi, const int *hj, const double *ha)
{
   const int nrows = M.nrows, ncols = M.ncols;
   const double t_copied = upload_now();
   // non-empty row list + row statistics
   std::vector<int> rn;
   int mx = 0;
   for (int r = 0; r < nrows; r++) {
      const int len = hi[r + 1] - hi[r];
      if (len > 0) rn.push_back(r);
      if (len > mx) mx = len;
   }
   M.max_row_nnz = mx;
   M.avg_row_nnz = nrows > 0 ? (double) M.nnz / (double) nrows : 0.0;
   M.num_rownnz = (int) rn.size();
   if (M.num_rownnz > 0) {
      HB_CUDA(cudaMalloc(&M.rownnz, sizeof(int) * rn.size()));
      HB_CUDA(cudaMemcpy(M.rownnz, rn.data(), sizeof(int) * rn.size(), cudaMemcpyHostToDevice));
   }
   // (the nnz-balanced partition of the stream kernel is built when that kernel is asked for:
   //  dcsr_ensure_formats)
   const double t_rows = upload_now();
   if (nrows > 0) HB_CHECK(dcsr_build_pat(M, hi, hj, ha));
   const double t_pat = upload_now();
   const bool eager = env_flag("HB200_EAGER_FORMATS", false);
   const bool square = nrows > 0 && nrows == ncols;
   // 16-bit offsets: the CSR kernel of blocks outside the row-pattern format and of the rows a pattern
   // table leaves out
   if (square && (eager || !M.has_pat || M.pat_nirr > 0)) HB_CHECK(dcsr_build_j16(M, hi, hj)); else M.defer_j16 = square;
   const double t_csr = upload_now();
   if (square && (eager || !M.has_pat)) HB_CHECK(dcsr_build_sell(M, hi, hj, ha)); else M.defer_sell = square;   // square (A_l) blocks only
   const double t_sell = upload_now();
   if (eager && nrows > 0) HB_CHECK(dcsr_build_partition(M, hi));
   dcsr_choose_kernel(M, SPMV_AUTO, 0);
   // uploads use the legacy default stream, kernels the non-blocking compute stream (see dcsr_ensure_formats)
   HB_CUDA(cudaDeviceSynchronize());
   if (nrows >= 1024) {
      HB_TRACE("analysis of a %d x %d block, %lld nnz: row lists %.3f s, row patterns %.3f s, 16-bit offsets %.3f s%s, "
               "packed SELL %.3f s%s -> kernel kind %d", nrows, ncols, M.nnz, t_rows - t_copied,
               t_pat - t_rows, t_csr - t_pat, M.defer_j16 ? " (deferred)" : "", t_sell - t_csr,
               M.defer_sell ? " (deferred)" : "", M.kind);
   }
   return 0;
}

// pageable -> pinned copy of one staging chunk.  One core moves ~12 GB/s (B200 box, r2 trace: A_0 in 0.43 s), a fifth
// of what the PCIe 5 link takes: the chunk is cut across a few threads (HB200_UPLOAD_THREADS, default 4; this container:
// 5.3 -> 16.8 GB/s from 1 to 4 threads)
static void stage_fill(char *dst, const char *src, size_t len)
{
   static int nthreads = 0;
   if (nthreads == 0) {
      const char *e = getenv("HB200_UPLOAD_THREADS");
      nthreads = e ? atoi(e) : 4;
      if (nthreads < 1) nthreads = 1;
      if (nthreads > 16) nthreads = 16;
   }
   if (nthreads == 1 || len < ((size_t) 4 << 20)) { memcpy(dst, src, len); return; }
   const size_t part = (len / (size_t) nthreads + 63) & ~(size_t) 63;
   std::vector<std::thread> th;
   for (int t = 1; t < nthreads; t++) {
      const size_t b = (size_t) t * part;
      if (b >= len) break;
      const size_t l = (b + part <= len) ? part : len - b;
      th.emplace_back([=]() { memcpy(dst + b, src + b, l); });
   }
   memcpy(dst, src, part < len ? part : len);
   for (auto &x : th) x.join();
}

// host -> device copy of one big array through two pinned staging buffers: the CPU fills one while
// the DMA engine drains the other (a pageable cudaMemcpy stages through one small driver buffer)
static int upload_array(void *dst, const void *src, size_t bytes)
{
   static char *stage[2] = {nullptr, nullptr};
   static cudaEvent_t done[2];
   static cudaStream_t st = nullptr;
   constexpr size_t kChunk = (size_t) 32 << 20;
   if (bytes == 0) return 0;
   if (bytes < ((size_t) 1 << 20) || env_flag("HB200_PAGEABLE_UPLOAD", false)) {
      HB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
      return 0;
   }
   if (!st) {
      HB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      for (int k = 0; k < 2; k++) {
         HB_CUDA(cudaMallocHost((void **) &stage[k], kChunk));
         HB_CUDA(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
      }
   }
   int k = 0;
   bool used[2] = {false, false};
   for (size_t off = 0; off < bytes; off += kChunk, k ^= 1) {
      const size_t len = bytes - off < kChunk ? bytes - off : kChunk;
      if (used[k]) HB_CUDA(cudaEventSynchronize(done[k]));
      stage_fill(stage[k], (const char *) src + off, len);
      HB_CUDA(cudaMemcpyAsync((char *) dst + off, stage[k], len, cudaMemcpyHostToDevice, st));
      HB_CUDA(cudaEventRecord(done[k], st));
      used[k] = true;
   }
   HB_CUDA(cudaStreamSynchronize(st));
   return 0;
}

int dcsr_upload(DCsr &M, int nrows, int ncols, const int *hi, const int *hj, const double *ha)
{
   const double t_start = upload_now();
   M.nrows = nrows;
   M.ncols = ncols;
   M.nnz = nrows > 0 ? hi[nrows] : 0;
   const size_t nnz_pad = (size_t) M.nnz + 8;   // the stream kernel reads up to 3 entries past the end
   HB_CUDA(cudaMalloc(&M.i, sizeof(int) * ((size_t) nrows + 1)));
   HB_CUDA(cudaMalloc(&M.j, sizeof(int) * nnz_pad));
   HB_CUDA(cudaMalloc(&M.a, sizeof(double) * nnz_pad));
   HB_CUDA(cudaMemset(M.j + M.nnz, 0, sizeof(int) * (nnz_pad - (size_t) M.nnz)));       // the pad only
   HB_CUDA(cudaMemset(M.a + M.nnz, 0, sizeof(double) * (nnz_pad - (size_t) M.nnz)));
   if (nrows > 0) {
      HB_CHECK(upload_array(M.i, hi, sizeof(int) * ((size_t) nrows + 1)));
   } else {
      int z = 0;
      HB_CUDA(cudaMemcpy(M.i, &z, sizeof(int), cudaMemcpyHostToDevice));
   }
   if (M.nnz > 0) {
      HB_CHECK(upload_array(M.j, hj, sizeof(int) * (size_t) M.nnz));
      HB_CHECK(upload_array(M.a, ha, sizeof(double) * (size_t) M.nnz));
   }
   if (nrows >= 1024) HB_TRACE("upload %d x %d block, %lld nnz: CSR copy %.3f s", nrows, ncols, M.nnz, upload_now() - t_start);
   return dcsr_analyze(M, hi, hj, ha);
}

int dcsr_free(DCsr &M)
{
   if (M.i) cudaFree(M.i);
   if (M.j) cudaFree(M.j);
   if (M.a) cudaFree(M.a);
   if (M.rownnz) cudaFree(M.rownnz);
   if (M.blk_row) cudaFree(M.blk_row);
   if (M.j16) cudaFree(M.j16);
   dcsr_free_sell(M);
   dcsr_free_pat(M);
   M = DCsr();
   return 0;
}

void host_csr_transpose(int nrows, int ncols, const int *ai, const int *aj, const double *aa,
                        std::vector<int> &ti, std::vector<int> &tj, std::vector<double> &ta)
{
   const int nnz = nrows > 0 ? ai[nrows] : 0;
   ti.assign((size_t) ncols + 1, 0);
   tj.resize(nnz);
   ta.resize(nnz);
   for (int p = 0; p < nnz; p++) ti[aj[p] + 1]++;
   for (int c = 0; c < ncols; c++) ti[c + 1] += ti[c];
   std::vector<int> next(ti.begin(), ti.end() - 1);
   for (int r = 0; r < nrows; r++) {
      for (int p = ai[r]; p < ai[r + 1]; p++) {
         const int q = next[aj[p]]++;
         tj[q] = r;
         ta[q] = aa[p];
      }
   }
}

}  // namespace hb

// host half of the restriction path, reachable without a GPU (CPU tests of the stored transpose)
extern "C" int hb200_host_csr_transpose(int num_rows, int num_cols, const int *row_ptr, const int *col_ind,
                                        const double *values, int *t_row_ptr, int *t_col_ind, double *t_values)
{
   using namespace hb;
   HB_REQUIRE(num_rows >= 0 && num_cols >= 0 && row_ptr && t_row_ptr, HB200_ERROR_ARG,
              "hb200_host_csr_transpose: null argument");
   const int nnz = num_rows > 0 ? row_ptr[num_rows] : 0;
   HB_REQUIRE(nnz == 0 || (col_ind && values && t_col_ind && t_values), HB200_ERROR_ARG,
              "hb200_host_csr_transpose: null matrix arrays");
   std::vector<int> ti, tj;
   std::vector<double> ta;
   host_csr_transpose(num_rows, num_cols, row_ptr, col_ind, values, ti, tj, ta);
   memcpy(t_row_ptr, ti.data(), sizeof(int) * ti.size());
   if (nnz) {
      memcpy(t_col_ind, tj.data(), sizeof(int) * (size_t) nnz);
      memcpy(t_values, ta.data(), sizeof(double) * (size_t) nnz);
   }
   return 0;
}
