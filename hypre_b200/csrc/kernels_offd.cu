// kernels_offd.cu — the offd pass of a ParCSR operation fused with the halo wait (peer-put halo,
// opt-in HB200_FUSE_WAIT=1).
//
// Unfused (parcsr_peer.cu): halo_wait_kernel polls the arrival flags, copies the receive buffer
// into x_ext, acks the senders; then spmv_vector<EPI_*_ACC> over the boundary rows reads x_ext.
// Here one kernel does both: every block polls the flags, the boundary rows read the NVLink
// receive buffer of this exchange directly, and the last block to finish acks the senders and
// advances the epoch.  One launch and one copy fewer per exchange (~1 000 exchanges per solve).
// The row arithmetic is that of spmv_vector (kernels_spmv.cu): K lanes per row over the row list
// of the offd block, shuffle reduction, the *_ACC epilogues.
#include "hb_internal.cuh"
#include "hb_epilogue.cuh"
#include "hb_peer.cuh"

namespace hb {

constexpr int kOffdThreads = 256;

template <int EPI, int K>
__global__ void __launch_bounds__(kOffdThreads)
spmv_offd_wait(int nlist, const int *__restrict__ rowlist, const int *__restrict__ rowptr,
               const int *__restrict__ colind, const double *__restrict__ val, PeerWaitArgs w, EpiArgs ea)
{
   __shared__ bool is_last;
   // ---- wait for this exchange's data (same protocol as halo_wait_kernel)
   const unsigned long long epoch = w.epoch_ctr[1] + 1;
   const int par = (int) (epoch & 1ull);
   SpinGuard guard;
   guard.err = w.err; guard.timeout_ns = w.timeout_ns;
   for (int j = threadIdx.x; j < w.n_in; j += kOffdThreads) spin_until_ge(w.flags + par * w.n_in + j, epoch, guard, 2, j);
   __syncthreads();
   const double *x = par ? w.buf1 : w.buf0;
   // ---- the boundary rows
   const int gtid = blockIdx.x * kOffdThreads + threadIdx.x;
   const int idx  = gtid / K;
   const int lane = threadIdx.x % K;
   double s = 0.0;
   int row = 0, p0 = 0;
   const bool active = idx < nlist;
   if (active) {
      row = rowlist[idx];
      p0 = rowptr[row];
      const int p1 = rowptr[row + 1];
      for (int p = p0 + lane; p < p1; p += K) s += val[p] * __ldcg(x + colind[p]);
   }
#pragma unroll
   for (int o = K / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, K);
   if (active && lane == 0) epi_apply<EPI>(ea, row, s, epi_needs_diag<EPI>() ? val[p0] : 0.0);
   // ---- everything of this exchange has been read: the last block tells the senders
   __syncthreads();
   if (threadIdx.x == 0) {
      if (gridDim.x == 1) {
         is_last = true;
      } else {
         __threadfence();
         const unsigned int t = atomicInc(w.ticket + 1, gridDim.x - 1);
         is_last = (t == gridDim.x - 1);
      }
   }
   __syncthreads();
   if (is_last) {
      for (int j = threadIdx.x; j < w.n_in; j += kOffdThreads) st_release_sys(w.in_ack[j], epoch);
      if (threadIdx.x == 0) w.epoch_ctr[1] = epoch;
   }
}

template <int EPI, int K>
static int offd_launch_K(const DCsr &M, const PeerWaitArgs &w, const EpiArgs &ea, cudaStream_t st)
{
   const long long threads = (long long) M.num_rownnz * K;
   int grid = (int) ((threads + kOffdThreads - 1) / kOffdThreads);
   if (grid < 1) grid = 1;
   HB_LAUNCH((spmv_offd_wait<EPI, K>), grid, kOffdThreads, 0, st, M.num_rownnz, M.rownnz, M.i, M.j, M.a, w, ea);
   HB_LAUNCH_CHECK();
   return 0;
}

template <int EPI>
static int offd_launch(const DCsr &M, const PeerWaitArgs &w, const EpiArgs &ea, cudaStream_t st)
{
   // lanes per row as the unfused offd pass picks them (latency-bound launches: spread the rows)
   const double avg = (double) M.nnz / (double) (M.num_rownnz ? M.num_rownnz : 1);
   int lanes = avg >= 150 ? 32 : avg >= 80 ? 16 : avg >= 36 ? 8 : avg >= 10 ? 2 : 1;
   while (lanes < 32 && (long long) M.num_rownnz * lanes < 148LL * 1024 && lanes < avg) lanes *= 2;
   switch (lanes) {
      case 1:  return offd_launch_K<EPI, 1>(M, w, ea, st);
      case 2:  return offd_launch_K<EPI, 2>(M, w, ea, st);
      case 4:  return offd_launch_K<EPI, 4>(M, w, ea, st);
      case 8:  return offd_launch_K<EPI, 8>(M, w, ea, st);
      case 16: return offd_launch_K<EPI, 16>(M, w, ea, st);
      default: return offd_launch_K<EPI, 32>(M, w, ea, st);
   }
}

// =======================================================================================
// parcsr_fused: ONE kernel = one complete ParCSR operation on a latency-bound level
// =======================================================================================
// On the coarse levels an operation y = epi(A x) is four dependent launches of a few microseconds each
// (halo put, diag pass, halo wait, offd pass) plus the flag round trip between them: about 40 such
// operations per V-cycle, and what keeps the N-GPU iteration above the single-GPU one.  Here the whole
// operation is one launch over peer memory:
//   put   : the first blocks gather x[send_map_elmts] and store it straight into the neighbours'
//           receive buffers over NVLink, fence, and the last of them raises the arrival flags;
//   diag  : every row group (K lanes per row) sums its diag-block row while the transfer is in flight;
//   wait  : only blocks that own a row with offd entries poll their arrival flags;
//   offd  : those rows add their offd part, read in place from the NVLink receive buffer;
//   epi   : one store per row — the value of the diag epilogue, then the offd accumulation applied to it
//           in registers (the same two roundings the two-kernel path makes through memory);
//   ack   : the last block to finish tells the senders that the buffer has been read.
// The protocol (double-buffered epochs, consumed-flags, bounded polling) is the one of parcsr_peer.cu;
// the blocks that put never wait for anything of the current exchange and come first in the grid, so a
// grid larger than the device still cannot starve the transfer.
template <int EPI>
__device__ __forceinline__ double fused_epi_value(const EpiArgs &ea, int row, double sd, double so, bool has_offd, double diag)
{
   if (EPI == EPI_AXPBY) {
      double v = (ea.beta == 0.0) ? ea.alpha * sd : epi_axpby_value(ea, __ldcs(ea.b + row), sd);
      if (has_offd) v += ea.alpha * so;
      return v;
   } else if (EPI == EPI_JACOBI7) {
      const double uo = ea.u[row];
      if (ea.cf == nullptr || __ldcs(ea.cf + row) == ea.relax_points) {
         const double d = __ldcs(ea.d + row);
         double v = epi_jacobi7_value(ea, uo, __ldcs(ea.b + row), d, sd);
         if (has_offd) v -= (ea.w * so) / d;
         return v;
      }
      return uo;
   } else {   // EPI_JACOBI_CORE
      const double uo = ea.u[row];
      const double di = ea.d ? ea.d[row] : diag;
      if ((ea.relax_points == 0 || ea.cf[row] == ea.relax_points) && di != 0.0) {
         const double res = ea.b[row] - sd;
         double v = ea.skip_diag ? uo * (1.0 - ea.w) + ea.w * res / di : uo + ea.w * res / di;
         if (has_offd) v -= ea.w * so / di;
         return v;
      }
      return uo;
   }
}

constexpr int kFusedThreads = 256;

// the put half of an exchange, run by the first `gput` blocks of a fused kernel
__device__ __forceinline__ void fused_put(const PeerFusedArgs &h, const double *__restrict__ src, int gput, const SpinGuard &guard, int *s_flag)
{
   const int tid = threadIdx.x;
   const unsigned long long epoch = h.w.epoch_ctr[0] + 1;
   const int par = (int) (epoch & 1ull);
   if (epoch > 2) {   // the receivers have read exchange (epoch - 2) out of this parity's buffers
      for (int i = tid; i < h.n_out; i += kFusedThreads) spin_until_ge(h.acks + i, epoch - 2, guard, 1, i);
      __syncthreads();
   }
   for (int k = blockIdx.x * kFusedThreads + tid; k < h.total_out; k += gput * kFusedThreads) {
      int lo = 0, hi = h.n_out - 1;
      while (lo < hi) {
         const int mid = (lo + hi + 1) >> 1;
         if (h.out_starts[mid] <= k) lo = mid; else hi = mid - 1;
      }
      h.dst2[par * h.n_out + lo][k - h.out_starts[lo]] = h.gather ? src[h.gather[k]] : src[k];
   }
   __syncthreads();
   if (tid == 0) {
      __threadfence_system();
      int last = 1;
      if (gput > 1) {
         const unsigned int t = atomicInc(h.w.ticket, (unsigned int) gput - 1);
         last = (t == (unsigned int) gput - 1);
         if (last) __threadfence_system();
      }
      *s_flag = last;
   }
   __syncthreads();
   if (*s_flag) {
      for (int i = tid; i < h.n_out; i += kFusedThreads) st_release_sys(h.flag2[par * h.n_out + i], epoch);
      if (tid == 0) h.w.epoch_ctr[0] = epoch;
   }
}

// the acknowledgement half: the last of the `nblocks` boundary blocks tells the senders
__device__ __forceinline__ void fused_ack(const PeerFusedArgs &h, unsigned long long epoch_in, int nblocks, int *s_flag)
{
   const int tid = threadIdx.x;
   __syncthreads();
   if (tid == 0) {
      int last = 1;
      if (nblocks > 1) {
         __threadfence();
         const unsigned int t = atomicInc(h.w.ticket + 1, (unsigned int) nblocks - 1);
         last = (t == (unsigned int) nblocks - 1);
      }
      *s_flag = last;
   }
   __syncthreads();
   if (*s_flag) {
      for (int j = tid; j < h.w.n_in; j += kFusedThreads) st_release_sys(h.w.in_ack[j], epoch_in);
      if (tid == 0) h.w.epoch_ctr[1] = epoch_in;
   }
}

// Block roles (by block index): [0, gput) put; [gput, gput + gint) the rows without offd entries — they
// never look at a flag; [gput + gint, grid) the boundary rows (the non-empty-row list of the offd block):
// diag part first, then the flags, then the offd part read in place from the NVLink receive buffer.  The
// boundary blocks come last in the grid: the interior is never held up behind a poll.
template <int EPI, int K, bool I16>
__global__ void __launch_bounds__(kFusedThreads)
parcsr_fused(int nrows, const int *__restrict__ di, const int *__restrict__ dj, const double *__restrict__ da,
             const int *__restrict__ oi, const int *__restrict__ oj, const double *__restrict__ oa,
             const int *__restrict__ bnd_rows, int nbnd, const double *__restrict__ x, PeerFusedArgs h,
             int gput, int gint, const double *__restrict__ xext_plain, EpiArgs ea)
{
   __shared__ int s_flag;
   const int tid = threadIdx.x;
   const int lane = tid % K;
   constexpr int G = kFusedThreads / K;
   SpinGuard guard;
   guard.err = h.w.err; guard.timeout_ns = h.w.timeout_ns;
   const short *__restrict__ dj16 = reinterpret_cast<const short *>(dj);
   const int skip = (EPI == EPI_JACOBI_CORE) ? ea.skip_diag : 0;
   const int b = (int) blockIdx.x;
   if (b < gput) { fused_put(h, x, gput, guard, &s_flag); return; }
   const bool boundary = b >= gput + gint;
   if (!boundary) {
      // ---- interior rows: never look at a flag
      const int r = (b - gput) * G + tid / K;
      const int row = (r < nrows && !(oi && oi[r + 1] > oi[r])) ? r : -1;   // rows with offd entries belong to the boundary blocks
      double sd = 0.0;
      int p0 = 0;
      if (row >= 0) {
         p0 = di[row];
         const int p1 = di[row + 1];
         if (I16) { const double *xr = x + row; for (int p = p0 + skip + lane; p < p1; p += K) sd += da[p] * __ldg(xr + dj16[p]); }
         else     { for (int p = p0 + skip + lane; p < p1; p += K) sd += da[p] * __ldg(x + dj[p]); }
      }
#pragma unroll
      for (int o = K / 2; o > 0; o >>= 1) sd += __shfl_down_sync(0xffffffffu, sd, o, K);
      if (row >= 0 && lane == 0) ea.y[row] = fused_epi_value<EPI>(ea, row, sd, 0.0, false, epi_needs_diag<EPI>() ? da[p0] : 0.0);
      return;
   }
   // ---- boundary rows: a bounded number of blocks walks the list; the diag part of a block's first rows is
   // summed before anybody looks at a flag (the transfer is in flight meanwhile)
   const int bb = b - gput - gint, gbnd = (int) gridDim.x - gput - gint;
   const int trips = (nbnd + gbnd * G - 1) / (gbnd * G);        // the same for every lane: shuffles stay converged
   unsigned long long epoch_in = 0;
   const double *xe = xext_plain;                                // (plain mode: the halo already sits in the receive buffer)
   for (int t = 0; t < (trips > 0 ? trips : 1); t++) {
      const int k = (t * gbnd + bb) * G + tid / K;
      const int row = (k < nbnd) ? bnd_rows[k] : -1;
      double sd = 0.0, so = 0.0;
      int p0 = 0;
      if (row >= 0) {
         p0 = di[row];
         const int p1 = di[row + 1];
         if (I16) { const double *xr = x + row; for (int p = p0 + skip + lane; p < p1; p += K) sd += da[p] * __ldg(xr + dj16[p]); }
         else     { for (int p = p0 + skip + lane; p < p1; p += K) sd += da[p] * __ldg(x + dj[p]); }
      }
#pragma unroll
      for (int o = K / 2; o > 0; o >>= 1) sd += __shfl_down_sync(0xffffffffu, sd, o, K);
      if (t == 0 && !xext_plain) {
         // block 0 of the boundary group polls the peers' flags and republishes the epoch locally
         epoch_in = h.w.epoch_ctr[1] + 1;
         const int par = (int) (epoch_in & 1ull);
         if (bb == 0) {
            for (int j = tid; j < h.w.n_in; j += kFusedThreads) spin_until_ge(h.w.flags + par * h.w.n_in + j, epoch_in, guard, 2, j);
            __syncthreads();
            if (tid == 0) st_release_gpu(h.w.epoch_ctr + 2, epoch_in);
         } else {
            if (tid == 0) spin_local_until_ge(h.w.epoch_ctr + 2, epoch_in, guard);
            __syncthreads();
         }
         xe = par ? h.w.buf1 : h.w.buf0;
      }
      if (row >= 0) { for (int q = oi[row] + lane; q < oi[row + 1]; q += K) so += oa[q] * __ldcg(xe + oj[q]); }
#pragma unroll
      for (int o = K / 2; o > 0; o >>= 1) so += __shfl_down_sync(0xffffffffu, so, o, K);
      if (row >= 0 && lane == 0) ea.y[row] = fused_epi_value<EPI>(ea, row, sd, so, true, epi_needs_diag<EPI>() ? da[p0] : 0.0);
   }
   if (!xext_plain) fused_ack(h, epoch_in, gbnd, &s_flag);
}

template <int EPI, int K>
static int fused_launch_K(const hb200_parcsr *A, const double *x, const PeerFusedArgs &h, const EpiArgs &ea, cudaStream_t st)
{
   const DCsr &D = A->diag, &O = A->offd;
   constexpr int G = kFusedThreads / K;
   const bool has_offd = A->num_cols_offd > 0 && O.num_rownnz > 0 && h.w.n_in > 0;
   int gput = h.n_out > 0 ? (h.total_out + 2047) / 2048 : 0;
   if (gput > 64) gput = 64;
#ifdef HB200_EMU
   if (gput > 1) gput = 1;   // (emulated blocks run one after the other: the put must be complete before a block waits)
#endif
   if (h.n_out > 0 && gput < 1) gput = 1;
   const int gint = (D.nrows + G - 1) / G;
   int gbnd = has_offd ? (O.num_rownnz + G - 1) / G : 0;
   if (gbnd > 2 * kNumSMs) gbnd = 2 * kNumSMs;          // (they loop over the list)
   if (h.w.n_in > 0 && gbnd < 1) gbnd = 1;              // (somebody sends to us: the exchange has to be consumed and acknowledged)
   const int grid = gput + gint + gbnd;
   if (grid < 1) return 0;
   if (D.kind == SPMV_VECTOR16 && D.j16) {
      HB_LAUNCH((parcsr_fused<EPI, K, true>), grid, kFusedThreads, 0, st, D.nrows, D.i, reinterpret_cast<const int *>(D.j16), D.a,
                has_offd ? O.i : (const int *) nullptr, O.j, O.a, O.rownnz, has_offd ? O.num_rownnz : 0, x, h, gput, gint,
                (const double *) nullptr, ea);
   } else {
      HB_LAUNCH((parcsr_fused<EPI, K, false>), grid, kFusedThreads, 0, st, D.nrows, D.i, D.j, D.a,
                has_offd ? O.i : (const int *) nullptr, O.j, O.a, O.rownnz, has_offd ? O.num_rownnz : 0, x, h, gput, gint,
                (const double *) nullptr, ea);
   }
   HB_LAUNCH_CHECK();
   return 0;
}

// the boundary half alone (gint = 0): put + boundary rows over the peer halo, or boundary rows over a receive
// buffer that has been filled already
template <int EPI, int K>
static int boundary_launch_K(const hb200_parcsr *A, const double *x, const PeerFusedArgs &h, const EpiArgs &ea, bool peer, cudaStream_t st)
{
   const DCsr &D = A->diag, &O = A->offd;
   constexpr int G = kFusedThreads / K;
   const int nb = A->num_cols_offd > 0 ? O.num_rownnz : 0;
   int gput = 0, gbnd = (nb + G - 1) / G;
   if (gbnd > 2 * kNumSMs) gbnd = 2 * kNumSMs;          // (they loop over the list)
   PeerFusedArgs hh = h;
   const double *plain = nullptr;
   if (peer) {
      gput = h.n_out > 0 ? (h.total_out + 2047) / 2048 : 0;
      if (gput > 64) gput = 64;
#ifdef HB200_EMU
      if (gput > 1) gput = 1;
#endif
      if (h.n_out > 0 && gput < 1) gput = 1;
      if (h.w.n_in > 0 && gbnd < 1) gbnd = 1;
   } else {
      hh = PeerFusedArgs();
      plain = A->pkg.d_recv_buf;
      if (nb == 0) return 0;
   }
   const int grid = gput + gbnd;
   if (grid < 1) return 0;
   HB_LAUNCH((parcsr_fused<EPI, K, false>), grid, kFusedThreads, 0, st, D.nrows, D.i, D.j, D.a,
             nb ? O.i : (const int *) nullptr, O.j, O.a, O.rownnz, nb, x, hh, gput, 0, plain, ea);
   HB_LAUNCH_CHECK();
   return 0;
}

int parcsr_boundary_launch(hb200_parcsr *A, const double *x, int epi_kind, const EpiArgs &ea, bool peer, cudaStream_t st)
{
   PeerFusedArgs h;
   if (peer) peer_fused_args(A->pkg.fwd, &h);
   // lanes per boundary row from the length of its diag part (the offd part is a handful of entries)
   const double avg = A->diag.avg_row_nnz;
   int lanes = avg >= 150 ? 32 : avg >= 80 ? 16 : avg >= 36 ? 8 : avg >= 10 ? 4 : 2;
   const long long nb = A->num_cols_offd > 0 ? A->offd.num_rownnz : 0;
   while (lanes < 32 && nb * lanes < 148LL * 512 && lanes < avg) lanes *= 2;
#define HB_BND(E) \
   switch (lanes) { \
      case 2:  return boundary_launch_K<E, 2>(A, x, h, ea, peer, st); \
      case 4:  return boundary_launch_K<E, 4>(A, x, h, ea, peer, st); \
      case 8:  return boundary_launch_K<E, 8>(A, x, h, ea, peer, st); \
      case 16: return boundary_launch_K<E, 16>(A, x, h, ea, peer, st); \
      default: return boundary_launch_K<E, 32>(A, x, h, ea, peer, st); \
   }
   switch (epi_kind) {
      case EPI_AXPBY:       HB_BND(EPI_AXPBY)
      case EPI_JACOBI7:     HB_BND(EPI_JACOBI7)
      case EPI_JACOBI_CORE: HB_BND(EPI_JACOBI_CORE)
      default: return set_error(HB200_ERROR_ARG, "parcsr_boundary_launch: epilogue %d", epi_kind);
   }
#undef HB_BND
}

template <int EPI>
static int fused_launch(const hb200_parcsr *A, const double *x, const PeerFusedArgs &h, const EpiArgs &ea, cudaStream_t st)
{
   switch (A->diag.lanes) {
      case 1:  return fused_launch_K<EPI, 1>(A, x, h, ea, st);
      case 2:  return fused_launch_K<EPI, 2>(A, x, h, ea, st);
      case 4:  return fused_launch_K<EPI, 4>(A, x, h, ea, st);
      case 8:  return fused_launch_K<EPI, 8>(A, x, h, ea, st);
      case 16: return fused_launch_K<EPI, 16>(A, x, h, ea, st);
      default: return fused_launch_K<EPI, 32>(A, x, h, ea, st);
   }
}

// the operation qualifies when the halo goes by peer puts, the diag block runs the CSR vector kernel
// and is small enough to be latency-bound (the fine levels hide the exchange behind a long diag pass
// and keep their structured formats)
int parcsr_fused_try(hb200_parcsr *A, const double *x, int epi_kind, const EpiArgs &ea, bool *done)
{
   *done = false;
   Ctx &c = ctx();
   static const bool on = env_flag("HB200_FUSED_HALO", true);
   static const long long max_threads = []() { const char *e = getenv("HB200_FUSED_HALO_MAX"); return e ? atoll(e) : 148LL * 2048 * 2; }();
   if (!on || c.halo_mode != 1 || c.nranks <= 1) return 0;
   if (!(epi_kind == EPI_AXPBY || epi_kind == EPI_JACOBI7 || epi_kind == EPI_JACOBI_CORE)) return 0;
   const DCsr &D = A->diag;
   if (!(D.kind == SPMV_VECTOR || D.kind == SPMV_VECTOR16) || D.lanes < 1) return 0;
   if ((long long) D.nrows * D.lanes > max_threads) return 0;
   if (ea.dot_slot >= 0) return 0;
   HB_CHECK(peer_plans_ensure(A, false));          // collective on first use, like the separate kernels
   if (A->pkg.peer_off) return 0;
   PeerFusedArgs h;
   peer_fused_args(A->pkg.fwd, &h);
   if (D.nrows == 0 && h.n_out == 0 && h.w.n_in == 0) { *done = true; return 0; }
   timer_tick(T_MATVEC_DIAG);
   int f;
   switch (epi_kind) {
      case EPI_AXPBY:   f = fused_launch<EPI_AXPBY>(A, x, h, ea, c.s_comp); break;
      case EPI_JACOBI7: f = fused_launch<EPI_JACOBI7>(A, x, h, ea, c.s_comp); break;
      default:          f = fused_launch<EPI_JACOBI_CORE>(A, x, h, ea, c.s_comp); break;
   }
   timer_tick(T_OTHER);
   if (!f) *done = true;
   return f;
}

// ---- the restriction y = alpha*A^T x + beta*y in one kernel (stored transposes diagT / offdT) -----------
//   put   : the first blocks compute y_tmp = alpha * offdT x row by row and store every value straight into
//           the owner's receive buffer (the reverse exchange, par_csr_matvec.c:402-470), then raise the flags;
//   diag  : every row of diagT;
//   wait  : blocks that own a row somebody contributes to poll the arrival flags;
//   unpack: those rows add the received contributions in ascending send-entry order (the reference's
//           sequential loop, par_csr_matvec.c:491-496), read in place from the NVLink receive buffer;
//   ack.
__global__ void unpack_slot_kernel(int n, const int *__restrict__ rows, int *__restrict__ slot)
{
   const int t = blockIdx.x * blockDim.x + threadIdx.x;
   if (t < n) slot[rows[t]] = t;
}

// block roles: [0, gput) compute y_tmp = alpha * offdT x and put it; [gput, gput + gint) the coarse rows nobody
// contributes to; [gput + gint, grid) the rows of the unpack plan: diagT part, flags, contributions in order
template <int K, bool I16>
__global__ void __launch_bounds__(kFusedThreads)
parcsr_fusedT(int ncoarse, const int *__restrict__ di, const int *__restrict__ dj, const double *__restrict__ da,
              int noffd, const int *__restrict__ oi, const int *__restrict__ oj, const double *__restrict__ oa,
              const double *__restrict__ x, double alpha, double beta, double *__restrict__ y,
              const int *__restrict__ unpack_slot, const int *__restrict__ unpack_rows, int nunpack,
              const int *__restrict__ unpack_ptr, const int *__restrict__ unpack_idx,
              PeerFusedArgs h, int gput, int gint)
{
   __shared__ int s_flag;
   const int tid = threadIdx.x;
   const int lane = tid % K;
   constexpr int G = kFusedThreads / K;
   SpinGuard guard;
   guard.err = h.w.err; guard.timeout_ns = h.w.timeout_ns;
   const int b = (int) blockIdx.x;
   if (b < gput) {
      // ---- put: y_tmp rows (the columns of the offd block), straight into the owners' buffers
      const unsigned long long epoch = h.w.epoch_ctr[0] + 1;
      const int par = (int) (epoch & 1ull);
      if (epoch > 2) {
         for (int i = tid; i < h.n_out; i += kFusedThreads) spin_until_ge(h.acks + i, epoch - 2, guard, 1, i);
         __syncthreads();
      }
      const int trips = (noffd + gput * G - 1) / (gput * G);     // same trip count for every lane: shuffles stay converged
      for (int t = 0; t < trips; t++) {
         const int c = (t * gput + b) * G + tid / K;
         double s = 0.0;
         if (c < noffd) { for (int q = oi[c] + lane; q < oi[c + 1]; q += K) s += oa[q] * __ldg(x + oj[q]); }
#pragma unroll
         for (int o = K / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, K);
         if (c < noffd && lane == 0) {
            int lo = 0, hi = h.n_out - 1;
            while (lo < hi) {
               const int mid = (lo + hi + 1) >> 1;
               if (h.out_starts[mid] <= c) lo = mid; else hi = mid - 1;
            }
            h.dst2[par * h.n_out + lo][c - h.out_starts[lo]] = alpha * s;
         }
      }
      __syncthreads();
      if (tid == 0) {
         __threadfence_system();
         int last = 1;
         if (gput > 1) {
            const unsigned int t = atomicInc(h.w.ticket, (unsigned int) gput - 1);
            last = (t == (unsigned int) gput - 1);
            if (last) __threadfence_system();
         }
         s_flag = last;
      }
      __syncthreads();
      if (s_flag) {
         for (int i = tid; i < h.n_out; i += kFusedThreads) st_release_sys(h.flag2[par * h.n_out + i], epoch);
         if (tid == 0) h.w.epoch_ctr[0] = epoch;
      }
      return;
   }
   const bool boundary = b >= gput + gint;
   const short *__restrict__ dj16 = reinterpret_cast<const short *>(dj);
   int row = -1, slot = -1;
   if (!boundary) {
      const int r = (b - gput) * G + tid / K;
      if (r < ncoarse && !(unpack_slot && unpack_slot[r] >= 0)) row = r;
   } else {
      const int t = (b - gput - gint) * G + tid / K;
      if (t < nunpack) { row = unpack_rows[t]; slot = t; }
   }
   double sd = 0.0;
   if (row >= 0) {
      const int p0 = di[row], p1 = di[row + 1];
      if (I16) { const double *xr = x + row; for (int p = p0 + lane; p < p1; p += K) sd += da[p] * __ldg(xr + dj16[p]); }
      else     { for (int p = p0 + lane; p < p1; p += K) sd += da[p] * __ldg(x + dj[p]); }
   }
#pragma unroll
   for (int o = K / 2; o > 0; o >>= 1) sd += __shfl_down_sync(0xffffffffu, sd, o, K);
   double v = 0.0;
   if (row >= 0 && lane == 0) v = (beta == 0.0) ? alpha * sd : beta * y[row] + alpha * sd;
   unsigned long long epoch_in = 0;
   if (boundary) {
      epoch_in = h.w.epoch_ctr[1] + 1;
      const int par = (int) (epoch_in & 1ull);
      for (int j = tid; j < h.w.n_in; j += kFusedThreads) spin_until_ge(h.w.flags + par * h.w.n_in + j, epoch_in, guard, 2, j);
      __syncthreads();
      const double *buf = par ? h.w.buf1 : h.w.buf0;
      if (row >= 0 && lane == 0) {
         for (int q = unpack_ptr[slot]; q < unpack_ptr[slot + 1]; q++) v = __dadd_rn(v, __ldcg(buf + unpack_idx[q]));
      }
   }
   if (row >= 0 && lane == 0) y[row] = v;
   if (boundary) fused_ack(h, epoch_in, (int) gridDim.x - gput - gint, &s_flag);
}

template <int K>
static int fusedT_launch_K(hb200_parcsr *A, const double *x, double alpha, double beta, double *y, const PeerFusedArgs &h, cudaStream_t st)
{
   const DCsr &D = A->diagT, &O = A->offdT;
   const CommPkgD &pk = A->pkg;
   constexpr int G = kFusedThreads / K;
   const bool has_offd = A->num_cols_offd > 0;
   int gput = h.n_out > 0 ? (A->num_cols_offd + G - 1) / G : 0;
   if (gput > 32) gput = 32;
#ifdef HB200_EMU
   if (gput > 1) gput = 1;   // (emulated blocks run one after the other: the put must be complete before a block waits)
#endif
   if (h.n_out > 0 && gput < 1) gput = 1;
   const int gint = (D.nrows + G - 1) / G;
   const int nunpack = h.w.n_in > 0 ? pk.n_unpack_rows : 0;
   int gbnd = (nunpack + G - 1) / G;
   if (h.w.n_in > 0 && gbnd < 1) gbnd = 1;
   const int grid = gput + gint + gbnd;
   if (grid < 1) return 0;
   if (D.kind == SPMV_VECTOR16 && D.j16) {
      HB_LAUNCH((parcsr_fusedT<K, true>), grid, kFusedThreads, 0, st, D.nrows, D.i, reinterpret_cast<const int *>(D.j16), D.a,
                has_offd ? O.nrows : 0, O.i, O.j, O.a, x, alpha, beta, y, nunpack ? A->d_unpack_slot : (int *) nullptr,
                pk.d_unpack_rows, nunpack, pk.d_unpack_ptr, pk.d_unpack_idx, h, gput, gint);
   } else {
      HB_LAUNCH((parcsr_fusedT<K, false>), grid, kFusedThreads, 0, st, D.nrows, D.i, D.j, D.a,
                has_offd ? O.nrows : 0, O.i, O.j, O.a, x, alpha, beta, y, nunpack ? A->d_unpack_slot : (int *) nullptr,
                pk.d_unpack_rows, nunpack, pk.d_unpack_ptr, pk.d_unpack_idx, h, gput, gint);
   }
   HB_LAUNCH_CHECK();
   return 0;
}

int parcsr_fusedT_try(hb200_parcsr *A, double alpha, const double *x, double beta, double *y, bool *done)
{
   *done = false;
   Ctx &c = ctx();
   static const bool on = env_flag("HB200_FUSED_HALO", true);
   static const long long max_threads = []() { const char *e = getenv("HB200_FUSED_HALO_MAX"); return e ? atoll(e) : 148LL * 2048 * 2; }();
   if (!on || c.halo_mode != 1 || c.nranks <= 1 || !A->has_T) return 0;
   const DCsr &D = A->diagT;
   if (!(D.kind == SPMV_VECTOR || D.kind == SPMV_VECTOR16) || D.lanes < 1) return 0;
   if ((long long) D.nrows * D.lanes > max_threads || (long long) A->num_cols_offd * D.lanes > max_threads) return 0;
   HB_CHECK(peer_plans_ensure(A, true));
   if (A->pkg.peer_off_rev) return 0;
   PeerFusedArgs h;
   peer_fused_args(A->pkg.rev, &h);
   if (D.nrows == 0 && h.n_out == 0 && h.w.n_in == 0) { *done = true; return 0; }
   // row -> position in the unpack plan (built once)
   if (!A->d_unpack_slot && A->pkg.n_unpack_rows > 0 && D.nrows > 0) {
      HB_CUDA(cudaMalloc(&A->d_unpack_slot, sizeof(int) * (size_t) D.nrows));
      HB_CUDA(cudaMemsetAsync(A->d_unpack_slot, 0xff, sizeof(int) * (size_t) D.nrows, c.s_comp));
      HB_LAUNCH(unpack_slot_kernel, (A->pkg.n_unpack_rows + 255) / 256, 256, 0, c.s_comp, A->pkg.n_unpack_rows, A->pkg.d_unpack_rows, A->d_unpack_slot);
      HB_LAUNCH_CHECK();
   }
   timer_tick(T_MATVEC_DIAG);
   int f;
   switch (D.lanes) {
      case 1:  f = fusedT_launch_K<1>(A, x, alpha, beta, y, h, c.s_comp); break;
      case 2:  f = fusedT_launch_K<2>(A, x, alpha, beta, y, h, c.s_comp); break;
      case 4:  f = fusedT_launch_K<4>(A, x, alpha, beta, y, h, c.s_comp); break;
      case 8:  f = fusedT_launch_K<8>(A, x, alpha, beta, y, h, c.s_comp); break;
      case 16: f = fusedT_launch_K<16>(A, x, alpha, beta, y, h, c.s_comp); break;
      default: f = fusedT_launch_K<32>(A, x, alpha, beta, y, h, c.s_comp); break;
   }
   timer_tick(T_OTHER);
   if (!f) *done = true;
   return f;
}

int spmv_offd_wait_launch(const DCsr &M, const PeerWaitArgs &w, int epi_kind, const EpiArgs &ea, cudaStream_t st)
{
   switch (epi_kind) {
      case EPI_ACC:             return offd_launch<EPI_ACC>(M, w, ea, st);
      case EPI_JACOBI7_ACC:     return offd_launch<EPI_JACOBI7_ACC>(M, w, ea, st);
      case EPI_JACOBI_CORE_ACC: return offd_launch<EPI_JACOBI_CORE_ACC>(M, w, ea, st);
      default: return set_error(HB200_ERROR_ARG, "spmv_offd_wait_launch: epilogue %d is not an offd pass", epi_kind);
   }
}

}  // namespace hb
