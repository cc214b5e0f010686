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
