// kernels_sell.cu — dictionary-packed sliced-ELL SpMV for structured-grid operators.
//
// The CSR model costs 12 B per nonzero (4 B column index + 8 B value).  On the fine level of a
// grid problem both fields are massively redundant:
//   * column index = row + d with d drawn from a handful of offsets (7 for the 7-point, 27 for the
//     27-point stencil, whatever the grid size);
//   * a constant-coefficient operator has a handful of distinct values (the 27-point Laplacian: 2).
// When a square block has <= 256 distinct offsets it is stored a second time as SELL-32: slices
// of 32 consecutive rows, entry k of lane r at [slice][k][r], one byte of offset code per entry and
// either one byte of value code (<= 256 distinct values) or the raw fp64 value; four consecutive
// entries of a row are stored contiguously so that a lane fetches 4 codes with one 32-bit load.  The fine-level
// SpMV then streams 2 B (or 9 B) per nonzero instead of 12, and a warp's gather of x for entry k
// touches 32 consecutive doubles (same offset across neighbouring rows) instead of 16+ sectors.
//
// One thread owns one row and adds its products in CSR order with separate multiply and add, i.e.
// exactly the reference's sequential loop (src/seq_mv/csr_matvec.c:683-721): results are
// bit-identical to the 1-thread CPU reference.  Lossless; falls back to CSR when a block does
// not qualify (every AMG coarse level: irregular offsets, all values distinct).
#include "hb_internal.cuh"
#include "hb_epilogue.cuh"
#include <stdlib.h>
#include <string.h>

namespace hb {

constexpr int kSellThreads = 256;

// tiny open-addressing map (<= 256 live keys in 2048 slots): the dictionaries are built with one
// lookup per nonzero on the host, so this has to be cheap
struct SmallMap {
   unsigned long long keys[2048];
   short code[2048];
   int   count = 0;
   SmallMap() { for (int k = 0; k < 2048; k++) code[k] = -1; }
   static inline unsigned slot(unsigned long long key) { return (unsigned) ((key * 0x9E3779B97F4A7C15ull) >> 53); }
   inline int find(unsigned long long key) const
   {
      unsigned s = slot(key);
      while (code[s] >= 0) { if (keys[s] == key) return code[s]; s = (s + 1) & 2047; }
      return -1;
   }
   inline int add(unsigned long long key)   // caller checked find() < 0 and count < 256
   {
      unsigned s = slot(key);
      while (code[s] >= 0) s = (s + 1) & 2047;
      keys[s] = key; code[s] = (short) count;
      return count++;
   }
};

// Layout: slice s (32 rows) holds G = ceil(maxlen/4) groups; group g of lane r sits at
// sell_ptr[s] + g*128 + r*4 .. +3  (4 consecutive entries of one row are contiguous), so a lane
// fetches 4 offset codes with one 32-bit load and a warp's load is one coalesced 128 B line.
template <int EPI, bool VCODED>
__global__ void __launch_bounds__(kSellThreads)
spmv_sell(int nrows, const int *__restrict__ rowptr, const long long *__restrict__ sptr,
          const unsigned char *__restrict__ cidx, const unsigned char *__restrict__ vidx,
          const double *__restrict__ vraw, const int *__restrict__ offdict,
          const double *__restrict__ valdict, const double *__restrict__ x, EpiArgs ea)
{
   __shared__ int    s_off[256];
   __shared__ double s_val[256];
   const int tid = threadIdx.x;
   s_off[tid] = offdict[tid];
   if (VCODED) s_val[tid] = valdict[tid];
   __syncthreads();
   const int row = blockIdx.x * kSellThreads + tid;
   if (row >= nrows) return;
   const int lane = tid & 31;
   const int slice = row >> 5;
   const int len = rowptr[row + 1] - rowptr[row];
   const long long base = sptr[slice] + lane * 4;
   const int skip = (EPI == EPI_JACOBI_CORE) ? ea.skip_diag : 0;
   double s = 0.0, diag = 0.0;
   if (epi_needs_diag<EPI>() && len > 0) diag = VCODED ? s_val[vidx[base]] : vraw[base];
   const double *xr = x + row;
   int k = 0;
   // two groups (8 entries) in flight per thread: codes first, then the dependent gathers
   for (; k + 8 <= len; k += 8) {
      const long long e0 = base + (long long) (k >> 2) * 128;
      const unsigned int c0 = *reinterpret_cast<const unsigned int *>(cidx + e0);
      const unsigned int c1 = *reinterpret_cast<const unsigned int *>(cidx + e0 + 128);
      double a[8], xv[8];
      if (VCODED) {
         const unsigned int v0 = *reinterpret_cast<const unsigned int *>(vidx + e0);
         const unsigned int v1 = *reinterpret_cast<const unsigned int *>(vidx + e0 + 128);
#pragma unroll
         for (int u = 0; u < 4; u++) { a[u] = s_val[(v0 >> (8 * u)) & 255u]; a[4 + u] = s_val[(v1 >> (8 * u)) & 255u]; }
      } else {
         const double2 p0 = *reinterpret_cast<const double2 *>(vraw + e0);
         const double2 p1 = *reinterpret_cast<const double2 *>(vraw + e0 + 2);
         const double2 p2 = *reinterpret_cast<const double2 *>(vraw + e0 + 128);
         const double2 p3 = *reinterpret_cast<const double2 *>(vraw + e0 + 130);
         a[0] = p0.x; a[1] = p0.y; a[2] = p1.x; a[3] = p1.y; a[4] = p2.x; a[5] = p2.y; a[6] = p3.x; a[7] = p3.y;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
         xv[u]     = __ldg(xr + s_off[(c0 >> (8 * u)) & 255u]);
         xv[4 + u] = __ldg(xr + s_off[(c1 >> (8 * u)) & 255u]);
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
         if (k + u >= skip) s = __dadd_rn(s, __dmul_rn(a[u], xv[u]));
      }
   }
   for (; k < len; k += 4) {
      const long long e0 = base + (long long) (k >> 2) * 128;
      const unsigned int c0 = *reinterpret_cast<const unsigned int *>(cidx + e0);
      unsigned int v0 = 0;
      if (VCODED) v0 = *reinterpret_cast<const unsigned int *>(vidx + e0);
#pragma unroll
      for (int u = 0; u < 4; u++) {
         if (k + u < len && k + u >= skip) {
            const double a = VCODED ? s_val[(v0 >> (8 * u)) & 255u] : vraw[e0 + u];
            s = __dadd_rn(s, __dmul_rn(a, __ldg(xr + s_off[(c0 >> (8 * u)) & 255u])));
         }
      }
   }
   epi_apply<EPI>(ea, row, s, diag);
}

template <int EPI>
static int sell_dispatch(const DCsr &M, const double *x, const EpiArgs &ea, cudaStream_t st)
{
   const int grid = (M.nrows + kSellThreads - 1) / kSellThreads;
   if (M.sell_vidx) {
      HB_LAUNCH((spmv_sell<EPI, true>), grid, kSellThreads, 0, st, M.nrows, M.i, M.sell_ptr, M.sell_cidx,
                M.sell_vidx, (const double *) nullptr, M.sell_offdict, M.sell_valdict, x, ea);
   } else {
      HB_LAUNCH((spmv_sell<EPI, false>), grid, kSellThreads, 0, st, M.nrows, M.i, M.sell_ptr, M.sell_cidx,
                (const unsigned char *) nullptr, M.sell_val, M.sell_offdict, M.sell_valdict, x, ea);
   }
   HB_LAUNCH_CHECK();
   return 0;
}

int spmv_sell_launch(const DCsr &M, const double *x, int epi_kind, const EpiArgs &ea, cudaStream_t st)
{
   switch (epi_kind) {
      case EPI_AXPBY:           return sell_dispatch<EPI_AXPBY>(M, x, ea, st);
      case EPI_ACC:             return sell_dispatch<EPI_ACC>(M, x, ea, st);
      case EPI_JACOBI7:         return sell_dispatch<EPI_JACOBI7>(M, x, ea, st);
      case EPI_JACOBI7_ACC:     return sell_dispatch<EPI_JACOBI7_ACC>(M, x, ea, st);
      case EPI_JACOBI_CORE:     return sell_dispatch<EPI_JACOBI_CORE>(M, x, ea, st);
      case EPI_JACOBI_CORE_ACC: return sell_dispatch<EPI_JACOBI_CORE_ACC>(M, x, ea, st);
      default: return set_error(HB200_ERROR_ARG, "spmv_sell_launch: unknown epilogue %d", epi_kind);
   }
}

int dcsr_free_sell(DCsr &M)
{
   if (M.sell_ptr) cudaFree(M.sell_ptr);
   if (M.sell_cidx) cudaFree(M.sell_cidx);
   if (M.sell_vidx) cudaFree(M.sell_vidx);
   if (M.sell_val) cudaFree(M.sell_val);
   if (M.sell_offdict) cudaFree(M.sell_offdict);
   if (M.sell_valdict) cudaFree(M.sell_valdict);
   M.sell_ptr = nullptr; M.sell_cidx = nullptr; M.sell_vidx = nullptr; M.sell_val = nullptr;
   M.sell_offdict = nullptr; M.sell_valdict = nullptr;
   M.has_sell = false;
   return 0;
}

int dcsr_build_sell(DCsr &M, const int *hi, const int *hj, const double *ha)
{
   if (env_flag("HB200_NO_SELL", false)) return 0;
   const int n = M.nrows;
   const long long nnz = M.nnz;
   if (n < 1024 || nnz == 0) return 0;               // tiny blocks: nothing to win
   // ---- offset dictionary (early exit: irregular blocks fail within the first rows)
   SmallMap offmap, valmap;
   std::vector<int> offdict;
   for (int r = 0; r < n; r++) {
      for (int p = hi[r]; p < hi[r + 1]; p++) {
         const unsigned long long d = (unsigned long long) (long long) (hj[p] - r);
         if (offmap.find(d) < 0) {
            if (offmap.count >= 256) return 0;
            offmap.add(d);
            offdict.push_back(hj[p] - r);
         }
      }
   }
   // ---- value dictionary (bit patterns, so that -0.0 / NaN payloads survive)
   std::vector<double> valdict;
   bool vcoded = true;
   {
      unsigned long long last = 0;
      bool have_last = false;
      for (long long p = 0; p < nnz; p++) {
         unsigned long long bits;
         memcpy(&bits, &ha[p], 8);
         if (have_last && bits == last) continue;
         last = bits; have_last = true;
         if (valmap.find(bits) < 0) {
            if (valmap.count >= 256) { vcoded = false; break; }
            valmap.add(bits);
            valdict.push_back(ha[p]);
         }
      }
   }
   // ---- slices: G = ceil(maxlen/4) groups of 4 entries per lane
   const int nslices = (n + 31) / 32;
   std::vector<long long> sptr((size_t) nslices + 1, 0);
   for (int s = 0; s < nslices; s++) {
      int mx = 0;
      for (int r = s * 32; r < n && r < s * 32 + 32; r++) mx = std::max(mx, hi[r + 1] - hi[r]);
      sptr[s + 1] = sptr[s] + (long long) ((mx + 3) / 4) * 128;
   }
   const long long total = sptr[nslices];
   if ((double) total > 1.6 * (double) nnz + 4096.0) return 0;   // too much padding: ragged rows
   std::vector<unsigned char> cidx((size_t) total, 0), vidx;
   std::vector<double> vraw;
   if (vcoded) vidx.assign((size_t) total, 0); else vraw.assign((size_t) total, 0.0);
   for (int r = 0; r < n; r++) {
      const long long base = sptr[r >> 5] + (long long) (r & 31) * 4;
      for (int p = hi[r], k = 0; p < hi[r + 1]; p++, k++) {
         const long long e = base + (long long) (k >> 2) * 128 + (k & 3);
         cidx[(size_t) e] = (unsigned char) offmap.find((unsigned long long) (long long) (hj[p] - r));
         if (vcoded) {
            unsigned long long bits;
            memcpy(&bits, &ha[p], 8);
            vidx[(size_t) e] = (unsigned char) valmap.find(bits);
         } else {
            vraw[(size_t) e] = ha[p];
         }
      }
   }
   offdict.resize(256, 0);
   valdict.resize(256, 0.0);
   HB_CUDA(cudaMalloc(&M.sell_ptr, sizeof(long long) * sptr.size()));
   HB_CUDA(cudaMemcpy(M.sell_ptr, sptr.data(), sizeof(long long) * sptr.size(), cudaMemcpyHostToDevice));
   HB_CUDA(cudaMalloc(&M.sell_cidx, (size_t) total + 64));
   HB_CUDA(cudaMemcpy(M.sell_cidx, cidx.data(), (size_t) total, cudaMemcpyHostToDevice));
   if (vcoded) {
      HB_CUDA(cudaMalloc(&M.sell_vidx, (size_t) total + 64));
      HB_CUDA(cudaMemcpy(M.sell_vidx, vidx.data(), (size_t) total, cudaMemcpyHostToDevice));
   } else {
      HB_CUDA(cudaMalloc(&M.sell_val, sizeof(double) * ((size_t) total + 8)));
      HB_CUDA(cudaMemcpy(M.sell_val, vraw.data(), sizeof(double) * (size_t) total, cudaMemcpyHostToDevice));
   }
   HB_CUDA(cudaMalloc(&M.sell_offdict, sizeof(int) * 256));
   HB_CUDA(cudaMemcpy(M.sell_offdict, offdict.data(), sizeof(int) * 256, cudaMemcpyHostToDevice));
   HB_CUDA(cudaMalloc(&M.sell_valdict, sizeof(double) * 256));
   HB_CUDA(cudaMemcpy(M.sell_valdict, valdict.data(), sizeof(double) * 256, cudaMemcpyHostToDevice));
   M.sell_nslices = nslices;
   M.sell_nd = offmap.count;
   M.sell_nv = vcoded ? valmap.count : 0;
   M.has_sell = true;
   return 0;
}

}  // namespace hb
