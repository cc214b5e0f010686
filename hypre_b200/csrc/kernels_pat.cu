// kernels_pat.cu — row-pattern SpMV for operators with few distinct rows.
//
// A row of a CSR block is the list (column - base, value) of its entries, with base = the row
// index for square blocks and the row's first column for rectangular ones.  On a constant-
// coefficient grid problem only a handful of distinct lists exist on the fine level — one for the
// interior and one per kind of boundary, 27 on a box for the 7-point and the 27-point stencil
// alike — and, where the coarsening comes out regular, the same holds for the interpolation P_0,
// its stored transpose and the Galerkin operator A_1 (about 100 patterns each on 27-pt).
// When a block has <= 256 distinct row patterns (<= 8192 table entries) it is stored a second
// time as
//     pat[row]           1 byte per ROW        (+ base[row], 4 bytes, for rectangular blocks)
//     table[pattern]     the (offset, value) list of each pattern, held in shared memory
// so the SpMV streams x, y and one byte per row — 17 B per row instead of 12 B per NONZERO — and
// the kernel is bound by the L1 gathers of x, not by HBM.
//
// One thread owns R rows that sit one block width apart (a block covers R*blockDim consecutive
// rows; for fixed j a warp's 32 rows are consecutive, so on square blocks every gather of x is a
// coalesced 256 B request).  When the R rows share a pattern — the interior of A, i.e. nearly
// always — each (offset, value) is read from shared memory once and used R times.  Every row
// adds its products in CSR order with separate multiply and add, exactly the reference's
// sequential loop (src/seq_mv/csr_matvec.c:683-721): results are bit-identical to the 1-thread
// CPU reference.  Lossless; blocks that do not qualify (variable coefficients, the deeper AMG
// levels) keep the packed-SELL or CSR path.
#include "hb_internal.cuh"
#include <atomic>
#include <thread>
#include "hb_epilogue.cuh"
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <unordered_map>

namespace hb {

constexpr int kPatMaxPatterns = 256;   // codes 0..254; 255 = row outside the table
constexpr int kPatMaxEntries = 8192;   // table entries over all patterns (96 KB of shared memory)

template <int EPI>
__device__ __forceinline__ void pat_one_row(const EpiArgs &ea, const double *__restrict__ xb, int row, int p,
                                            const int *s_ptr, const int *s_off, const double *s_val, int skip)
{
   const int b = s_ptr[p], e = s_ptr[p + 1];
   double s = 0.0;
   for (int k = b + skip; k < e; k++) s = __dadd_rn(s, __dmul_rn(s_val[k], __ldg(xb + s_off[k])));
   epi_apply<EPI>(ea, row, s, e > b ? s_val[b] : 0.0);
}

// ---- fused dot (DOT variants): <y, w> of the vector the kernel writes, two-stage and in a fixed
// order (bitwise reproducible): per thread over its rows, per block by warp shuffles, the last
// block to arrive adds the block partials.  Region 3 of the reduction scratch of the context.
__device__ double       *g_pd_partials = nullptr;
__device__ unsigned int *g_pd_counter = nullptr;
__device__ double       *g_pd_scalars = nullptr;

template <int NT>
__device__ __forceinline__ void pat_dot_finish(double acc, int slot)
{
   __shared__ double sm[NT / 32];
   __shared__ bool   is_last;
   double *partials = g_pd_partials + 3 * (size_t) kRedBlocksMax;
   unsigned int *counter = g_pd_counter + 3;
   const int tid = threadIdx.x;
   double s = acc;
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
   if ((tid & 31) == 0) sm[tid >> 5] = s;
   __syncthreads();
   if (tid == 0) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < NT / 32; w++) t += sm[w];
      partials[blockIdx.x] = t;
      __threadfence();
      const unsigned int ticket = atomicInc(counter, gridDim.x - 1);
      is_last = (ticket == gridDim.x - 1);
   }
   __syncthreads();
   if (is_last) {
      __threadfence();
      double t = 0.0;
      for (int b = tid; b < (int) gridDim.x; b += NT) t += ((volatile double *) partials)[b];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
      __syncthreads();
      if ((tid & 31) == 0) sm[tid >> 5] = t;
      __syncthreads();
      if (tid == 0) {
         double r = 0.0;
#pragma unroll
         for (int w = 0; w < NT / 32; w++) r += sm[w];
         g_pd_scalars[slot] = r;
      }
   }
}

// column of entry k of a row = base + off[k]; base = the row itself for square blocks (BASE =
// false), else base[row] (the row's first column: interpolation and its stored transpose).
// NT threads per block, R rows per thread.
// WIDE (HB200_PAT_WIDE=1, default on N > 1): 16-bit row codes (<= 65534 patterns, 65535 = row outside
// the table) and the table read from global memory through L1 instead of shared memory — the
// partitioned coarse operators, whose numbering next to a rank boundary multiplies the patterns.
template <int EPI, bool BASE, int NT, int R, bool WIDE = false, bool DOT = false>
__global__ void __launch_bounds__(NT, 1536 / NT)
spmv_pat(int nrows, int ntiles, const unsigned char *__restrict__ pat, const int *__restrict__ base, int npat,
         int nent, const int *__restrict__ tab_ptr, const int *__restrict__ tab_off,
         const double *__restrict__ tab_val, const double *__restrict__ x, EpiArgs ea)
{
   HB_DYN_SHARED(double, s_mem);
   const double *s_val;
   const int    *s_off, *s_ptr;
   const int tid = threadIdx.x;
   if (WIDE) {
      s_val = tab_val; s_off = tab_off; s_ptr = tab_ptr;
   } else {
      double *w_val = s_mem;
      int    *w_off = reinterpret_cast<int *>(w_val + nent);
      int    *w_ptr = w_off + nent;
      for (int k = tid; k < nent; k += NT) { w_val[k] = tab_val[k]; w_off[k] = tab_off[k]; }
      for (int k = tid; k <= npat; k += NT) w_ptr[k] = tab_ptr[k];
      __syncthreads();
      s_val = w_val; s_off = w_off; s_ptr = w_ptr;
   }
   const int skip = (EPI == EPI_JACOBI_CORE) ? ea.skip_diag : 0;
   double dacc = 0.0;                     // DOT: this thread's share of <y, dotw>
   for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int r0 = tile * (NT * R) + tid;
      int p[R];
      int bs[R];
      bool same = true;
#pragma unroll
      for (int j = 0; j < R; j++) {
         const int row = r0 + j * NT;
         if (WIDE) {
            p[j] = row < nrows ? (int) reinterpret_cast<const unsigned short *>(pat)[row] : 65535;
            if (p[j] == 65535) p[j] = -1;
         } else {
            p[j] = row < nrows ? (int) pat[row] : 255;
            if (p[j] == 255) p[j] = -1;        // past the end, or an irregular row (CSR pass)
         }
         if (BASE) bs[j] = row < nrows ? base[row] : 0;
         same = same && (p[j] == p[0]);
      }
      if (same && p[0] >= 0) {
         // the R rows share a pattern: one table read per entry, R gathers
         const int b = s_ptr[p[0]], e = s_ptr[p[0] + 1];
         double s[R];
#pragma unroll
         for (int j = 0; j < R; j++) s[j] = 0.0;
#pragma unroll 2
         for (int k = b + skip; k < e; k++) {
            const double a = s_val[k];
            const double *xk = x + s_off[k];
            double xv[R];
#pragma unroll
            for (int j = 0; j < R; j++) xv[j] = BASE ? __ldg(xk + bs[j]) : __ldg(xk + r0 + j * NT);
#pragma unroll
            for (int j = 0; j < R; j++) s[j] = __dadd_rn(s[j], __dmul_rn(a, xv[j]));
         }
         const double diag = e > b ? s_val[b] : 0.0;
         if (DOT) {
#pragma unroll
            for (int j = 0; j < R; j++) {
               const int row = r0 + j * NT;
               dacc += epi_apply_ret<EPI>(ea, row, s[j]) * __ldg(ea.dotw + row);
            }
         } else {
#pragma unroll
            for (int j = 0; j < R; j++) epi_apply<EPI>(ea, r0 + j * NT, s[j], diag);
         }
      } else {
#pragma unroll
         for (int j = 0; j < R; j++) {
            if (p[j] >= 0) {
               const int row = r0 + j * NT;
               if (DOT) {
                  const int b = s_ptr[p[j]], e = s_ptr[p[j] + 1];
                  const double *xb = x + (BASE ? bs[j] : row);
                  double sr = 0.0;
                  for (int k = b + skip; k < e; k++) sr = __dadd_rn(sr, __dmul_rn(s_val[k], __ldg(xb + s_off[k])));
                  dacc += epi_apply_ret<EPI>(ea, row, sr) * __ldg(ea.dotw + row);
               } else {
                  pat_one_row<EPI>(ea, x + (BASE ? bs[j] : row), row, p[j], s_ptr, s_off, s_val, skip);
               }
            }
         }
      }
   }
   if (DOT) pat_dot_finish<NT>(dacc, ea.dot_slot);
}

static int pat_rows_per_thread()
{
   static int r = 0;
   if (r == 0) {
      const char *e = getenv("HB200_PAT_ROWS");
      r = (e && atoi(e) == 8) ? 8 : 4;
   }
   return r;
}

template <int EPI, bool BASE, int NT, int R, bool WIDE = false, bool DOT = false>
static int pat_launch_t(const DCsr &M, const double *x, const EpiArgs &ea, cudaStream_t st)
{
   const size_t smem = WIDE ? 0 : (size_t) M.pat_nent * 12 + (size_t) (M.pat_npat + 1) * 4 + 8;
   static bool opted = false;
   if (!opted && !WIDE) {
      HB_CUDA(cudaFuncSetAttribute(spmv_pat<EPI, BASE, NT, R, WIDE, DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kPatMaxEntries * 12 + (kPatMaxPatterns + 1) * 4 + 8));
      opted = true;
   }
   const int ntiles = (M.nrows + NT * R - 1) / (NT * R);
   // the table is loaded once per block and the resident blocks walk the tiles: the grid is exactly
   // one wave (a partial second wave would double the time)
   static size_t occ_smem = (size_t) -1;
   static int occ_blocks = 1;
   if (occ_smem != smem) {
      int nb = 0;
      HB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, spmv_pat<EPI, BASE, NT, R, WIDE, DOT>, NT, smem));
      occ_blocks = nb > 0 ? nb : 1;
      occ_smem = smem;
   }
   int grid = 148 * occ_blocks;
   if (grid > ntiles) grid = ntiles;
   HB_LAUNCH((spmv_pat<EPI, BASE, NT, R, WIDE, DOT>), grid, NT, smem, st, M.nrows, ntiles, M.pat_code, M.pat_base, M.pat_npat,
             M.pat_nent, M.pat_ptr, M.pat_off, M.pat_val, x, ea);
   HB_LAUNCH_CHECK();
   return 0;
}

bool fused_dots_enabled()
{
   // on by default (B200, 27-pt 256^3: 49.97 -> 48.18 ms per solve, same iterations and residual;
   // profiles/r2_session_log.md); HB200_FUSED_DOTS=0 keeps the separate dot kernels
   static const bool on = env_flag("HB200_FUSED_DOTS", true);
   return on;
}

bool spmv_can_fuse_dot(const DCsr &M, int epi_kind)
{
   return fused_dots_enabled() && (epi_kind == EPI_AXPBY || epi_kind == EPI_JACOBI7) &&
          (M.kind == SPMV_PAT || M.kind == SPMV_BOX) && M.has_pat && !M.pat_wide && !M.pat_base && M.pat_nirr == 0;
}

// the fused-dot instantiations exist for the two epilogues that produce a Krylov dot operand
template <int EPI>
static int pat_dispatch_dot(const DCsr &M, const double *x, const EpiArgs &ea, cudaStream_t st)
{
   static bool bound = false;
   if (!bound) {
      Ctx &c = ctx();
      HB_CUDA(cudaMemcpyToSymbol(g_pd_partials, &c.d_partials, sizeof(double *)));
      HB_CUDA(cudaMemcpyToSymbol(g_pd_counter, &c.d_counter, sizeof(unsigned int *)));
      HB_CUDA(cudaMemcpyToSymbol(g_pd_scalars, &c.d_scalars, sizeof(double *)));
      bound = true;
   }
   const size_t smem = (size_t) M.pat_nent * 12 + (size_t) (M.pat_npat + 1) * 4 + 8;
   return smem > 40 * 1024 ? pat_launch_t<EPI, false, 512, 4, false, true>(M, x, ea, st)
                           : pat_launch_t<EPI, false, 256, 4, false, true>(M, x, ea, st);
}

template <int EPI>
static int pat_dispatch(const DCsr &M, const double *x, const EpiArgs &ea, cudaStream_t st)
{
   if (M.pat_wide) {
      // 16-bit codes, table in global memory (no shared-memory limit on the block size)
      if (M.pat_base) {
         return M.avg_row_nnz < 8.0 ? pat_launch_t<EPI, true, 256, 4, true>(M, x, ea, st)
                                    : pat_launch_t<EPI, true, 256, 1, true>(M, x, ea, st);
      }
      return pat_launch_t<EPI, false, 256, 4, true>(M, x, ea, st);
   }
   const size_t smem = (size_t) M.pat_nent * 12 + (size_t) (M.pat_npat + 1) * 4 + 8;
   // big tables leave room for few blocks per SM: use larger ones
   const bool big = smem > 40 * 1024;
   if (M.pat_base) {
      // rectangular blocks.  Interpolation (3-4 entries per row): 4 rows per thread keep enough gathers
      // in flight.  Restriction (~30 entries per row): one row per thread, a thread's rows would
      // otherwise touch R distant pieces of x and thrash L1 (measured: profiles/r1_level_sweep_27pt.txt)
      if (M.avg_row_nnz < 8.0) {
         return big ? pat_launch_t<EPI, true, 512, 4>(M, x, ea, st) : pat_launch_t<EPI, true, 256, 4>(M, x, ea, st);
      }
      return big ? pat_launch_t<EPI, true, 512, 1>(M, x, ea, st) : pat_launch_t<EPI, true, 256, 1>(M, x, ea, st);
   }
   if (pat_rows_per_thread() == 8) {
      return big ? pat_launch_t<EPI, false, 512, 8>(M, x, ea, st) : pat_launch_t<EPI, false, 256, 8>(M, x, ea, st);
   }
   return big ? pat_launch_t<EPI, false, 512, 4>(M, x, ea, st) : pat_launch_t<EPI, false, 256, 4>(M, x, ea, st);
}

int spmv_pat_launch(const DCsr &M, const double *x, int epi_kind, const EpiArgs &ea, cudaStream_t st)
{
   if (ea.dot_slot >= 0 && ea.dotw != nullptr) {
      if (!spmv_can_fuse_dot(M, epi_kind)) return set_error(HB200_ERROR_GENERIC, "fused dot requested on a block that cannot fuse it");
      return epi_kind == EPI_AXPBY ? pat_dispatch_dot<EPI_AXPBY>(M, x, ea, st) : pat_dispatch_dot<EPI_JACOBI7>(M, x, ea, st);
   }
   switch (epi_kind) {
      case EPI_AXPBY:           return pat_dispatch<EPI_AXPBY>(M, x, ea, st);
      case EPI_ACC:             return pat_dispatch<EPI_ACC>(M, x, ea, st);
      case EPI_JACOBI7:         return pat_dispatch<EPI_JACOBI7>(M, x, ea, st);
      case EPI_JACOBI7_ACC:     return pat_dispatch<EPI_JACOBI7_ACC>(M, x, ea, st);
      case EPI_JACOBI_CORE:     return pat_dispatch<EPI_JACOBI_CORE>(M, x, ea, st);
      case EPI_JACOBI_CORE_ACC: return pat_dispatch<EPI_JACOBI_CORE_ACC>(M, x, ea, st);
      case EPI_CHEBY_FIRST:     return pat_dispatch<EPI_CHEBY_FIRST>(M, x, ea, st);
      case EPI_CHEBY_STEP:      return pat_dispatch<EPI_CHEBY_STEP>(M, x, ea, st);
      default: return set_error(HB200_ERROR_ARG, "spmv_pat_launch: unknown epilogue %d", epi_kind);
   }
}

// =======================================================================================
// Box view of the row-pattern format: compact 3 x 3 x 3 stencils (spmv_box)
// =======================================================================================
// When every pattern of the table is a subset of the 27 offsets {dz*sz + dy*sy + dx : d in {-1,0,1}}
// for one pair of strides (sy, sz), stored diagonal first and then in ascending order — the 7-point
// and 27-point operators of `ij` on a box, with every boundary variant — a row is the 27-slot list
// (presence bit, value) and the SpMV is a stencil sweep:
//   * a thread owns one in-plane position q (< sz) and walks a run of planes z: its rows are
//     q + z*sz.  The 27 x values of row z are 9 in-plane neighbours (dy, dx) on the planes z-1, z,
//     z+1; moving to z+1 keeps 18 of them in registers and loads 9.  One row costs 9 coalesced
//     loads of x instead of 27 gathers (the L1 wavefronts that bounded spmv_pat), no table read in
//     the interior (the full pattern's 27 values are a kernel argument: constant bank operands);
//   * a warp whose 32 rows all carry the full pattern takes that path; any other row (the faces,
//     edges and corners of the box: 2 % at 256^3) reads its mask and values from shared memory and
//     runs the same 27 slots predicated;
//   * products are added in CSR order — diagonal, then ascending offsets — with separate multiply
//     and add: bit-identical to spmv_pat and to the 1-thread CPU reference (csr_matvec.c:683-721).
// Bound: HBM (17 B per row + the epilogue vectors: x is read once, the z-halo of a run comes
// from L2) and the FP64 pipe (54 instructions per row), no longer L1.
struct BoxP0 { double a[27]; };

constexpr int kBoxThreads = 256;   // in-plane positions per block (one per thread)
constexpr int kBoxRows    = 2;     // U: planes (rows per thread) computed per step — independent summation chains
constexpr int kBoxStages  = 3 * kBoxRows + 1;   // planes in the shared-memory ring per block (two steps of look-ahead)

__device__ __forceinline__ unsigned int box_smem_u32(const void *p)
{
#ifndef HB200_EMU
   return (unsigned int) __cvta_generic_to_shared(p);
#else
   return 0u;
#endif
}
// ---- the two ways a plane reaches shared memory ----------------------------------------------
// (a) TMA bulk copy (cp.async.bulk, SASS UBLKCP): ONE thread issues one instruction per contiguous
//     piece, the copy engine moves it and signals an mbarrier with the byte count; needs 16-byte
//     aligned addresses and sizes (the launcher checks the pointers, the kernel pads the segment);
// (b) per-thread 8-byte cp.async (LDGSTS): no alignment demand, ~3 instructions per thread and plane
__device__ __forceinline__ void box_cp_async8(double *smem_dst, const double *gsrc)
{
#ifndef HB200_EMU
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(box_smem_u32(smem_dst)), "l"(gsrc) : "memory");
#else
   *smem_dst = *gsrc;
#endif
}
__device__ __forceinline__ void box_cp_commit()
{
#ifndef HB200_EMU
   asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void box_cp_wait()
{
#ifndef HB200_EMU
   asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}
__device__ __forceinline__ void box_mbar_init(unsigned long long *bar, int count)
{
#ifndef HB200_EMU
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(box_smem_u32(bar)), "r"(count) : "memory");
#else
   *bar = 0;
#endif
}
__device__ __forceinline__ void box_mbar_fence_init()
{
#ifndef HB200_EMU
   asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
__device__ __forceinline__ void box_mbar_expect(unsigned long long *bar, unsigned int bytes)
{
#ifndef HB200_EMU
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(box_smem_u32(bar)), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void box_bulk_copy(void *smem_dst, const void *gsrc, unsigned int bytes, unsigned long long *bar)
{
#ifndef HB200_EMU
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(box_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(box_smem_u32(bar)) : "memory");
#else
   memcpy(smem_dst, gsrc, bytes);
#endif
}
__device__ __forceinline__ void box_mbar_wait(unsigned long long *bar, unsigned int parity)
{
#ifndef HB200_EMU
   unsigned int done = 0;
   const unsigned int a = box_smem_u32(bar);
   while (!done) {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done) : "r"(a), "r"(parity) : "memory");
   }
#endif
}

// The kernel.  A block owns kBoxThreads consecutive in-plane positions [q0, q0 + NT) and a run of
// planes; plane p of its x slab is the contiguous segment x[p*sz + q0 - sy - 1 .. p*sz + q0 + NT + sy + 1),
// brought into a ring of shared-memory stages up to two steps ahead of the compute (the same stage
// carries the per-row inputs of the epilogue: b, or f and l1).  A step computes the rows of U
// consecutive planes: a thread reads its 9 in-plane neighbours of the U NEW planes from shared memory
// into the register window (U + 2 planes; 2 stay from the step before) and runs the 27 slots of its U
// rows — U independent summation chains per thread.  The chain (27 dependent FP64 adds, in CSR order
// for bit-identity with the reference) is what one row costs in latency: with 2 blocks x 8 warps x U
// chains per SM the FP64 pipe and the memory system stay busy while every chain waits for itself.
//   STREAMS = number of epilogue vectors staged with the slab (0: none, 1: b, 2: f and l1)
//   BULK    = planes arrive by TMA bulk copies on an mbarrier per stage, else by per-thread cp.async
//   NEG1    = every off-diagonal coefficient of the reference pattern is exactly -1.0 (the Laplacians):
//             a * x is -x, bit for bit, so the rows add -x instead of multiplying
//   UNI     = every pattern of the table is the reference pattern with some slots missing (a constant-
//             coefficient stencil: the boundary rows only LOSE entries): one branch-free path for all rows,
//             the row's presence mask predicates the 27 slots, the values are kernel arguments.  Without it
//             a warp with one boundary lane reads masks and values from shared memory for all its rows, and
//             the per-step barrier makes the whole block as slow as that warp.
//   GEO     = (with UNI) the reference pattern is the full 27-point stencil and a row misses exactly the neighbours
//             that lie outside the box sy x (sz / sy) x nplanes (checked row by row at upload): the missing
//             operands are ZERO instead of the additions being predicated — an in-plane neighbour outside the
//             box reads a zero word of the stage (its offset is fixed per thread), a plane outside the box
//             loads zeros — so every row runs the 27 unconditional slots of the interior and no row code is
//             read at all.  s + a * 0 == s: same values as the predicated form (a partial sum of -0.0 may
//             become +0.0).  One instruction per slot instead of three.
template <int EPI, bool DOT, int STREAMS, bool BULK, bool NEG1, bool UNI, bool GEO = false>
__global__ void __launch_bounds__(kBoxThreads, 2)
spmv_box(int nrows, int sy, int sz, int zrun, int nplanes, int gx, const unsigned char *__restrict__ pat, int npat, int p0,
         const unsigned int *__restrict__ masks, const double *__restrict__ vals, BoxP0 P0,
         const double *__restrict__ x, EpiArgs ea)
{
   constexpr int NT = kBoxThreads, NS = kBoxStages, U = kBoxRows, NW = U + 2;
   HB_DYN_SHARED(double, s_mem);
   const int seg = NT + 2 * sy + 2;                    // doubles of one plane segment
   const int segpad = (seg + 2) & ~1;                  // + the parity shift of an aligned copy, even
   const int zslot = segpad + STREAMS * NT;            // GEO: a zero word behind the staged data of every stage
   const int stage_len = zslot + (GEO ? 2 : 0);        // + the staged epilogue vectors (+ the zero pair)
   double             *s_ring = s_mem;                                          // NS stages
   unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(s_ring + (size_t) NS * stage_len);   // NS mbarriers
   double             *s_val = reinterpret_cast<double *>(s_bar + NS);          // npat x 27
   unsigned int       *s_mask = reinterpret_cast<unsigned int *>(s_val + npat * 27);
   const int tid = threadIdx.x;
   for (int k = tid; k < npat * 27; k += NT) s_val[k] = vals[k];
   for (int k = tid; k < npat; k += NT) s_mask[k] = masks[k];
   if (GEO && tid < NS) { s_ring[(size_t) tid * stage_len + zslot] = 0.0; s_ring[(size_t) tid * stage_len + zslot + 1] = 0.0; }
   if (BULK && tid == 0) {
      for (int k = 0; k < NS; k++) box_mbar_init(s_bar + k, 1);
      box_mbar_fence_init();
   }
   __syncthreads();
   // blocks are numbered in-plane first: neighbours in q run the same planes at the same time (L2)
   const int q0 = (int) (blockIdx.x % (unsigned) gx) * NT;
   const int q = q0 + tid;                                                // in-plane position
   const int z0 = (int) (blockIdx.x / (unsigned) gx) * zrun;
   const int z1 = min(z0 + zrun, nplanes);
   const bool qok = q < sz;
   const int skip_c = (EPI == EPI_JACOBI_CORE && ea.skip_diag) ? 1 : 0;   // leave the diagonal out
   const double *st0 = (STREAMS >= 1) ? ea.b : nullptr;                   // AXPBY: b ; JACOBI7: f
   const double *st1 = (STREAMS >= 2) ? ea.d : nullptr;                   // JACOBI7: l1 norms
   // global index of xs[0] of plane p is p*sz + q0 - sy - 1 - dpar: dpar = parity of the segment start,
   // so that an aligned (even) global index lands on an even shared-memory index (BULK: sz is even,
   // the parity is the same for every plane)
   const int dpar = BULK ? ((q0 - sy - 1) & 1) : 0;
   const int nq = min(NT, sz - q0);                                       // in-plane positions of this block

   // ---- producer: plane p (x segment) and the epilogue inputs of the rows of plane p - 1 go to stage
   //      (p - z0 + 1) % NS
   auto issue = [&](int p, int stage) {
      double *dst = s_ring + (size_t) stage * stage_len;
      const long long gstart = (long long) p * sz + q0 - sy - 1;          // global index of the segment start
      if (BULK) {
         if (tid == 0) {
            unsigned long long *bar = s_bar + stage;
            // clamp to the vector, align to 16 bytes (pairs of doubles)
            long long cs = gstart < 0 ? 0 : gstart;
            long long ce = gstart + seg;
            if (ce > (long long) nrows) ce = nrows;
            long long as = cs & ~1LL, ae = (ce + 1) & ~1LL;
            bool tail = false;
            if (ae > (long long) nrows) { ae = ce - 1; tail = true; }     // odd vector length: last element by hand
            const long long r0 = (long long) (p - 1) * sz + q0;           // first row of the staged epilogue inputs
            long long rcnt = 0;
            if (STREAMS >= 1 && p - 1 >= z0 && p - 1 < z1) {
               rcnt = nq;
               if (r0 + rcnt > (long long) nrows) rcnt = (long long) nrows - r0;
               if (rcnt < 0) rcnt = 0;
            }
            const long long rev = rcnt & ~1LL;                            // even part by bulk copy
            unsigned int bytes = 0;
            if (ae > as) bytes += (unsigned int) ((ae - as) * 8);
            bytes += (unsigned int) (rev * 8 * STREAMS);
            box_mbar_expect(bar, bytes);
            if (ae > as) box_bulk_copy(dst + (as - gstart + dpar), x + as, (unsigned int) ((ae - as) * 8), bar);
            if (tail && ce - 1 >= cs) dst[ce - 1 - gstart + dpar] = x[ce - 1];
            if (STREAMS >= 1 && rev > 0) {
               box_bulk_copy(dst + segpad, st0 + r0, (unsigned int) (rev * 8), bar);
               if (STREAMS >= 2) box_bulk_copy(dst + segpad + NT, st1 + r0, (unsigned int) (rev * 8), bar);
            }
            if (STREAMS >= 1 && rcnt > rev) {
               dst[segpad + rev] = st0[r0 + rev];
               if (STREAMS >= 2) dst[segpad + NT + rev] = st1[r0 + rev];
            }
         }
      } else {
         for (int k = tid; k < seg; k += NT) {
            const long long gi = gstart + k;
            if (gi >= 0 && gi < (long long) nrows) box_cp_async8(dst + k, x + gi);
         }
         if (STREAMS >= 1) {
            const long long r = (long long) (p - 1) * sz + q;
            if (p - 1 >= z0 && p - 1 < z1 && qok && r < (long long) nrows) {
               box_cp_async8(dst + segpad + tid, st0 + r);
               if (STREAMS >= 2) box_cp_async8(dst + segpad + NT + tid, st1 + r);
            }
         }
         box_cp_commit();
      }
   };
   // planes are issued in order, one cp.async group / one mbarrier phase each
   int next_plane = z0 - 1, next_stage = 0;          // the next plane to issue and its stage
   auto issue_upto = [&](int last) {
      for (; next_plane <= last; next_plane++) {
         issue(next_plane, next_stage);
         if (++next_stage == NS) next_stage = 0;
      }
   };
   // prologue: planes z0 - 1 .. z0 + NS - 2 fill the ring
   issue_upto(z0 + NS - 2);
#ifdef HB200_EMU
   __syncthreads();   // (the emulated copies are synchronous, made by thread 0: let it run first)
#endif
   // consumer side: stage / use count of the next plane to take out of the ring (planes leave in order)
   int rd_stage = 0, rd_use = 0;
   auto landed_bulk = [&]() { box_mbar_wait(s_bar + rd_stage, (unsigned int) (rd_use & 1)); };
   auto advance = [&]() { if (++rd_stage == NS) { rd_stage = 0; rd_use++; } };
   // in-plane neighbour offsets inside a segment, class c = (dy+1)*3 + (dx+1)
   int offc[9];
#pragma unroll
   for (int c = 0; c < 9; c++) offc[c] = tid + sy + 1 + dpar + (c / 3 - 1) * sy + (c % 3 - 1);
   if (GEO) {
      // in-plane neighbours outside the box: the same for every plane of this thread
      const int ny = sz / sy, xq = q % sy, yq = q / sy;
#pragma unroll
      for (int c = 0; c < 9; c++) {
         const bool in = qok && (unsigned int) (xq + c % 3 - 1) < (unsigned int) sy && (unsigned int) (yq + c / 3 - 1) < (unsigned int) ny;
         if (!in) offc[c] = zslot;
      }
   }
   double W[NW][9];                                                       // planes z-1 .. z+U (rotating roles)
   // planes z0 - 1 and z0 into the window (cp.async groups complete in order: the NS - 2 planes issued
   // after them may still be in flight)
   if (BULK) { landed_bulk(); advance(); landed_bulk(); advance(); }
   else      { box_cp_wait<NS - 2>(); advance(); advance(); }
   __syncthreads();   // (cp.async data of the other threads; the hand-copied tail elements of thread 0)
   {
      const double *sm = s_ring, *sc = s_ring + stage_len;
#pragma unroll
      for (int c = 0; c < 9; c++) { W[0][c] = (GEO && z0 == 0) ? 0.0 : sm[offc[c]]; W[1][c] = sc[offc[c]]; }
   }
   // row codes are prefetched two steps ahead in registers (32-bit row arithmetic: rows are ints)
   const unsigned int urows = (unsigned int) nrows, usz = (unsigned int) sz;
   unsigned int rowq = (unsigned int) z0 * usz + (unsigned int) q;        // this thread's row in plane z
   auto ldcode = [&](unsigned int r, int z) -> int {
      if (GEO) return (qok && z < z1 && r < urows) ? 0 : 255;   // every row of the box is a table row: nothing to read
      return (qok && z < z1 && r < urows) ? (int) __ldg(pat + r) : 255;
   };
   int code_n1[U], code_n2[U];
#pragma unroll
   for (int i = 0; i < U; i++) {
      code_n1[i] = ldcode(rowq + (unsigned int) i * usz, z0 + i);
      code_n2[i] = ldcode(rowq + (unsigned int) (U + i) * usz, z0 + U + i);
   }
   double dacc = 0.0;
   // the window rotates by U planes per step: its roles repeat every NW / gcd(U, NW) steps
   constexpr int PERIOD = (NW % U == 0) ? NW / U : NW;
   for (int zb = z0; zb < z1; zb += PERIOD * U) {
#pragma unroll
      for (int ph = 0; ph < PERIOD; ph++) {
         const int z = zb + ph * U;
         if (z >= z1) break;                              // block-uniform
         int code[U];
#pragma unroll
         for (int i = 0; i < U; i++) { code[i] = code_n1[i]; code_n1[i] = code_n2[i]; }
#pragma unroll
         for (int i = 0; i < U; i++) code_n2[i] = ldcode(rowq + (unsigned int) (2 * U + i) * usz, z + 2 * U + i);
         // ---- the U new planes z + 1 .. z + U: shared memory -> register window
         double e0[U], e1[U];
         if (!BULK) {
            // they have landed when only the planes issued after them are pending (cp.async groups complete in
            // order): the first step sees the prologue's planes up to z0 + NS - 2, the others up to z + NS
            if (z == z0) box_cp_wait<NS - 2 - U>(); else box_cp_wait<NS - U>();
            __syncthreads();                              // (everybody's part of them)
         }
#pragma unroll
         for (int i = 0; i < U; i++) {
            double (&Wn)[9] = W[(ph * U + 2 + i) % NW];   // plane z + 1 + i: role (ph*U + 2 + i) % NW
            if (BULK) landed_bulk();
            const double *sp = s_ring + (size_t) rd_stage * stage_len;
#pragma unroll
            for (int c = 0; c < 9; c++) Wn[c] = (GEO && z + 1 + i >= nplanes) ? 0.0 : sp[offc[c]];   // (block-uniform)
            e0[i] = (STREAMS >= 1) ? sp[segpad + tid] : 0.0;
            e1[i] = (STREAMS >= 2) ? sp[segpad + NT + tid] : 0.0;
            advance();
         }
         // every thread has the planes <= z + U of the ring in its registers: their stages are refilled
         // (thread 0; it overlaps with the other warps' arithmetic below) up to plane z + U + NS
         __syncthreads();
         issue_upto(z + U + NS);
         // ---- the U rows of this step.  Interior warps: U independent chains in ONE basic block, so that the
         // compiler interleaves them (a chain is 27 dependent FP64 adds: latency, not throughput)
         double s[U];
         bool all_full = true;
#pragma unroll
         for (int i = 0; i < U; i++) all_full = all_full && (code[i] == p0);
         if (GEO || (!UNI && __all_sync(0xffffffffu, all_full))) {
#pragma unroll
            for (int i = 0; i < U; i++) {
               double (&Wc)[9] = W[(ph * U + i + 1) % NW];
               s[i] = 0.0;
               if (!skip_c) s[i] = __dadd_rn(s[i], __dmul_rn(P0.a[13], Wc[4]));
            }
#pragma unroll
            for (int t = 0; t < 27; t++) {
               if (t == 13) continue;
#pragma unroll
               for (int i = 0; i < U; i++) {
                  double (&Wm)[9] = W[(ph * U + i) % NW];
                  double (&Wc)[9] = W[(ph * U + i + 1) % NW];
                  double (&Wp)[9] = W[(ph * U + i + 2) % NW];
                  const double w = t < 9 ? Wm[t] : t < 18 ? Wc[t - 9] : Wp[t - 18];
                  if (NEG1) s[i] = __dadd_rn(s[i], -w);     // (-1.0) * w == -w exactly
                  else      s[i] = __dadd_rn(s[i], __dmul_rn(P0.a[t], w));
               }
            }
         } else if (UNI) {
            unsigned int m[U];
#pragma unroll
            for (int i = 0; i < U; i++) {
               double (&Wc)[9] = W[(ph * U + i + 1) % NW];
               m[i] = code[i] != 255 ? s_mask[code[i]] : 0u;
               s[i] = 0.0;
               if (!skip_c && (m[i] & (1u << 13))) s[i] = __dadd_rn(s[i], __dmul_rn(P0.a[13], Wc[4]));
            }
#pragma unroll
            for (int t = 0; t < 27; t++) {
               if (t == 13) continue;
#pragma unroll
               for (int i = 0; i < U; i++) {
                  double (&Wm)[9] = W[(ph * U + i) % NW];
                  double (&Wc)[9] = W[(ph * U + i + 1) % NW];
                  double (&Wp)[9] = W[(ph * U + i + 2) % NW];
                  const double w = t < 9 ? Wm[t] : t < 18 ? Wc[t - 9] : Wp[t - 18];
                  if (m[i] & (1u << t)) {
                     if (NEG1) s[i] = __dadd_rn(s[i], -w);  // (-1.0) * w == -w exactly
                     else      s[i] = __dadd_rn(s[i], __dmul_rn(P0.a[t], w));
                  }
               }
            }
         } else {
#pragma unroll
            for (int i = 0; i < U; i++) {
               double (&Wm)[9] = W[(ph * U + i) % NW];
               double (&Wc)[9] = W[(ph * U + i + 1) % NW];
               double (&Wp)[9] = W[(ph * U + i + 2) % NW];
               double acc = 0.0;
               if (code[i] != 255) {
                  const unsigned int m = s_mask[code[i]];
                  const double *a = s_val + code[i] * 27;
                  if (!skip_c && (m & (1u << 13))) acc = __dadd_rn(acc, __dmul_rn(a[13], Wc[4]));
#pragma unroll
                  for (int t = 0; t < 27; t++) {
                     if (t == 13) continue;
                     const double w = t < 9 ? Wm[t] : t < 18 ? Wc[t - 9] : Wp[t - 18];
                     if (m & (1u << t)) acc = __dadd_rn(acc, __dmul_rn(a[t], w));
                  }
               }
               s[i] = acc;
            }
         }
#pragma unroll
         for (int i = 0; i < U; i++) {
            if (code[i] != 255) {                         // (255: past the end, or a row of the CSR pass)
               const int r = (int) (rowq + (unsigned int) i * usz);
               double (&Wc)[9] = W[(ph * U + i + 1) % NW];
               double v;
               // the epilogues of hb_epilogue.cuh with their per-row inputs already at hand
               if (EPI == EPI_AXPBY) {
                  v = (STREAMS == 0) ? ea.alpha * s[i] : epi_axpby_value(ea, e0[i], s[i]);
                  __stcs(ea.y + r, v);
               } else if (EPI == EPI_JACOBI7) {
                  const double uo = Wc[4];                // u_in is the vector the sweep multiplies
                  if (ea.cf == nullptr || __ldcs(ea.cf + r) == ea.relax_points) v = epi_jacobi7_value(ea, uo, e0[i], e1[i], s[i]);
                  else v = uo;
                  __stcs(ea.y + r, v);
               } else {
                  epi_apply<EPI>(ea, r, s[i], (UNI || code[i] == p0) ? P0.a[13] : s_val[code[i] * 27 + 13]);
                  v = 0.0;
               }
               if (DOT) dacc += v * __ldg(ea.dotw + r);
            }
         }
         rowq += (unsigned int) U * usz;
      }
   }
   // the planes issued beyond the run are still in flight: let them land before the block leaves
   __syncthreads();
   if (BULK) {
      for (int k = rd_use * NS + rd_stage; k < next_plane - (z0 - 1); k++) { landed_bulk(); advance(); }
   } else {
      box_cp_wait<0>();
   }
   if (DOT) pat_dot_finish<NT>(dacc, ea.dot_slot);
}

bool spmv_box_supports(int epi_kind) { return epi_kind == EPI_AXPBY || epi_kind == EPI_JACOBI7 || epi_kind == EPI_JACOBI_CORE; }

static int box_zrun()
{
   static int z = 0;
   if (z == 0) {
      const char *e = getenv("HB200_BOX_ZRUN");
      z = e ? atoi(e) : 0;
      if (z < 2) z = 0; else z = (z + 3) / 4 * 4;
      if (z == 0) z = -1;
   }
   return z;
}

static size_t box_smem_bytes(const DCsr &M, int streams)
{
   const size_t seg = (size_t) kBoxThreads + 2 * (size_t) M.box_sy + 2;
   const size_t segpad = (seg + 2) & ~(size_t) 1;
   return sizeof(double) * ((size_t) kBoxStages * (segpad + (size_t) streams * kBoxThreads + 2) + kBoxStages + (size_t) M.pat_npat * 27) +
          sizeof(unsigned int) * (size_t) M.pat_npat + 16;
}

// the slab ring has to fit one block's shared memory: in-plane strides up to ~4000 (a 4000-wide grid)
static bool box_fits(const DCsr &M) { return box_smem_bytes(M, 2) <= 200 * 1024; }

template <int EPI, bool DOT, int STREAMS, bool BULK, bool NEG1, bool UNI, bool GEO = false>
static int box_launch_v(const DCsr &M, const double *x, const EpiArgs &ea, cudaStream_t st)
{
   const size_t smem = box_smem_bytes(M, STREAMS);
   static size_t opted = 0;
   if (opted < smem) {
      HB_CUDA(cudaFuncSetAttribute(spmv_box<EPI, DOT, STREAMS, BULK, NEG1, UNI, GEO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      opted = smem;
   }
   if (DOT) {
      static bool bound = false;
      if (!bound) {
         Ctx &c = ctx();
         HB_CUDA(cudaMemcpyToSymbol(g_pd_partials, &c.d_partials, sizeof(double *)));
         HB_CUDA(cudaMemcpyToSymbol(g_pd_counter, &c.d_counter, sizeof(unsigned int *)));
         HB_CUDA(cudaMemcpyToSymbol(g_pd_scalars, &c.d_scalars, sizeof(double *)));
         bound = true;
      }
   }
   const int nplanes = (int) (((long long) M.nrows + M.box_sz - 1) / M.box_sz);
   const int gx = (M.box_sz + kBoxThreads - 1) / kBoxThreads;
   // planes per block: long runs amortise the pipeline fill, short ones fill the GPU: aim at >= 3
   // waves of two blocks per SM
   int zrun = box_zrun();
   if (zrun < 0) {
      zrun = 48;
      while (zrun > 8 && (long long) gx * ((nplanes + zrun - 1) / zrun) < 3LL * 2 * kNumSMs) zrun -= 8;
   }
   int gy = (nplanes + zrun - 1) / zrun;
   if (DOT && (long long) gx * gy > kRedBlocksMax) {
      // the fused dot keeps one partial per block: fewer, longer runs
      gy = kRedBlocksMax / gx; if (gy < 1) gy = 1;
      zrun = ((nplanes + gy - 1) / gy + 3) / 4 * 4;
      gy = (nplanes + zrun - 1) / zrun;
      if ((long long) gx * gy > kRedBlocksMax) return set_error(HB200_ERROR_GENERIC, "spmv_box: fused dot on a plane of more than %d blocks", kRedBlocksMax);
   }
   BoxP0 P0;
   for (int t = 0; t < 27; t++) P0.a[t] = M.box_p0_val[t];
   HB_LAUNCH((spmv_box<EPI, DOT, STREAMS, BULK, NEG1, UNI, GEO>), gx * gy, kBoxThreads, smem, st, M.nrows, M.box_sy, M.box_sz, zrun, nplanes, gx, M.pat_code,
             M.pat_npat, M.box_p0, M.box_mask, M.box_val, P0, x, ea);
   HB_LAUNCH_CHECK();
   return 0;
}

template <int EPI, bool DOT, int STREAMS>
static int box_launch_t(const DCsr &M, const double *x, const EpiArgs &ea, cudaStream_t st)
{
   // TMA bulk copies need 16-byte aligned pieces: an even plane size and aligned vectors (every cudaMalloc'ed
   // vector is; a GMRES basis vector inside a slab of odd-length vectors is not)
   auto al16 = [](const void *p) { return (((uintptr_t) p) & 15u) == 0; };
   static const bool no_bulk = env_flag("HB200_BOX_NO_BULK", false);
   const bool bulk = !no_bulk && (M.box_sz % 2 == 0) && al16(x) && (STREAMS < 1 || al16(ea.b)) && (STREAMS < 2 || al16(ea.d));
   // the Laplacians: every off-diagonal coefficient of the reference pattern is exactly -1
   bool neg1 = M.box_p0 >= 0 || M.box_uniform;
   for (int t = 0; t < 27 && neg1; t++) if (t != 13 && (M.box_ref_mask & (1u << t)) && M.box_p0_val[t] != -1.0) neg1 = false;
   static const bool no_neg1 = env_flag("HB200_BOX_NO_NEG1", false);
   if (no_neg1) neg1 = false;
   static const bool no_uni = env_flag("HB200_BOX_NO_UNI", false);
   const bool uni = M.box_uniform && !no_uni;
   static const bool no_geo = env_flag("HB200_BOX_NO_GEO", false);
   if (uni && M.box_geo && !no_geo) {
      if (bulk) return neg1 ? box_launch_v<EPI, DOT, STREAMS, true, true, true, true>(M, x, ea, st) : box_launch_v<EPI, DOT, STREAMS, true, false, true, true>(M, x, ea, st);
      return neg1 ? box_launch_v<EPI, DOT, STREAMS, false, true, true, true>(M, x, ea, st) : box_launch_v<EPI, DOT, STREAMS, false, false, true, true>(M, x, ea, st);
   }
   if (uni) {
      if (bulk) return neg1 ? box_launch_v<EPI, DOT, STREAMS, true, true, true>(M, x, ea, st) : box_launch_v<EPI, DOT, STREAMS, true, false, true>(M, x, ea, st);
      return neg1 ? box_launch_v<EPI, DOT, STREAMS, false, true, true>(M, x, ea, st) : box_launch_v<EPI, DOT, STREAMS, false, false, true>(M, x, ea, st);
   }
   if (M.box_p0 < 0) return spmv_pat_launch(M, x, EPI, ea, st);   // no full pattern and no uniform table: the generic kernel
   if (bulk) return neg1 ? box_launch_v<EPI, DOT, STREAMS, true, true, false>(M, x, ea, st) : box_launch_v<EPI, DOT, STREAMS, true, false, false>(M, x, ea, st);
   return neg1 ? box_launch_v<EPI, DOT, STREAMS, false, true, false>(M, x, ea, st) : box_launch_v<EPI, DOT, STREAMS, false, false, false>(M, x, ea, st);
}

int spmv_box_launch(const DCsr &M, const double *x, int epi_kind, const EpiArgs &ea, cudaStream_t st)
{
   const bool dot = ea.dot_slot >= 0 && ea.dotw != nullptr;
   if (dot && !spmv_can_fuse_dot(M, epi_kind)) return set_error(HB200_ERROR_GENERIC, "fused dot requested on a block that cannot fuse it");
   switch (epi_kind) {
      case EPI_AXPBY:
         if (ea.beta == 0.0) return dot ? box_launch_t<EPI_AXPBY, true, 0>(M, x, ea, st) : box_launch_t<EPI_AXPBY, false, 0>(M, x, ea, st);
         return dot ? box_launch_t<EPI_AXPBY, true, 1>(M, x, ea, st) : box_launch_t<EPI_AXPBY, false, 1>(M, x, ea, st);
      case EPI_JACOBI7:
         if (ea.u != x) return spmv_pat_launch(M, x, epi_kind, ea, st);   // (the sweep always multiplies the vector it updates)
         return dot ? box_launch_t<EPI_JACOBI7, true, 2>(M, x, ea, st) : box_launch_t<EPI_JACOBI7, false, 2>(M, x, ea, st);
      case EPI_JACOBI_CORE: return box_launch_t<EPI_JACOBI_CORE, false, 0>(M, x, ea, st);
      default: return set_error(HB200_ERROR_ARG, "spmv_box_launch: epilogue %d has no box kernel", epi_kind);
   }
}

// Host side: does the pattern table describe compact stencils?  Finds the strides (sy, sz) for which
// every offset of every pattern is dz*sz + dy*sy + dx with d in {-1,0,1} (the decomposition has to
// be unique: sy >= 3, sz >= 2*sy + 3), checks the storage order (diagonal first, then ascending
// slots: the order `ij`'s generators and hypre's IJ assembly produce), and lays every pattern out
// as 27 (presence, value) slots.  Pure host code.
struct BoxHost {
   bool ok = false, uniform = false;
   int sy = 0, sz = 0, p0 = -1, pref = -1;    // p0: the full (27-slot) pattern; pref: the reference pattern (most slots)
   unsigned int ref_mask = 0;
   std::vector<unsigned int> mask;
   std::vector<double> val;
};

static bool box_slot_of(int off, int sy, int sz, int *slot)
{
   for (int dz = -1; dz <= 1; dz++) {
      for (int dy = -1; dy <= 1; dy++) {
         const int dx = off - dz * sz - dy * sy;
         if (dx >= -1 && dx <= 1) { *slot = (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1); return true; }
      }
   }
   return false;
}

static void box_analyze_host(const PatHost &ph, int nrows, BoxHost &out)
{
   out = BoxHost();
   if (!ph.ok || !ph.square || ph.wide) return;
   const int npat = (int) ph.ptr.size() - 1;
   if (npat < 1) return;
   // candidate strides from the positive offsets of the table
   std::vector<int> pos;
   for (int o : ph.off) if (o > 1) pos.push_back(o);
   std::sort(pos.begin(), pos.end());
   pos.erase(std::unique(pos.begin(), pos.end()), pos.end());
   if (pos.empty() || pos.size() > 16) return;
   std::vector<int> cand;
   for (int o : pos) { for (int d = -1; d <= 1; d++) if (o + d >= 3) cand.push_back(o + d); }
   std::sort(cand.begin(), cand.end());
   cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
   int best_sy = 0, best_sz = 0;
   for (size_t i = 0; i < cand.size() && !best_sy; i++) {
      for (size_t j = 0; j < cand.size() && !best_sy; j++) {
         const int sy = cand[i], sz = cand[j];
         if (sz < 2 * sy + 3) continue;
         bool all = true;
         int slot;
         for (int o : ph.off) { if (!box_slot_of(o, sy, sz, &slot)) { all = false; break; } }
         if (all) { best_sy = sy; best_sz = sz; }
      }
   }
   // a one- or two-dimensional operator (no offset beyond the row / the plane) is a box with a
   // stride nobody reaches
   if (!best_sy) return;
   if ((long long) best_sz > (long long) nrows) return;
   out.sy = best_sy; out.sz = best_sz;
   out.mask.assign((size_t) npat, 0u);
   out.val.assign((size_t) npat * 27, 0.0);
   for (int p = 0; p < npat; p++) {
      int prev = -1;
      for (int k = ph.ptr[p]; k < ph.ptr[p + 1]; k++) {
         int slot = -1;
         if (!box_slot_of(ph.off[k], best_sy, best_sz, &slot)) return;
         if (k == ph.ptr[p]) {
            if (slot != 13) return;                       // the diagonal comes first (par_relax.c:274)
         } else {
            if (slot == 13 || slot <= prev) return;       // then ascending
            prev = slot;
         }
         if (out.mask[(size_t) p] & (1u << slot)) return;
         out.mask[(size_t) p] |= 1u << slot;
         out.val[(size_t) p * 27 + slot] = ph.val[(size_t) k];
      }
      if (out.mask[(size_t) p] == 0x7ffffffu && out.p0 < 0) out.p0 = p;   // patterns are sorted by frequency
   }
   // uniform table: every pattern is the reference pattern (the one with the most slots) minus some slots
   int best = 0;
   for (int p = 1; p < npat; p++) if (__builtin_popcount(out.mask[(size_t) p]) > __builtin_popcount(out.mask[(size_t) best])) best = p;
   out.pref = best;
   out.ref_mask = out.mask[(size_t) best];
   out.uniform = true;
   for (int p = 0; p < npat && out.uniform; p++) {
      if (out.mask[(size_t) p] & ~out.ref_mask) { out.uniform = false; break; }
      for (int t = 0; t < 27; t++) {
         if ((out.mask[(size_t) p] & (1u << t)) &&
             memcmp(&out.val[(size_t) p * 27 + t], &out.val[(size_t) best * 27 + t], sizeof(double)) != 0) { out.uniform = false; break; }
      }
   }
   out.ok = true;
}

// GEO (spmv_box): the table is uniform with the full 27-point reference pattern, the rows fill the box
// sy x (sz / sy) x (nrows / sz) exactly, and the pattern of every row is the reference pattern minus the
// neighbours outside that box.  One pass over the row codes.
static bool box_geo_check(const PatHost &ph, const BoxHost &bh, int n)
{
   if (!bh.ok || !bh.uniform || bh.ref_mask != 0x7ffffffu || !ph.irr.empty() || ph.skips_boundary) return false;
   const int sy = bh.sy, sz = bh.sz;
   if (sy < 1 || sz % sy != 0 || n % sz != 0 || (int) ph.code.size() != n) return false;
   const int ny = sz / sy, nz = n / sz;
   // masks of the 64 combinations of "first / last" along x, y, z
   unsigned int tab[64];
   for (int f = 0; f < 64; f++) {
      unsigned int m = 0;
      for (int t = 0; t < 27; t++) {
         const int dx = t % 3 - 1, dy = (t / 3) % 3 - 1, dz = t / 9 - 1;
         const bool out = (dx < 0 && (f & 1)) || (dx > 0 && (f & 2)) || (dy < 0 && (f & 4)) || (dy > 0 && (f & 8)) ||
                          (dz < 0 && (f & 16)) || (dz > 0 && (f & 32));
         if (!out) m |= 1u << t;
      }
      tab[f] = m;
   }
   const unsigned char *code = ph.code.data();
   const int npat = (int) bh.mask.size();
   size_t r = 0;
   for (int z = 0; z < nz; z++) {
      const int fz = (z == 0 ? 16 : 0) | (z == nz - 1 ? 32 : 0);
      for (int y = 0; y < ny; y++) {
         const int fy = fz | (y == 0 ? 4 : 0) | (y == ny - 1 ? 8 : 0);
         for (int x = 0; x < sy; x++, r++) {
            const int f = fy | (x == 0 ? 1 : 0) | (x == sy - 1 ? 2 : 0);
            const int c = code[r];
            if (c >= npat || bh.mask[(size_t) c] != tab[f]) return false;
         }
      }
   }
   return true;
}

int dcsr_free_pat(DCsr &M)
{
   if (M.pat_code) cudaFree(M.pat_code);
   if (M.pat_ptr) cudaFree(M.pat_ptr);
   if (M.pat_off) cudaFree(M.pat_off);
   if (M.pat_val) cudaFree(M.pat_val);
   if (M.pat_base) cudaFree(M.pat_base);
   if (M.pat_irr) cudaFree(M.pat_irr);
   if (M.box_mask) cudaFree(M.box_mask);
   if (M.box_val) cudaFree(M.box_val);
   M.box_mask = nullptr; M.box_val = nullptr; M.has_box = false; M.box_geo = false;
   M.pat_code = nullptr; M.pat_ptr = nullptr; M.pat_off = nullptr; M.pat_val = nullptr; M.pat_base = nullptr; M.pat_irr = nullptr;
   M.pat_nirr = 0;
   M.pat_wide = false;
   M.pat_skips_boundary = false;
   M.has_pat = false;
   return 0;
}

// Pattern detection on the host, one pass over the block.  Consecutive rows of a grid operator
// nearly always repeat the previous row's pattern, so the common case is one compare per entry
// against that pattern; only a change of pattern goes through the hash.  The 255 most frequent
// patterns that fit the table are kept; rows outside them (the irregular strip next to a rank
// boundary, where the coarsening of a partitioned grid loses its regularity) get code 255 and are
// swept by the CSR vector kernel over a row list (pat_irr).  The block qualifies when the table
// covers at least 70% of the rows.  Irregular blocks leave after their first 64K rows.
// Pure host code (no CUDA call): also reachable through hb200_host_pattern_analyze for CPU tests.
// rows with entries in the offd block of the ParCSR matrix this diag block belongs to (set around the upload
// of a diag block on N > 1 ranks, parcsr.cu): they are left out of the table AND of the irregular-row list —
// the boundary kernel of the split ParCSR operation computes them completely (kernels_offd.cu)
const int *g_pat_boundary_offd_i = nullptr;

int pat_analyze_host(int n, int ncols, const int *hi, const int *hj, const double *ha, PatHost &out, bool wide)
{
   const int *bnd = g_pat_boundary_offd_i;
   // narrow: 1-byte codes, table in shared memory; wide: 2-byte codes, table in global memory
   const int max_pat = wide ? 65534 : kPatMaxPatterns - 1;
   const int max_ent = wide ? (1 << 20) : kPatMaxEntries;
   const int kMaxCand = wide ? (1 << 18) : 16384;
   out = PatHost();
   const long long nnz = n > 0 ? hi[n] : 0;
   if (n < 1024 || nnz < 2LL * n) return 0;          // tiny or nearly empty (offd) blocks: nothing to win
   // square blocks: column = row + offset; rectangular ones (P, P^T): column = first column + offset
   const bool square = (n == ncols);
   std::vector<int> basev;
   if (!square) {
      basev.resize((size_t) n);
      for (int r = 0; r < n; r++) basev[r] = hi[r + 1] > hi[r] ? hj[hi[r]] : 0;
   }
   auto base_of = [&](int r) -> int { return square ? r : basev[r]; };
   // candidate patterns, each known by a representative row
   std::vector<int> rep;
   std::vector<long long> count;
   std::vector<int> rid((size_t) n, -1);
   std::unordered_multimap<unsigned long long, int> by_hash;
   auto same_rows = [&](int a, int r) -> bool {
      const int len = hi[a + 1] - hi[a];
      if (hi[r + 1] - hi[r] != len) return false;
      const int *ja = hj + hi[a], *jr = hj + hi[r];
      const double *va = ha + hi[a], *vr = ha + hi[r];
      const int shift = base_of(r) - base_of(a);
      for (int k = 0; k < len; k++) {
         if (jr[k] - ja[k] != shift || memcmp(&va[k], &vr[k], 8) != 0) return false;   // bit patterns
      }
      return true;
   };
   int prev = -1;
   long long misses = 0;
   long long nbnd = 0;
   auto row_hash = [&](int r) -> unsigned long long {
      unsigned long long h = 1469598103934665603ull ^ (unsigned long long) (hi[r + 1] - hi[r]);
      for (int q = hi[r]; q < hi[r + 1]; q++) {
         unsigned long long bits;
         memcpy(&bits, &ha[q], 8);
         h = (h ^ (unsigned long long) (long long) (hj[q] - base_of(r))) * 1099511628211ull;
         h = (h ^ bits) * 1099511628211ull;
      }
      return h;
   };
   // ---- big blocks: the classification is one memory-bound pass over the block (0.63 s for A_0 of 27-pt 256^3 on one
   // core): rows are cut into contiguous chunks, every thread classifies its chunk against its own candidate list,
   // the lists are merged in chunk order.  Candidates are numbered by first appearance either way, so the result is the
   // sequential one; a block that turns out irregular or overflows the candidate list is left to the sequential pass.
   bool classified = false;
   int nthreads = 1;
   if ((long long) n >= (1LL << 18)) {
      const char *e = getenv("HB200_ANALYSIS_THREADS");
      nthreads = e ? atoi(e) : (int) std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
      if (nthreads < 1) nthreads = 1;
      if (nthreads > 64) nthreads = 64;
   }
   if (nthreads > 1) {
      struct Local {
         std::vector<int> rep;
         std::vector<long long> count;
         std::vector<unsigned long long> hash;
         long long nbnd = 0;
         bool bad = false;
      };
      std::vector<Local> loc((size_t) nthreads);
      std::atomic<bool> abort_all(false);
      auto work = [&](int t) {
         Local &L = loc[(size_t) t];
         const int r0 = (int) ((long long) n * t / nthreads), r1 = (int) ((long long) n * (t + 1) / nthreads);
         std::unordered_multimap<unsigned long long, int> local_hash;
         int lprev = -1;
         long long lmiss = 0;
         for (int r = r0; r < r1; r++) {
            if (((r - r0) & 65535) == 0 && r > r0 && (lmiss > (r - r0) / 2 || abort_all.load(std::memory_order_relaxed))) { L.bad = true; break; }
            if (bnd && bnd[r + 1] > bnd[r]) { rid[r] = -2; L.nbnd++; continue; }
            if (lprev >= 0 && same_rows(L.rep[(size_t) lprev], r)) { rid[r] = lprev; L.count[(size_t) lprev]++; continue; }
            const unsigned long long h = row_hash(r);
            int found = -1;
            auto range = local_hash.equal_range(h);
            for (auto it = range.first; it != range.second; ++it) {
               if (same_rows(L.rep[(size_t) it->second], r)) { found = it->second; break; }
            }
            if (found < 0) {
               if ((int) L.rep.size() >= kMaxCand) { lmiss++; L.bad = true; continue; }
               found = (int) L.rep.size();
               L.rep.push_back(r);
               L.count.push_back(0);
               L.hash.push_back(h);
               local_hash.emplace(h, found);
            }
            rid[r] = found;
            L.count[(size_t) found]++;
            lprev = found;
         }
         if (L.bad) abort_all.store(true, std::memory_order_relaxed);
      };
      {
         std::vector<std::thread> th;
         for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
         work(0);
         for (auto &x : th) x.join();
      }
      bool ok = !abort_all.load();
      std::vector<std::vector<int>> remap((size_t) nthreads);
      for (int t = 0; ok && t < nthreads; t++) {
         Local &L = loc[(size_t) t];
         remap[(size_t) t].resize(L.rep.size());
         for (size_t k = 0; k < L.rep.size(); k++) {
            int found = -1;
            auto range = by_hash.equal_range(L.hash[k]);
            for (auto it = range.first; it != range.second; ++it) {
               if (same_rows(rep[(size_t) it->second], L.rep[k])) { found = it->second; break; }
            }
            if (found < 0) {
               if ((int) rep.size() >= kMaxCand) { ok = false; break; }
               found = (int) rep.size();
               rep.push_back(L.rep[k]);
               count.push_back(0);
               by_hash.emplace(L.hash[k], found);
            }
            remap[(size_t) t][k] = found;
            count[(size_t) found] += L.count[k];
         }
         nbnd += L.nbnd;
      }
      if (ok) {
         auto fix = [&](int t) {
            const int r0 = (int) ((long long) n * t / nthreads), r1 = (int) ((long long) n * (t + 1) / nthreads);
            const std::vector<int> &m = remap[(size_t) t];
            for (int r = r0; r < r1; r++) if (rid[r] >= 0) rid[r] = m[(size_t) rid[r]];
         };
         std::vector<std::thread> th;
         for (int t = 1; t < nthreads; t++) th.emplace_back(fix, t);
         fix(0);
         for (auto &x : th) x.join();
         classified = true;
      } else {
         // the sequential pass decides (and numbers) everything itself
         rep.clear(); count.clear(); by_hash.clear(); nbnd = 0;
         std::fill(rid.begin(), rid.end(), -1);
      }
   }
   for (int r = 0; !classified && r < n; r++) {
      if ((r & 65535) == 0 && r > 0 && misses > r / 2) return 0;      // irregular block
      if (bnd && bnd[r + 1] > bnd[r]) { rid[r] = -2; nbnd++; continue; }   // a boundary row: not this format's business
      if (prev >= 0 && same_rows(rep[prev], r)) { rid[r] = prev; count[prev]++; continue; }
      const unsigned long long h = row_hash(r);
      int found = -1;
      auto range = by_hash.equal_range(h);
      for (auto it = range.first; it != range.second; ++it) {
         if (same_rows(rep[it->second], r)) { found = it->second; break; }
      }
      if (found < 0) {
         if ((int) rep.size() >= kMaxCand) { misses++; continue; }
         found = (int) rep.size();
         rep.push_back(r);
         count.push_back(0);
         by_hash.emplace(h, found);
      }
      rid[r] = found;
      count[found]++;
      prev = found;
   }
   // ---- keep the most frequent patterns that fit the table
   std::vector<int> order(rep.size());
   for (size_t k = 0; k < order.size(); k++) order[k] = (int) k;
   std::sort(order.begin(), order.end(), [&](int a, int b) { return count[a] != count[b] ? count[a] > count[b] : a < b; });
   std::vector<int> code_of(rep.size(), -1);
   out.ptr.assign(1, 0);
   long long covered = 0, all_entries = 0;
   for (int c : order) all_entries += hi[rep[c] + 1] - hi[rep[c]];
   // a fully regular block keeps every pattern, rare ones included; otherwise a pattern has to
   // earn its table slot (shared memory per block) with a few rows
   const bool all_fit = ((int) rep.size() <= max_pat && all_entries <= max_ent && misses == 0);
   // (wide: a one-off row in the table is a 1-thread CSR row; the vector kernel sweeps those better)
   const long long min_count = wide ? 4 : (all_fit ? 1 : 8);
   for (int c : order) {
      const int len = hi[rep[c] + 1] - hi[rep[c]];
      if ((int) out.ptr.size() - 1 >= max_pat || count[c] < min_count) break;
      if ((int) out.off.size() + len > max_ent) continue;
      code_of[c] = (int) out.ptr.size() - 1;
      for (int q = hi[rep[c]]; q < hi[rep[c] + 1]; q++) {
         out.off.push_back(hj[q] - base_of(rep[c]));
         out.val.push_back(ha[q]);
      }
      out.ptr.push_back((int) out.off.size());
      covered += count[c];
   }
   if ((double) covered < 0.7 * (double) (n - nbnd)) { out = PatHost(); return 0; }
   if (wide) out.code16.resize((size_t) n); else out.code.resize((size_t) n);
   for (int r = 0; r < n; r++) {
      const int c = rid[r] >= 0 ? code_of[rid[r]] : -1;
      if (c < 0 && rid[r] != -2) { out.irr.push_back(r); out.irr_nnz += hi[r + 1] - hi[r]; }
      if (wide) out.code16[r] = (unsigned short) (c >= 0 ? c : 65535);
      else      out.code[r] = (unsigned char) (c >= 0 ? c : 255);
   }
   out.wide = wide;
   out.skips_boundary = (bnd != nullptr);
   out.base.swap(basev);
   out.square = square;
   out.ok = true;
   return 0;
}

int dcsr_build_pat(DCsr &M, const int *hi, const int *hj, const double *ha)
{
   if (env_flag("HB200_NO_PAT", false)) return 0;
   PatHost ph;
   HB_CHECK(pat_analyze_host(M.nrows, M.ncols, hi, hj, ha, ph, false));
   // the wide variant: on request, and by default on a partitioned problem, where the coarse
   // operators lose the regular numbering next to the rank boundary
   const bool try_wide = env_flag("HB200_PAT_WIDE", false) || (ctx().nranks > 1 && !env_flag("HB200_NO_PAT_WIDE", false));
   if (!ph.ok && try_wide) HB_CHECK(pat_analyze_host(M.nrows, M.ncols, hi, hj, ha, ph, true));
   if (!ph.ok) return 0;
   const int n = M.nrows;
   const int npat = (int) ph.ptr.size() - 1, nent = (int) ph.off.size();
   if (ph.wide) {
      HB_CUDA(cudaMalloc(&M.pat_code, 2 * (size_t) n + 64));
      HB_CUDA(cudaMemcpy(M.pat_code, ph.code16.data(), 2 * (size_t) n, cudaMemcpyHostToDevice));
   } else {
      HB_CUDA(cudaMalloc(&M.pat_code, (size_t) n + 64));
      HB_CUDA(cudaMemcpy(M.pat_code, ph.code.data(), (size_t) n, cudaMemcpyHostToDevice));
   }
   M.pat_wide = ph.wide;
   HB_CUDA(cudaMalloc(&M.pat_ptr, sizeof(int) * ((size_t) npat + 1)));
   HB_CUDA(cudaMemcpy(M.pat_ptr, ph.ptr.data(), sizeof(int) * ((size_t) npat + 1), cudaMemcpyHostToDevice));
   HB_CUDA(cudaMalloc(&M.pat_off, sizeof(int) * ((size_t) nent + 1)));
   HB_CUDA(cudaMemcpy(M.pat_off, ph.off.data(), sizeof(int) * (size_t) nent, cudaMemcpyHostToDevice));
   HB_CUDA(cudaMalloc(&M.pat_val, sizeof(double) * ((size_t) nent + 1)));
   HB_CUDA(cudaMemcpy(M.pat_val, ph.val.data(), sizeof(double) * (size_t) nent, cudaMemcpyHostToDevice));
   if (!ph.square) {
      HB_CUDA(cudaMalloc(&M.pat_base, sizeof(int) * (size_t) n));
      HB_CUDA(cudaMemcpy(M.pat_base, ph.base.data(), sizeof(int) * (size_t) n, cudaMemcpyHostToDevice));
   }
   if (!ph.irr.empty()) {
      HB_CUDA(cudaMalloc(&M.pat_irr, sizeof(int) * ph.irr.size()));
      HB_CUDA(cudaMemcpy(M.pat_irr, ph.irr.data(), sizeof(int) * ph.irr.size(), cudaMemcpyHostToDevice));
   }
   M.pat_irr_nnz = ph.irr_nnz;
   M.pat_nirr = (int) ph.irr.size();
   M.pat_skips_boundary = ph.skips_boundary;
   M.pat_npat = npat;
   M.pat_nent = nent;
   M.has_pat = true;
   // compact stencils: the register-window kernel (spmv_box) takes the block
   if (!env_flag("HB200_NO_BOX", false)) {
      BoxHost bh;
      box_analyze_host(ph, n, bh);
      if (bh.ok) {
         HB_CUDA(cudaMalloc(&M.box_mask, sizeof(unsigned int) * (size_t) npat));
         HB_CUDA(cudaMemcpy(M.box_mask, bh.mask.data(), sizeof(unsigned int) * (size_t) npat, cudaMemcpyHostToDevice));
         HB_CUDA(cudaMalloc(&M.box_val, sizeof(double) * (size_t) npat * 27));
         HB_CUDA(cudaMemcpy(M.box_val, bh.val.data(), sizeof(double) * (size_t) npat * 27, cudaMemcpyHostToDevice));
         M.box_sy = bh.sy; M.box_sz = bh.sz; M.box_p0 = bh.p0;
         M.box_uniform = bh.uniform; M.box_ref_mask = bh.ref_mask;
         // the kernel-argument values: the reference pattern's when the table is uniform, else the full pattern's
         const int pv = bh.uniform ? bh.pref : bh.p0;
         for (int t = 0; t < 27; t++) M.box_p0_val[t] = pv >= 0 ? bh.val[(size_t) pv * 27 + t] : 0.0;
         // the stencil sweep pays its fixed cost per row (9 shared-memory reads, 27 predicated slots) back only on
         // dense stencils: a 7-point operator stays with the generic row-pattern kernel (7 gathers per row;
         // B200, 256^3: 0.071 ms against 0.19 ms)
         M.has_box = box_fits(M) && (bh.uniform || bh.p0 >= 0) && __builtin_popcount(bh.ref_mask) >= 15;
         M.box_geo = M.has_box && box_geo_check(ph, bh, n);
      }
   }
   return 0;
}

}  // namespace hb

extern "C" int hb200_host_pattern_analyze(int num_rows, int num_cols, const int *row_ptr, const int *col_ind,
                                          const double *values, unsigned char *row_code, int *row_base,
                                          int *num_patterns, int *pattern_ptr, int *pattern_offset,
                                          double *pattern_value, int *num_irregular, int *irregular_rows)
{
   using namespace hb;
   HB_REQUIRE(num_rows >= 0 && num_cols >= 0 && row_ptr && num_patterns && num_irregular, HB200_ERROR_ARG,
              "hb200_host_pattern_analyze: null argument");
   HB_REQUIRE(row_ptr[num_rows] == 0 || (col_ind && values), HB200_ERROR_ARG, "hb200_host_pattern_analyze: null matrix arrays");
   PatHost ph;
   HB_CHECK(pat_analyze_host(num_rows, num_cols, row_ptr, col_ind, values, ph, false));
   *num_patterns = 0;
   *num_irregular = 0;
   if (!ph.ok) return 0;
   const int npat = (int) ph.ptr.size() - 1;
   *num_patterns = npat;
   *num_irregular = (int) ph.irr.size();
   if (row_code) memcpy(row_code, ph.code.data(), (size_t) num_rows);
   if (row_base) {
      for (int r = 0; r < num_rows; r++) row_base[r] = ph.square ? r : ph.base[r];
   }
   if (pattern_ptr) memcpy(pattern_ptr, ph.ptr.data(), sizeof(int) * ((size_t) npat + 1));
   if (pattern_offset) memcpy(pattern_offset, ph.off.data(), sizeof(int) * ph.off.size());
   if (pattern_value) memcpy(pattern_value, ph.val.data(), sizeof(double) * ph.val.size());
   if (irregular_rows && !ph.irr.empty()) memcpy(irregular_rows, ph.irr.data(), sizeof(int) * ph.irr.size());
   return 0;
}

// the experimental wide variant of the analysis (HB200_PAT_WIDE=1): 16-bit codes, 65535 = row outside
// the table; pattern_capacity = entries the caller's pattern_offset / pattern_value arrays hold
extern "C" int hb200_host_pattern_analyze_wide(int num_rows, int num_cols, const int *row_ptr, const int *col_ind,
                                               const double *values, unsigned short *row_code, int *row_base,
                                               int *num_patterns, int *num_entries, int pattern_capacity,
                                               int *pattern_ptr, int *pattern_offset, double *pattern_value,
                                               int *num_irregular, int *irregular_rows)
{
   using namespace hb;
   HB_REQUIRE(num_rows >= 0 && num_cols >= 0 && row_ptr && num_patterns && num_entries && num_irregular,
              HB200_ERROR_ARG, "hb200_host_pattern_analyze_wide: null argument");
   HB_REQUIRE(row_ptr[num_rows] == 0 || (col_ind && values), HB200_ERROR_ARG,
              "hb200_host_pattern_analyze_wide: null matrix arrays");
   PatHost ph;
   HB_CHECK(pat_analyze_host(num_rows, num_cols, row_ptr, col_ind, values, ph, true));
   *num_patterns = 0; *num_entries = 0; *num_irregular = 0;
   if (!ph.ok) return 0;
   const int npat = (int) ph.ptr.size() - 1;
   *num_patterns = npat;
   *num_entries = (int) ph.off.size();
   *num_irregular = (int) ph.irr.size();
   if (row_code) memcpy(row_code, ph.code16.data(), 2 * (size_t) num_rows);
   if (row_base) {
      for (int r = 0; r < num_rows; r++) row_base[r] = ph.square ? r : ph.base[r];
   }
   if (pattern_ptr) memcpy(pattern_ptr, ph.ptr.data(), sizeof(int) * ((size_t) npat + 1));
   if ((int) ph.off.size() <= pattern_capacity) {
      if (pattern_offset) memcpy(pattern_offset, ph.off.data(), sizeof(int) * ph.off.size());
      if (pattern_value) memcpy(pattern_value, ph.val.data(), sizeof(double) * ph.val.size());
   }
   if (irregular_rows && !ph.irr.empty()) memcpy(irregular_rows, ph.irr.data(), sizeof(int) * ph.irr.size());
   return 0;
}
