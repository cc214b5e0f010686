// gs.cu — hybrid Gauss-Seidel / SOR family on the diag block (relax types 3, 4, 6, 8, 13, 14,
// 88, 89): Jacobi across ranks (frozen halo), exact sequential Gauss-Seidel inside the rank.
//
// Reference: hypre_BoomerAMGRelaxHybridGaussSeidel_core (src/parcsr_ls/par_relax.c:692-945)
// and its inner loops hypre_HybridGaussSeidelNS / hypre_HybridGaussSeidel
// (src/parcsr_ls/par_relax.h:12-110, 238-330), single-thread semantics (the reference's
// OpenMP variant changes the result with the thread count).
//
// Two device formulations, selected per matrix by hb200_parcsr_set_gs_chunks (the reference's
// hypre_NumThreads(), par_relax.c:727):
//
//  * chunks <= 1: the 1-thread reference, exactly (below: wavefronts);
//  * chunks = T > 1: the reference's OpenMP semantics with T threads (par_relax.c:868-896 over
//    hypre_HybridGaussSeidelNSThreads / hypre_HybridGaussSeidelThreads, par_relax.h:107-218, 332-440): rows are
//    cut into T contiguous chunks by hypre_partition1D, Gauss-Seidel inside a chunk, the values of all other
//    chunks frozen at their state before the call (Vtemp).  One group of K lanes walks one chunk row after
//    row, 32 / K chunks per warp in lock step, ONE launch per call whatever the matrix: with a few thousand
//    chunks the sweep is bound by the HBM stream of the CSR block like the Jacobi sweep, not by launches.
//    The result is the reference's at OMP_NUM_THREADS = T (tests run both at the same T); the l1 norms of
//    the l1 variants belong to the same T (hypre_ParCSRComputeL1NormsThreads, ams.c:4523).
//
// Wavefront formulation: the sequential sweep is a sparse triangular dependency graph.  At first
// use the rows are grouped into wavefronts (level schedule) of the *symmetrised* pattern, so
// that every row in a wavefront (a) has all its lower-index neighbours finished and (b) is
// not read by any row of the same wavefront; the sweep is then one launch per wavefront over
// a level-sorted row list, updating u in place, K lanes per row.  Results are identical to the
// 1-thread reference up to the summation order inside a row.
#include "hb_internal.cuh"
#include "hb_ew.cuh"
#include "relax.cuh"
#include <algorithm>

namespace hb {

struct GsSched {
   int  nlev_f = 0, nlev_b = 0;
   std::vector<int> lptr_f, lptr_b;          // host level pointers
   int *d_perm_f = nullptr, *d_perm_b = nullptr;
};

int gs_sched_free(void *p)
{
   GsSched *s = (GsSched *) p;
   if (!s) return 0;
   if (s->d_perm_f) cudaFree(s->d_perm_f);
   if (s->d_perm_b) cudaFree(s->d_perm_b);
   delete s;
   return 0;
}

static void build_levels(int n, const int *ai, const int *aj, bool forward, std::vector<int> &perm,
                         std::vector<int> &lptr)
{
   std::vector<int> lev(n, 0);
   if (forward) {
      for (int i = 0; i < n; i++) {
         int l = lev[i];   // holds the anti-dependency bound accumulated so far
         for (int p = ai[i]; p < ai[i + 1]; p++) {
            const int j = aj[p];
            if (j < i && lev[j] + 1 > l) l = lev[j] + 1;
         }
         lev[i] = l;
         for (int p = ai[i]; p < ai[i + 1]; p++) {
            const int j = aj[p];
            if (j > i && j < n && lev[j] < l + 1) lev[j] = l + 1;
         }
      }
   } else {
      for (int i = n - 1; i >= 0; i--) {
         int l = lev[i];
         for (int p = ai[i]; p < ai[i + 1]; p++) {
            const int j = aj[p];
            if (j > i && j < n && lev[j] + 1 > l) l = lev[j] + 1;
         }
         lev[i] = l;
         for (int p = ai[i]; p < ai[i + 1]; p++) {
            const int j = aj[p];
            if (j < i && lev[j] < l + 1) lev[j] = l + 1;
         }
      }
   }
   int nlev = 0;
   for (int i = 0; i < n; i++) nlev = std::max(nlev, lev[i] + 1);
   lptr.assign((size_t) nlev + 1, 0);
   for (int i = 0; i < n; i++) lptr[lev[i] + 1]++;
   for (int l = 0; l < nlev; l++) lptr[l + 1] += lptr[l];
   perm.resize(n);
   std::vector<int> next(lptr.begin(), lptr.end() - 1);
   if (forward) { for (int i = 0; i < n; i++) perm[next[lev[i]]++] = i; }
   else         { for (int i = n - 1; i >= 0; i--) perm[next[lev[i]]++] = i; }
}

static int gs_get_sched(hb200_parcsr *A, GsSched **out)
{
   if (A->gs_sched) { *out = (GsSched *) A->gs_sched; return 0; }
   const int n = A->num_rows;
   std::vector<int> hi((size_t) n + 1), hj((size_t) A->diag.nnz);
   HB_CUDA(cudaMemcpy(hi.data(), A->diag.i, sizeof(int) * hi.size(), cudaMemcpyDeviceToHost));
   if (A->diag.nnz) HB_CUDA(cudaMemcpy(hj.data(), A->diag.j, sizeof(int) * hj.size(), cudaMemcpyDeviceToHost));
   GsSched *s = new GsSched();
   std::vector<int> perm;
   build_levels(n, hi.data(), hj.data(), true, perm, s->lptr_f);
   s->nlev_f = (int) s->lptr_f.size() - 1;
   if (n) {
      HB_CUDA(cudaMalloc(&s->d_perm_f, sizeof(int) * (size_t) n));
      HB_CUDA(cudaMemcpy(s->d_perm_f, perm.data(), sizeof(int) * (size_t) n, cudaMemcpyHostToDevice));
   }
   build_levels(n, hi.data(), hj.data(), false, perm, s->lptr_b);
   s->nlev_b = (int) s->lptr_b.size() - 1;
   if (n) {
      HB_CUDA(cudaMalloc(&s->d_perm_b, sizeof(int) * (size_t) n));
      HB_CUDA(cudaMemcpy(s->d_perm_b, perm.data(), sizeof(int) * (size_t) n, cudaMemcpyHostToDevice));
   }
   A->gs_sched = s;
   *out = s;
   return 0;
}

struct GsArgs {
   const int *di, *dj; const double *da;
   const int *oi, *oj; const double *oa;
   const double *f, *l1, *vtmp, *vext;
   const int *cf;
   double *u;
   int relax_points, skip_diag, non_scale;
   double w, omega, one_minus_omega, prod;
};

template <int K>
__global__ void __launch_bounds__(256)
gs_level_kernel(int nrows, const int *__restrict__ rows, GsArgs g)
{
   const int gtid = blockIdx.x * 256 + threadIdx.x;
   const int idx  = gtid / K;
   const int lane = threadIdx.x % K;
   const bool active = idx < nrows;
   int i = 0;
   bool relax = false;
   double d = 0.0;
   double s0 = 0.0, s2 = 0.0, so = 0.0;   // sum a*u (current), sum a*vtmp (old), sum offd
   if (active) {
      i = rows[idx];
      d = g.l1 ? g.l1[i] : g.da[g.di[i]];
      relax = (g.relax_points == 0 || g.cf[i] == g.relax_points) && d != 0.0;
      if (relax) {
         const int p1 = g.di[i + 1];
         // u is updated in place by other wavefronts only: plain (coherent) loads
         for (int p = g.di[i] + g.skip_diag + lane; p < p1; p += K) {
            const int j = g.dj[p];
            const double a = g.da[p];
            s0 += a * g.u[j];
            if (!g.non_scale) s2 += a * g.vtmp[j];
         }
         if (g.oi) {
            const int q1 = g.oi[i + 1];
            for (int q = g.oi[i] + lane; q < q1; q += K) so += g.oa[q] * g.vext[g.oj[q]];
         }
      }
   }
#pragma unroll
   for (int o = K / 2; o > 0; o >>= 1) {
      s0 += __shfl_down_sync(0xffffffffu, s0, o, K);
      s2 += __shfl_down_sync(0xffffffffu, s2, o, K);
      so += __shfl_down_sync(0xffffffffu, so, o, K);
   }
   if (active && relax && lane == 0) {
      if (g.non_scale) {
         // par_relax.h:44-70: res = f - sum_diag - sum_offd ; u = res/d  or  u += res/d
         const double res = g.f[i] - s0 - so;
         if (g.skip_diag) g.u[i] = res / d;
         else             g.u[i] = g.u[i] + res / d;
      } else {
         // par_relax.h:262-282
         const double res  = g.f[i] - so;
         const double res0 = -s0;
         const double res2 = s2;
         double ui = g.u[i];
         if (g.skip_diag) ui *= g.prod;
         ui += g.w * (g.omega * res + res0 + g.one_minus_omega * res2) / d;
         g.u[i] = ui;
      }
   }
}

static int gs_sweep(hb200_parcsr *A, const GsSched *s, bool forward, const GsArgs &g, int K)
{
   Ctx &c = ctx();
   const std::vector<int> &lp = forward ? s->lptr_f : s->lptr_b;
   const int *perm = forward ? s->d_perm_f : s->d_perm_b;
   const int nlev = forward ? s->nlev_f : s->nlev_b;
   for (int l = 0; l < nlev; l++) {
      const int cnt = lp[l + 1] - lp[l];
      if (cnt == 0) continue;
      const long long threads = (long long) cnt * K;
      const int grid = (int) ((threads + 255) / 256);
      switch (K) {
         case 1:  HB_LAUNCH((gs_level_kernel<1>),  grid, 256, 0, c.s_comp, cnt, perm + lp[l], g); break;
         case 2:  HB_LAUNCH((gs_level_kernel<2>),  grid, 256, 0, c.s_comp, cnt, perm + lp[l], g); break;
         case 4:  HB_LAUNCH((gs_level_kernel<4>),  grid, 256, 0, c.s_comp, cnt, perm + lp[l], g); break;
         case 8:  HB_LAUNCH((gs_level_kernel<8>),  grid, 256, 0, c.s_comp, cnt, perm + lp[l], g); break;
         case 16: HB_LAUNCH((gs_level_kernel<16>), grid, 256, 0, c.s_comp, cnt, perm + lp[l], g); break;
         default: HB_LAUNCH((gs_level_kernel<32>), grid, 256, 0, c.s_comp, cnt, perm + lp[l], g); break;
      }
   }
   HB_LAUNCH_CHECK();
   (void) A;
   return 0;
}

// ---------------------------------------------------------------------------------------
// T chunks (the reference with T OpenMP threads)
// ---------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256)
gs_chunk_kernel(int n, int T, int nsweeps, int gs_order, GsArgs g)
{
   constexpr int G = 32 / K;   // chunks per warp
   const int warp   = (int) ((blockIdx.x * 256u + threadIdx.x) >> 5);
   const int lane32 = threadIdx.x & 31;
   const int lane   = lane32 % K;
   const int chunk  = warp * G + lane32 / K;
   // hypre_partition1D (utilities/general.c): the first `rest` chunks hold one row more
   const int size = n / T, rest = n - size * T;
   int ns = 0, ne = 0;
   if (chunk < T) {
      if (chunk < rest) { ns = chunk * (size + 1); ne = ns + size + 1; }
      else              { ns = chunk * size + rest; ne = ns + size; }
   }
   const int cnt = ne - ns;
   const int maxrows = size + (rest ? 1 : 0);   // the same for every lane: the loops below stay warp-uniform
   for (int sw = 0; sw < nsweeps; sw++) {
      const int iorder = nsweeps == 1 ? gs_order : (sw == 0 ? 1 : -1);
      for (int step = 0; step < maxrows; step++) {
         const bool have = step < cnt;
         const int i = iorder > 0 ? ns + step : ne - 1 - step;
         bool relax = false;
         double d = 0.0;
         double res = 0.0, res0 = 0.0, res2 = 0.0;   // frozen part (other chunks + offd), own chunk (current), own chunk (old)
         if (have) {
            d = g.l1 ? g.l1[i] : g.da[g.di[i]];
            relax = (g.relax_points == 0 || g.cf[i] == g.relax_points) && d != 0.0;
         }
         if (relax) {
            const int p1 = g.di[i + 1];
            for (int p = g.di[i] + g.skip_diag + lane; p < p1; p += K) {
               const int ii = g.dj[p];
               const double a = g.da[p];
               if (ii >= ns && ii < ne) {
                  res0 -= a * g.u[ii];                       // rows of this chunk: written by lane 0 of this group only
                  if (!g.non_scale) res2 += a * g.vtmp[ii];
               } else {
                  res -= a * g.vtmp[ii];
               }
            }
            if (g.oi) {
               const int q1 = g.oi[i + 1];
               for (int q = g.oi[i] + lane; q < q1; q += K) res -= g.oa[q] * g.vext[g.oj[q]];
            }
         }
#pragma unroll
         for (int o = K / 2; o > 0; o >>= 1) {
            res  += __shfl_down_sync(0xffffffffu, res, o, K);
            res0 += __shfl_down_sync(0xffffffffu, res0, o, K);
            res2 += __shfl_down_sync(0xffffffffu, res2, o, K);
         }
         if (relax && lane == 0) {
            if (g.non_scale) {
               // par_relax.h:150-176: res = f - sum ; u = res/d  or  u += res/d
               const double r = g.f[i] + res + res0;
               if (g.skip_diag) g.u[i] = r / d;
               else             g.u[i] = g.u[i] + r / d;
            } else {
               // par_relax.h:376-400
               const double r = g.f[i] + res;
               double ui = g.u[i];
               if (g.skip_diag) ui *= g.prod;
               ui += g.w * (g.omega * r + res0 + g.one_minus_omega * res2) / d;
               g.u[i] = ui;
            }
         }
         __syncwarp();   // the next row of the chunk reads what lane 0 just wrote
      }
   }
}

static int gs_chunks_launch(hb200_parcsr *A, int nsweeps, int gs_order, const GsArgs &g)
{
   Ctx &c = ctx();
   const int n = A->num_rows, T = A->gs_chunks;
   if (n == 0) return 0;
   const double avg = A->diag.avg_row_nnz;
   const int K = avg >= 40 ? 32 : avg >= 20 ? 16 : avg >= 6 ? 8 : 4;
   const int G = 32 / K;
   const long long warps = ((long long) T + G - 1) / G;
   const int grid = (int) ((warps + 7) / 8);
   switch (K) {
      case 4:  HB_LAUNCH((gs_chunk_kernel<4>),  grid, 256, 0, c.s_comp, n, T, nsweeps, gs_order, g); break;
      case 8:  HB_LAUNCH((gs_chunk_kernel<8>),  grid, 256, 0, c.s_comp, n, T, nsweeps, gs_order, g); break;
      case 16: HB_LAUNCH((gs_chunk_kernel<16>), grid, 256, 0, c.s_comp, n, T, nsweeps, gs_order, g); break;
      default: HB_LAUNCH((gs_chunk_kernel<32>), grid, 256, 0, c.s_comp, n, T, nsweeps, gs_order, g); break;
   }
   HB_LAUNCH_CHECK();
   return 0;
}

static int gs_core(hb200_parcsr *A, const double *f, const int *cf, int relax_points, double w,
                   double omega, const double *l1, double *u, double *vtemp, int gs_order,
                   int symm, int skip_diag)
{
   // hypre_BoomerAMGRelaxHybridGaussSeidel_core, num_threads == 1 branch (par_relax.c:905-936)
   Ctx &c = ctx();
   const bool chunked = A->gs_chunks > 1 && A->num_rows > 0;
   GsSched *s = nullptr;
   if (!chunked) HB_CHECK(gs_get_sched(A, &s));
   const int non_scale = (w == 1.0 && omega == 1.0);
   // halo exchange of u, once per call (par_relax.c:806-835), frozen during the sweep(s)
   HB_CHECK(parcsr_halo_begin(A, u, c.s_comp));
   if (!non_scale || chunked) {
      HB_REQUIRE(vtemp != nullptr || A->num_rows == 0, HB200_ERROR_ARG, "scaled / chunked hybrid GS needs vtemp");
      HB_CHECK(vec_copy(u, vtemp, (size_t) A->num_rows, c.s_comp));   // par_relax.c:857-866
   }
   HB_CHECK(parcsr_halo_end(A, c.s_comp));
   HB_REQUIRE(relax_points == 0 || cf != nullptr || A->num_rows == 0, HB200_ERROR_ARG, "CF relaxation needs cf_marker");
   GsArgs g;
   g.di = A->diag.i; g.dj = A->diag.j; g.da = A->diag.a;
   if (A->num_cols_offd > 0) { g.oi = A->offd.i; g.oj = A->offd.j; g.oa = A->offd.a; g.vext = A->pkg.d_recv_buf; }
   else { g.oi = nullptr; g.oj = nullptr; g.oa = nullptr; g.vext = nullptr; }
   g.f = f; g.l1 = l1; g.vtmp = vtemp; g.cf = cf; g.u = u;
   g.relax_points = relax_points; g.skip_diag = skip_diag; g.non_scale = non_scale;
   g.w = w; g.omega = omega; g.one_minus_omega = 1.0 - omega; g.prod = 1.0 - w * omega;
   const double avg = A->diag.avg_row_nnz;
   const int K = avg >= 48 ? 16 : avg >= 20 ? 8 : avg >= 8 ? 4 : avg >= 3 ? 2 : 1;
   const int nsweeps = symm ? 2 : 1;
   // T chunks: both sweeps of a symmetric call stay inside the chunk, against the same frozen copy
   // (par_relax.c:876-895)
   if (chunked) return gs_chunks_launch(A, nsweeps, gs_order > 0 ? 1 : -1, g);
   for (int sw = 0; sw < nsweeps; sw++) {
      const int iorder = nsweeps == 1 ? (gs_order > 0 ? 1 : -1) : (sw == 0 ? 1 : -1);
      HB_CHECK(gs_sweep(A, s, iorder > 0, g, K));
   }
   return 0;
}

int relax_hybrid_gs(hb200_parcsr *A, const double *f, const int *cf, int relax_type,
                    int relax_points, double w, double omega, const double *l1, double *u,
                    double *vtemp)
{
   const int skip_l1 = (w == 1.0 && omega == 1.0) ? 0 : 1;   // par_relax.c:1277, 1348, 1372
   switch (relax_type) {
      case 3:  return gs_core(A, f, cf, relax_points, w, omega, nullptr, u, vtemp, 1, 0, 1);
      case 4:  return gs_core(A, f, cf, relax_points, w, omega, nullptr, u, vtemp, -1, 0, 1);
      case 6:  return gs_core(A, f, cf, relax_points, w, omega, nullptr, u, vtemp, 1, 1, 1);
      case 8:
      case 88: HB_REQUIRE(l1 || A->num_rows == 0, HB200_ERROR_ARG, "l1 GS needs l1_norms");
               return gs_core(A, f, cf, relax_points, w, omega, l1, u, vtemp, 1, 1, skip_l1);
      case 13: HB_REQUIRE(l1 || A->num_rows == 0, HB200_ERROR_ARG, "l1 GS needs l1_norms");
               return gs_core(A, f, cf, relax_points, w, omega, l1, u, vtemp, 1, 0, skip_l1);
      case 14: HB_REQUIRE(l1 || A->num_rows == 0, HB200_ERROR_ARG, "l1 GS needs l1_norms");
               return gs_core(A, f, cf, relax_points, w, omega, l1, u, vtemp, -1, 0, skip_l1);
      case 89: HB_REQUIRE(l1 || A->num_rows == 0, HB200_ERROR_ARG, "l1 GS needs l1_norms");
               HB_CHECK(gs_core(A, f, cf, relax_points, w, omega, l1, u, vtemp, 1, 0, skip_l1));
               return gs_core(A, f, cf, relax_points, w, omega, l1, u, vtemp, -1, 0, skip_l1);
      default: return set_error(HB200_ERROR_ARG, "relax_hybrid_gs: unsupported type %d", relax_type);
   }
}

}  // namespace hb

// hypre_NumThreads() of the hybrid Gauss-Seidel sweeps on this matrix (see the header of this file)
extern "C" int hb200_parcsr_set_gs_chunks(hb200_parcsr *A, int num_chunks)
{
   using namespace hb;
   HB_REQUIRE(A != nullptr && num_chunks >= 0, HB200_ERROR_ARG, "hb200_parcsr_set_gs_chunks: bad argument");
   A->gs_chunks = (num_chunks > A->num_rows) ? A->num_rows : num_chunks;
   return 0;
}

// the chunk count the device prefers for a block of num_rows rows: enough chunks to keep every SM's
// warps busy (148 SMs x 32 warps x 4 chunks), at least 32 rows per chunk
extern "C" int hb200_gs_auto_chunks(int num_rows)
{
   const long long cap = (long long) hb::kNumSMs * 32 * 4;
   long long t = num_rows / 32;
   if (t > cap) t = cap;
   return t < 1 ? 1 : (int) t;
}

// host half of the hybrid-GS path, reachable without a GPU (CPU tests of the schedule)
extern "C" int hb200_host_gs_schedule(int num_rows, const int *row_ptr, const int *col_ind, int forward,
                                      int *perm, int *level_ptr, int *num_levels)
{
   using namespace hb;
   HB_REQUIRE(num_rows >= 0 && row_ptr && num_levels, HB200_ERROR_ARG, "hb200_host_gs_schedule: null argument");
   HB_REQUIRE(row_ptr[num_rows] == 0 || col_ind, HB200_ERROR_ARG, "hb200_host_gs_schedule: null column indices");
   std::vector<int> p, lp;
   build_levels(num_rows, row_ptr, col_ind, forward != 0, p, lp);
   *num_levels = (int) lp.size() - 1;
   if (perm && num_rows > 0) memcpy(perm, p.data(), sizeof(int) * (size_t) num_rows);
   if (level_ptr) memcpy(level_ptr, lp.data(), sizeof(int) * lp.size());
   return 0;
}
