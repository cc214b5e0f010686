// relax.cu — relaxation sweeps of the BoomerAMG solve phase.
//
// Reference: hypre_BoomerAMGRelax dispatch (src/parcsr_ls/par_relax.c:23-173),
// hypre_BoomerAMGRelaxWeightedJacobi_core (:180-314), Relax18WeightedL1Jacobi (:338-369),
// Relax7Jacobi (:1178-1254), hypre_BoomerAMGRelaxIF (par_relax_interface.c:19-65),
// hypre_ParCSRRelax_Cheby_SolveHost (par_cheby_solve.c:194-345).
//
// Jacobi-type sweeps are ONE SpMV-shaped pass: the reference's  copy f->Vtemp ; Vtemp = w f -
// w A u ; u += Vtemp ./ l1  (3 vector kernels + 1 SpMV) becomes the SpMV kernel with the
// EPI_JACOBI7 epilogue writing u_new out of place; the zero-initial-guess sweep (flagged by
// hypre_ParVectorAllZeros) is a single element-wise kernel with no matrix traffic.
#include "hb_internal.cuh"
#include "hb_ew.cuh"
#include "relax.cuh"

namespace hb {

bool relax_is_jacobi(int relax_type) { return relax_type == 0 || relax_type == 7 || relax_type == 18; }
bool relax_is_gs(int t)
{
   return t == 3 || t == 4 || t == 6 || t == 8 || t == 13 || t == 14 || t == 88 || t == 89;
}

// split operation (parcsr.cu, parcsr_matvec): main kernel over the rows without offd entries, boundary kernel
// (put + flags + complete boundary rows) beside it on the side stream
static int relax_split(hb200_parcsr *A, const double *x, int epi, const EpiArgs &ea)
{
   Ctx &c = ctx();
   const bool peer = (c.halo_mode == 1 && c.nranks > 1);
   if (peer) HB_CHECK(peer_plans_ensure(A, false));
   if (peer && !A->pkg.peer_off) {
      timer_tick(T_HALO_START);
      HB_CUDA(cudaEventRecord(c.ev_a, c.s_comp));
      HB_CUDA(cudaStreamWaitEvent(c.s_comm, c.ev_a, 0));
      HB_CHECK(parcsr_boundary_launch(A, x, epi, ea, true, c.s_comm));
      HB_CUDA(cudaEventRecord(c.ev_b, c.s_comm));
      timer_tick(T_MATVEC_DIAG);
      HB_CHECK(spmv_launch(A->diag, x, epi, ea, false, c.s_comp));
      HB_CUDA(cudaStreamWaitEvent(c.s_comp, c.ev_b, 0));
   } else {
      timer_tick(T_HALO_START);
      HB_CHECK(parcsr_halo_begin(A, x, c.s_comp));
      timer_tick(T_MATVEC_DIAG);
      HB_CHECK(spmv_launch(A->diag, x, epi, ea, false, c.s_comp));
      timer_tick(T_HALO_WAIT);
      HB_CHECK(parcsr_halo_end(A, c.s_comp));
      timer_tick(T_MATVEC_OFFD);
      HB_CHECK(parcsr_boundary_launch(A, x, epi, ea, false, c.s_comp));
   }
   timer_tick(T_OTHER);
   return 0;
}

// One out-of-place Jacobi-type sweep: u_out = sweep(u_in).  u_in may be NULL when
// zero_guess (u == 0 by flag, memory content undefined).
int relax_jacobi_oop(hb200_parcsr *A, const double *f, const int *cf, int relax_type,
                     int relax_points, double w, const double *l1, const double *u_in,
                     double *u_out, bool zero_guess, bool *used_shortcut)
{
   Ctx &c = ctx();
   const int n = A->num_rows;
   if (used_shortcut) *used_shortcut = false;
   // an armed fused-dot request (<u_out, w>) is honoured by the matvec-form sweep only
   const bool want_dot = c.dot_req_armed;
   c.dot_req_armed = false;
   c.last_dot_fused = false;
   // which reference routine does this (type, points) pair run?
   //   0            -> core(Skip_diag = 1, d = diagonal)
   //   7            -> Relax7Jacobi (matvec form, marked divpy)
   //   18, points=0 -> Relax7Jacobi ; 18, points!=0 -> core(Skip_diag = 0, d = l1)   (:355-368)
   const bool form7 = (relax_type == 7) || (relax_type == 18 && relax_points == 0);
   if (form7) {
      HB_REQUIRE(l1 != nullptr || n == 0, HB200_ERROR_ARG, "relax type 7/18 needs l1_norms");   // (a rank may own no rows of a level)
      if (zero_guess) {
         // Vtemp = w*f (Scale), u = 0 + Vtemp ./ l1   (par_relax.c:1221-1244)
         if (used_shortcut) *used_shortcut = true;
         timer_tick(T_RELAX_ZERO);
         int fl = vec_scale_div(w, f, l1, u_out, relax_points ? cf : nullptr, relax_points, nullptr,
                                (size_t) n, c.s_comp);
         timer_tick(T_OTHER);
         return fl;
      }
      EpiArgs ea;
      ea.w = w; ea.b = f; ea.u = u_in; ea.d = l1; ea.y = u_out;
      ea.cf = relax_points ? cf : nullptr; ea.relax_points = relax_points;
      {
         bool done = false;
         HB_CHECK(parcsr_fused_try(A, u_in, EPI_JACOBI7, ea, &done));   // one kernel on a latency-bound level
         if (done) return 0;
      }
      if (parcsr_main_skips_boundary(A)) return relax_split(A, u_in, EPI_JACOBI7, ea);
      timer_tick(T_HALO_START);
      HB_CHECK(parcsr_halo_begin(A, u_in, c.s_comp));
      timer_tick(T_MATVEC_DIAG);
      if (want_dot && c.nranks == 1 && A->num_cols_offd == 0 && spmv_can_fuse_dot(A->diag, EPI_JACOBI7)) {
         ea.dotw = c.dot_req_w; ea.dot_slot = c.dot_req_slot;
         c.last_dot_fused = true;
      }
      HB_CHECK(spmv_launch(A->diag, u_in, EPI_JACOBI7, ea, false, c.s_comp));
      HB_CHECK(parcsr_offd_pass(A, EPI_JACOBI7_ACC, ea));
      timer_tick(T_OTHER);
      return 0;
   }
   // core form
   const double *uin = u_in;
   if (zero_guess) {
      // the reference runs the full sweep on an all-zero u; materialise the zeros
      HB_REQUIRE(u_in != nullptr || n == 0, HB200_ERROR_ARG, "core Jacobi needs a u_in buffer");
      HB_CHECK(vec_set((double *) u_in, 0.0, (size_t) n, c.s_comp));
   }
   HB_REQUIRE(relax_points == 0 || cf != nullptr || n == 0, HB200_ERROR_ARG, "CF relaxation needs cf_marker");
   EpiArgs ea;
   ea.w = w; ea.b = f; ea.u = uin; ea.y = u_out;
   ea.cf = cf; ea.relax_points = relax_points;
   if (relax_type == 0) {
      // the divisor is the diagonal of the diag block for BOTH passes (the offd pass cannot
      // find it in its own block): use the extracted diagonal vector
      const double *dg = nullptr;
      HB_CHECK(parcsr_diag(A, &dg));
      ea.d = dg; ea.skip_diag = 1;
   } else { ea.d = l1; ea.skip_diag = 0; }
   {
      bool done = false;
      HB_CHECK(parcsr_fused_try(A, uin, EPI_JACOBI_CORE, ea, &done));
      if (done) return 0;
   }
   if (parcsr_main_skips_boundary(A)) return relax_split(A, uin, EPI_JACOBI_CORE, ea);
   HB_CHECK(parcsr_halo_begin(A, uin, c.s_comp));
   HB_CHECK(spmv_launch(A->diag, uin, EPI_JACOBI_CORE, ea, false, c.s_comp));
   HB_CHECK(parcsr_offd_pass(A, EPI_JACOBI_CORE_ACC, ea));
   timer_tick(T_OTHER);
   return 0;
}

// ---------------------------------------------------------------------------------------
// Chebyshev (par_cheby_solve.c:194-345).  v, r, tmp: level-sized scratch.
// ---------------------------------------------------------------------------------------
struct FChebyStart {   // r = ds*(f + tmp) [scaled] ; o = u ; u = r*c
   const double *f; const double *tmp; const double *ds; double *r; double *o; double *u; double c;
   int scale;
   __device__ void operator()(size_t i) const
   {
      double rr;
      if (scale) { rr = __dmul_rn(ds[i], __dadd_rn(f[i], tmp[i])); r[i] = rr; }
      else       { rr = r[i]; }
      o[i] = u[i];
      u[i] = __dmul_rn(rr, c);
   }
};
struct FMul { const double *a; const double *b; double *y;
   __device__ void operator()(size_t i) const { y[i] = __dmul_rn(a[i], b[i]); } };
struct FChebyStep {    // u = mult*r + ds*v  (or + v unscaled)
   const double *r; const double *v; const double *ds; double *u; double mult; int scale;
   __device__ void operator()(size_t i) const
   {
      const double t = scale ? __dmul_rn(ds[i], v[i]) : v[i];
      u[i] = __dadd_rn(__dmul_rn(mult, r[i]), t);
   }
};
struct FChebyEnd {     // u = o + ds*u  (or o + u)
   const double *o; const double *ds; double *u; int scale;
   __device__ void operator()(size_t i) const
   {
      const double t = scale ? __dmul_rn(ds[i], u[i]) : u[i];
      u[i] = __dadd_rn(o[i], t);
   }
};

int cheby_solve(hb200_parcsr *A, const double *f, const double *ds, const double *coefs,
                int order, int scale, double *u, double *v, double *r, double *orig_u, double *tmp)
{
   Ctx &c = ctx();
   const size_t n = (size_t) A->num_rows;
   if (order > 4) order = 4;
   if (order < 1) order = 1;
   const int cheby_order = order - 1;
   HB_REQUIRE(!scale || ds != nullptr || n == 0, HB200_ERROR_ARG, "scaled Chebyshev needs ds");
   // ---- fused form: every element-wise step rides in the epilogue of the SpMV that produces its input
   // (`order` launches per sweep instead of 3*order + 2).  Blocks with an offd part keep the unfused form:
   // the epilogue needs the complete row sum, which the offd pass adds later.
   const bool fuse = env_flag("HB200_FUSED_CHEBY", true);
   const int kd = A->diag.kind;
   if (fuse && n > 0 && A->num_cols_offd == 0 && c.nranks == 1 &&
       (kd == SPMV_VECTOR || kd == SPMV_VECTOR16 || kd == SPMV_PAT || kd == SPMV_BOX)) {
      const double *dsv = scale ? ds : nullptr;
      // SpMV inputs ping-pong between two scratch vectors; orig_u is u itself, untouched until the last step
      double *in_a = scale ? tmp : v, *in_b = orig_u;    // scaled: t = ds*u' ; unscaled: u' itself
      EpiArgs e1;
      e1.b = f; e1.d = dsv; e1.w = coefs[cheby_order]; e1.r_out = r; e1.u = u;
      e1.cheby_last = (cheby_order == 0);
      // (order 1: the only SpMV multiplies u itself, so the result cannot land in u directly)
      e1.y = v; e1.y2 = in_a;
      if (!scale && !e1.cheby_last) e1.y = in_a;
      HB_CHECK(spmv_launch(A->diag, u, EPI_CHEBY_FIRST, e1, false, c.s_comp));
      if (e1.cheby_last) return vec_copy(v, u, n, c.s_comp);
      const double *xin = in_a;
      for (int i = cheby_order - 1; i >= 0; i--) {
         EpiArgs es;
         es.d = dsv; es.w = coefs[i]; es.r = r; es.u = u;
         es.cheby_last = (i == 0);
         double *next_in = (xin == in_a) ? in_b : in_a;
         if (es.cheby_last) { es.y = u; }
         else if (scale)    { es.y = v; es.y2 = next_in; }
         else               { es.y = next_in; }
         HB_CHECK(spmv_launch(A->diag, xin, EPI_CHEBY_STEP, es, false, c.s_comp));
         xin = next_in;
      }
      return 0;
   }
   if (!scale) {
      HB_CHECK(parcsr_matvec(A, -1.0, u, 1.0, f, r));
      FChebyStart fs{f, nullptr, nullptr, r, orig_u, u, coefs[cheby_order], 0};
      HB_EW(fs, n, c.s_comp);
      for (int i = cheby_order - 1; i >= 0; i--) {
         HB_CHECK(parcsr_matvec(A, 1.0, u, 0.0, v, v));
         FChebyStep st{r, v, nullptr, u, coefs[i], 0};
         HB_EW(st, n, c.s_comp);
      }
      FChebyEnd fe{orig_u, nullptr, u, 0};
      HB_EW(fe, n, c.s_comp);
   } else {
      HB_CHECK(parcsr_matvec(A, -1.0, u, 0.0, tmp, tmp));
      FChebyStart fs{f, tmp, ds, r, orig_u, u, coefs[cheby_order], 1};
      HB_EW(fs, n, c.s_comp);
      for (int i = cheby_order - 1; i >= 0; i--) {
         FMul fm{ds, u, tmp};
         HB_EW(fm, n, c.s_comp);
         HB_CHECK(parcsr_matvec(A, 1.0, tmp, 0.0, v, v));
         FChebyStep st{r, v, ds, u, coefs[i], 1};
         HB_EW(st, n, c.s_comp);
      }
      FChebyEnd fe{orig_u, ds, u, 1};
      HB_EW(fe, n, c.s_comp);
   }
   return 0;
}

// ---------------------------------------------------------------------------------------
// coarsest-level Gaussian elimination (hypre_gselim, src/utilities/gselim.h:11-66;
// hypre_GaussElimSolve types 9/19, src/parcsr_ls/par_gauss_elim.c:457-695)
// ---------------------------------------------------------------------------------------
int ge_factor_host(const double *A_mat, int n, std::vector<double> &LfT, std::vector<double> &UT,
                   std::vector<double> &Udiag)
{
   // run the reference's elimination once on a copy, recording the multipliers it would use
   // (same skip-on-zero tests); the per-solve work is then the two triangular sweeps only
   std::vector<double> A(A_mat, A_mat + (size_t) n * n);
   LfT.assign((size_t) n * n, 0.0);
   for (int k = 0; k < n - 1; k++) {
      if (A[(size_t) k * n + k] != 0.0) {
         const double divA = 1.0 / A[(size_t) k * n + k];
         for (int j = k + 1; j < n; j++) {
            if (A[(size_t) j * n + k] != 0.0) {
               const double factor = A[(size_t) j * n + k] * divA;
               for (int m = k + 1; m < n; m++) {
                  A[(size_t) j * n + m] -= factor * A[(size_t) k * n + m];
               }
               LfT[(size_t) k * n + j] = factor;   // column k of L, stored contiguously
            }
         }
      }
   }
   UT.assign((size_t) n * n, 0.0);
   Udiag.assign(n, 0.0);
   for (int k = 0; k < n; k++) {
      Udiag[k] = A[(size_t) k * n + k];
      for (int j = 0; j < k; j++) UT[(size_t) k * n + j] = A[(size_t) j * n + k];   // column k of U
   }
   return 0;
}

__global__ void ge_solve_kernel(int n, const double *__restrict__ LfT, const double *__restrict__ UT,
                                const double *__restrict__ Udiag, const double *__restrict__ b,
                                double *__restrict__ u_local, int first_row, int num_local)
{
   HB_DYN_SHARED(double, x);
   const int tid = threadIdx.x;
   for (int j = tid; j < n; j += blockDim.x) x[j] = b[j];
   __syncthreads();
   if (n == 1) {
      if (tid == 0 && Udiag[0] != 0.0) x[0] = x[0] / Udiag[0];
   } else {
      // forward elimination applied to the right-hand side: x[j] -= factor(j,k) * x[k]
      for (int k = 0; k < n - 1; k++) {
         const double xk = x[k];
         for (int j = k + 1 + tid; j < n; j += blockDim.x) {
            const double fac = LfT[(size_t) k * n + j];
            if (fac != 0.0) x[j] = __dadd_rn(x[j], -__dmul_rn(fac, xk));
         }
         __syncthreads();
      }
      // back substitution
      for (int k = n - 1; k > 0; --k) {
         const double dk = Udiag[k];
         if (dk != 0.0) {
            if (tid == 0) x[k] = x[k] / dk;
            __syncthreads();
            const double xk = x[k];
            for (int j = tid; j < k; j += blockDim.x) {
               const double ujk = UT[(size_t) k * n + j];
               if (ujk != 0.0) x[j] = __dadd_rn(x[j], -__dmul_rn(xk, ujk));
            }
         }
         __syncthreads();
      }
      if (tid == 0 && Udiag[0] != 0.0) x[0] = x[0] / Udiag[0];
   }
   __syncthreads();
   for (int j = tid; j < num_local; j += blockDim.x) u_local[j] = x[first_row + j];
}

__global__ void ge_scatter_kernel(int n, int first_row, int num_local, const double *__restrict__ f,
                                  double *__restrict__ b)
{
   const int j = blockIdx.x * blockDim.x + threadIdx.x;
   if (j < n) {
      const int l = j - first_row;
      b[j] = (l >= 0 && l < num_local) ? f[l] : 0.0;
   }
}

int ge_solve(const GEData &ge, const double *f_local, double *u_local)
{
   Ctx &c = ctx();
   if (ge.n <= 0) return 0;
   timer_tick(T_GE_SOLVE);
   if (c.nranks > 1) {
      // hypre_MPI_Allgatherv of the coarse right-hand side (par_gauss_elim.c:577): every rank
      // deposits its slice into a zeroed n-vector and the vector is summed (x + 0 is exact)
      HB_LAUNCH(ge_scatter_kernel, (ge.n + 255) / 256, 256, 0, c.s_comp, ge.n, ge.first_row,
                ge.num_local, f_local, ge.d_b);
      HB_LAUNCH_CHECK();
#ifdef HB200_WITH_NCCL
      HB_NCCL(nccl_api().AllReduce(ge.d_b, ge.d_b, ge.n, ncclDouble, ncclSum, c.nccl, c.s_comp));
#endif
      if (ge.num_local == 0) return 0;   // par_gauss_elim.c:600-617: ranks without rows leave
   } else {
      HB_CUDA(cudaMemcpyAsync(ge.d_b, f_local, sizeof(double) * (size_t) ge.n, cudaMemcpyDeviceToDevice, c.s_comp));
   }
   int nt = 32;
   while (nt < ge.n && nt < 1024) nt <<= 1;
   HB_LAUNCH(ge_solve_kernel, 1, nt, sizeof(double) * (size_t) ge.n, c.s_comp, ge.n, ge.d_LfT,
             ge.d_UT, ge.d_Udiag, ge.d_b, u_local, ge.first_row, ge.num_local);
   HB_LAUNCH_CHECK();
   timer_tick(T_OTHER);
   return 0;
}

}  // namespace hb

using namespace hb;

extern "C" {

int hb200_relax(hb200_parcsr *A, const double *f, const int *cf_marker, int relax_type,
                int relax_points, double relax_weight, double omega, const double *l1_norms,
                double *u, int u_all_zeros, double *vtemp)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A && ((f && u) || A->num_rows == 0), HB200_ERROR_ARG, "null argument");
   Ctx &c = ctx();
   if (relax_is_jacobi(relax_type)) {
      HB_REQUIRE(vtemp != nullptr || A->num_rows == 0, HB200_ERROR_ARG, "Jacobi relaxation needs vtemp scratch");
      bool shortcut = false;
      // out of place into vtemp, then back into u (the AMG cycle avoids this copy by swapping)
      const bool core_zero = u_all_zeros && !((relax_type == 7) || (relax_type == 18 && relax_points == 0));
      if (core_zero) HB_CHECK(vec_set(u, 0.0, (size_t) A->num_rows, c.s_comp));
      HB_CHECK(relax_jacobi_oop(A, f, cf_marker, relax_type, relax_points, relax_weight, l1_norms,
                                u, vtemp, u_all_zeros && !core_zero, &shortcut));
      return vec_copy(vtemp, u, (size_t) A->num_rows, c.s_comp);
   }
   if (relax_is_gs(relax_type)) {
      if (u_all_zeros) HB_CHECK(vec_set(u, 0.0, (size_t) A->num_rows, c.s_comp));
      return relax_hybrid_gs(A, f, cf_marker, relax_type, relax_points, relax_weight, omega,
                             l1_norms, u, vtemp);
   }
   return set_error(HB200_ERROR_ARG, "hb200_relax: relax_type %d is not on the B200 path", relax_type);
}

int hb200_relax_if(hb200_parcsr *A, const double *f, const int *cf_marker, int relax_type,
                   int relax_order, int cycle_param, double relax_weight, double omega,
                   const double *l1_norms, double *u, int u_all_zeros, double *vtemp)
{
   // par_relax_interface.c:19-65
   if (relax_order == 1 && cycle_param < 3) {
      const int pts[2] = {cycle_param < 2 ? 1 : -1, cycle_param < 2 ? -1 : 1};
      for (int i = 0; i < 2; i++) {
         HB_CHECK(hb200_relax(A, f, cf_marker, relax_type, pts[i], relax_weight, omega, l1_norms, u,
                              u_all_zeros, vtemp));
         u_all_zeros = 0;
      }
      return 0;
   }
   return hb200_relax(A, f, cf_marker, relax_type, 0, relax_weight, omega, l1_norms, u, u_all_zeros, vtemp);
}

int hb200_cheby_solve(hb200_parcsr *A, const double *f, const double *ds, const double *coefs,
                      int order, int scale, int variant, double *u)
{
   (void) variant;   // unused by the reference's solve too (par_cheby_solve.c:207)
   HB_CHECK(require_ready());
   HB_REQUIRE(A && coefs && ((f && u) || A->num_rows == 0), HB200_ERROR_ARG, "null argument");
   const size_t n = (size_t) (A->num_rows ? A->num_rows : 1);
   double *w = nullptr;
   HB_CUDA(cudaMalloc(&w, sizeof(double) * n * 4));
   int fl = cheby_solve(A, f, ds, coefs, order, scale, u, w, w + n, w + 2 * n, w + 3 * n);
   cudaStreamSynchronize(ctx().s_comp);
   cudaFree(w);
   return fl;
}

}  // extern "C"
