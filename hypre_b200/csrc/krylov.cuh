// krylov.cuh — what the Krylov drivers share (krylov.cu: PCG, GMRES; krylov_ext.cu: BiCGSTAB,
// FlexGMRES, COGMRES): the preconditioner call, the device scalar slots, dots that stay on the device.
#pragma once
#include "hb_internal.cuh"
#include "hb_ew.cuh"
#include "relax.cuh"
#include <float.h>
#include <math.h>
#include <string.h>

namespace hb {
int amg_solve(hb200_amg *amg, hb200_parcsr *A, const double *f, double *u, bool u_all_zeros,
              int *num_iterations, double *rel_resid_norm, double *resid_norms = nullptr, double *rhs_norm_out = nullptr);
// fused-dot request for the next amg_solve (preconditioner use): slot >= 0 asks the cycle's last
// level-0 sweep for <u, f>; amg_dot_fused tells whether it delivered (else the caller runs dot_kernel)
void amg_set_dot_request(hb200_amg *amg, int slot);
bool amg_dot_fused(const hb200_amg *amg);
}

namespace hb {

// device scalar slots used by the Krylov drivers
// PCG: the three dots that close an iteration (<r,s> = gamma, <r,r>, flexible <r_old,s>) sit in one
// block of adjacent slots so that ONE all-reduce serves them; two blocks alternate between
// iterations because the previous gamma is still needed (beta = gamma / gamma_old)
enum {
   S_BB = 0, S_SDOTP = 1, S_FLAG = 2, S_ALPHA = 3,   // S_ALPHA = S_FLAG + 1 (pcg_update_xr_kernel)
   S_BLK0 = 4, S_BLK1 = 8,                           // {gamma, rr, delta} of even / odd iterations
   B_GAMMA = 0, B_RR = 1, B_DELTA = 2,
   S_T0 = 12, S_T1 = 13,
   S_NFETCH = 12,                                    // slots the host reads once per iteration
   S_H0 = 16   // GMRES: hh column (k_dim + 1 entries, k_dim <= 100)
};

// dot_slot >= 0: the caller wants <r, z> in that scalar slot next; *dot_done says whether the
// preconditioner's last kernel already produced it (fused epilogue, single rank, row-pattern A_0)
[[maybe_unused]] static int precond_apply(int kind, hb200_amg *amg, hb200_parcsr *A, const double *r, double *z,
                         int dot_slot = -1, bool *dot_done = nullptr)
{
   // the Krylov solvers always ClearVector(z) first => zero initial guess
   Ctx &c = ctx();
   const size_t n = (size_t) A->num_rows;
   if (dot_done) *dot_done = false;
   switch (kind) {
      case HB200_PRECOND_AMG: {
         amg_set_dot_request(amg, (dot_done && fused_dots_enabled() && c.nranks == 1) ? dot_slot : -1);
         const int fl = amg_solve(amg, A, r, z, true, nullptr, nullptr) & ~HB200_ERROR_CONV;
         if (dot_done) *dot_done = (fl == 0) && amg_dot_fused(amg);
         amg_set_dot_request(amg, -1);
         return fl;
      }
      case HB200_PRECOND_DIAGSCALE: {
         const double *dg = nullptr;
         HB_CHECK(parcsr_diag(A, &dg));
         return vec_diag_scale(dg, r, z, n, c.s_comp);
      }
      default:   // hypre_ParKrylovIdentity: copy
         return vec_copy(r, z, n, c.s_comp);
   }
}

struct FAxpyDev {   // y += sign*S[slot] * x
   const double *x; double *y; const double *S; int slot; double sign;
   __device__ void operator()(size_t i) const
   {
      const double a = sign * S[slot];
      y[i] = __dadd_rn(y[i], __dmul_rn(a, x[i]));
   }
};
struct FScaleInvSqrtDev {   // y *= 1/sqrt(S[slot]) unless S[slot] == 0
   double *y; const double *S; int slot;
   __device__ void operator()(size_t i) const
   {
      const double t = sqrt(S[slot]);
      if (t != 0.0) y[i] = __dmul_rn(y[i], 1.0 / t);
   }
};

[[maybe_unused]] static int dot_global(const double *x, const double *y, size_t n, int slot)
{
   Ctx &c = ctx();
   timer_tick(T_BLAS1);
   HB_CHECK(vec_dot_dev(x, y, n, slot, c.s_comp));
   timer_tick(T_OTHER);
   return scalars_allreduce(slot, 1, c.s_comp);
}

[[maybe_unused]] static int dot_global_host(const double *x, const double *y, size_t n, double *out)
{
   HB_CHECK(dot_global(x, y, n, S_T0));
   return scalars_fetch(S_T0, 1, out, ctx().s_comp);
}

}  // namespace hb
