// hb_epilogue.cuh — the fused SpMV epilogues shared by every SpMV kernel.
#pragma once
#include "hb_internal.cuh"

namespace hb {

// ---------------------------------------------------------------------------------------
// epilogues
// ---------------------------------------------------------------------------------------
// The epilogue vectors (b/f, l1 norms, cf marker) are read once per sweep and y is written once:
// streaming (evict-first) accesses keep them from displacing the gathered x lines in L1/L2.
// the arithmetic of the two epilogues that also exist with their per-row inputs staged by the caller
// (spmv_box reads b / f / l1 through its shared-memory ring): one definition, identical rounding
__device__ __forceinline__ double epi_axpby_value(const EpiArgs &ea, double b, double sum)
{
   // reference: y = (beta/alpha)*b; y += sum; y *= alpha  ==  beta*b + alpha*sum
   // (csr_matvec.c:836-845); for alpha = +-1 the specialised branches are exact copies.
   return ea.beta * b + ea.alpha * sum;
}
__device__ __forceinline__ double epi_jacobi7_value(const EpiArgs &ea, double uo, double f, double d, double sum)
{
   // Vtemp = w*f - w*A*u ; u += Vtemp ./ l1   (par_relax.c:1216-1244)
   const double vt = (ea.w == 1.0) ? (f - sum) : (ea.w * f - ea.w * sum);
   return uo + vt / d;
}

template <int EPI>
__device__ __forceinline__ void epi_apply(const EpiArgs &ea, int row, double sum, double diag)
{
   if (EPI == EPI_AXPBY) {
      double v;
      if (ea.beta == 0.0) { v = ea.alpha * sum; }
      else                { v = epi_axpby_value(ea, __ldcs(ea.b + row), sum); }
      __stcs(ea.y + row, v);
   }
   else if (EPI == EPI_ACC) {
      ea.y[row] += ea.alpha * sum;
   }
   else if (EPI == EPI_JACOBI7) {
      // Vtemp = w*f - w*A*u ; u += Vtemp ./ l1   (par_relax.c:1216-1244)
      const double uo = ea.u[row];
      if (ea.cf == nullptr || __ldcs(ea.cf + row) == ea.relax_points) {
         __stcs(ea.y + row, epi_jacobi7_value(ea, uo, __ldcs(ea.b + row), __ldcs(ea.d + row), sum));
      } else {
         __stcs(ea.y + row, uo);
      }
   }
   else if (EPI == EPI_JACOBI7_ACC) {
      if (ea.cf == nullptr || ea.cf[row] == ea.relax_points) {
         ea.y[row] -= (ea.w * sum) / ea.d[row];
      }
   }
   else if (EPI == EPI_JACOBI_CORE) {
      // hypre_BoomerAMGRelaxWeightedJacobi_core (par_relax.c:258-295)
      const double uo = ea.u[row];
      const double di = ea.d ? ea.d[row] : diag;
      if ((ea.relax_points == 0 || ea.cf[row] == ea.relax_points) && di != 0.0) {
         const double res = ea.b[row] - sum;
         if (ea.skip_diag) { ea.y[row] = uo * (1.0 - ea.w) + ea.w * res / di; }
         else              { ea.y[row] = uo + ea.w * res / di; }
      } else {
         ea.y[row] = uo;
      }
   }
   else if (EPI == EPI_CHEBY_FIRST || EPI == EPI_CHEBY_STEP) {
      // hypre_ParCSRRelax_Cheby_SolveHost (par_cheby_solve.c:284-341), one row: d = ds (NULL: unscaled
      // variant), w = the coefficient of this step, b = f, u = orig_u; same operations, same order
      const double ds = ea.d ? __ldcs(ea.d + row) : 1.0;
      double un;
      if (EPI == EPI_CHEBY_FIRST) {
         double rr = __dadd_rn(__ldcs(ea.b + row), -sum);          // f + (-A u)
         if (ea.d) rr = __dmul_rn(ds, rr);
         __stcs(ea.r_out + row, rr);
         un = __dmul_rn(rr, ea.w);
      } else {
         const double t = ea.d ? __dmul_rn(ds, sum) : sum;
         un = __dadd_rn(__dmul_rn(ea.w, __ldcs(ea.r + row)), t);
      }
      if (ea.cheby_last) {
         const double t = ea.d ? __dmul_rn(ds, un) : un;
         ea.y[row] = __dadd_rn(ea.u[row], t);                      // u = orig_u + ds*u
      } else {
         ea.y[row] = un;
         if (ea.d) ea.y2[row] = __dmul_rn(ds, un);                 // the next SpMV multiplies ds*u
      }
   }
   else if (EPI == EPI_JACOBI_CORE_ACC) {
      const double di = ea.d ? ea.d[row] : diag;
      if ((ea.relax_points == 0 || ea.cf[row] == ea.relax_points) && di != 0.0) {
         ea.y[row] -= ea.w * sum / di;
      }
   }
}

// The same arithmetic as epi_apply<EPI_AXPBY / EPI_JACOBI7>, returning the value it stored: used by the
// kernels that fuse a dot product with the vector they write (kernels_pat.cu, DOT variants).
template <int EPI>
__device__ __forceinline__ double epi_apply_ret(const EpiArgs &ea, int row, double sum)
{
   double v;
   if (EPI == EPI_AXPBY) {
      if (ea.beta == 0.0) { v = ea.alpha * sum; }
      else                { v = epi_axpby_value(ea, __ldcs(ea.b + row), sum); }
   } else {   // EPI_JACOBI7
      const double uo = ea.u[row];
      if (ea.cf == nullptr || __ldcs(ea.cf + row) == ea.relax_points) {
         v = epi_jacobi7_value(ea, uo, __ldcs(ea.b + row), __ldcs(ea.d + row), sum);
      } else {
         v = uo;
      }
   }
   __stcs(ea.y + row, v);
   return v;
}

template <int EPI>
__device__ __forceinline__ constexpr bool epi_needs_diag()
{
   return EPI == EPI_JACOBI_CORE || EPI == EPI_JACOBI_CORE_ACC;
}


}  // namespace hb
