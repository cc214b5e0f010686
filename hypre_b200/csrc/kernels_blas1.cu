// kernels_blas1.cu — Krylov / relaxation vector kernels (fp64, HBM-bound).
//
// Replaces hypre_SeqVector{Axpy,InnerProd,Scale,Copy,SetConstantValues,PointwiseDivpy}Host
// (src/seq_mv/vector.c:889-1010, 1083-1358) and their cuBLAS/thrust device twins
// (src/seq_mv/vector_device.c:103-311, src/utilities/device_utils.c:627-716).
//
// * element-wise kernels keep the reference's operation order (separate multiply and add,
//   no FMA contraction) so that their results are bit-identical to the CPU reference;
// * dots are two-stage: per-CTA partials (warp shuffles), the last CTA to finish adds the
//   partials in a fixed order -> bitwise reproducible run to run; the result stays in a
//   device scalar slot so that dependent kernels (alpha, beta of PCG) never wait for the host;
// * the PCG update  x += alpha p, r -= alpha s, <r,r>  is one pass over 4 vectors.
#include "hb_internal.cuh"
#include "hb_ew.cuh"
#include <float.h>

namespace hb {

struct FSet { double *y; double v; __device__ void operator()(size_t i) const { y[i] = v; } };
struct FCopy { const double *x; double *y; __device__ void operator()(size_t i) const { y[i] = x[i]; } };
struct FScale { double *y; double a; __device__ void operator()(size_t i) const { y[i] = __dmul_rn(y[i], a); } };
struct FAxpy {
   const double *x; double *y; double a;
   __device__ void operator()(size_t i) const { y[i] = __dadd_rn(y[i], __dmul_rn(a, x[i])); }
};
struct FAxpbyOut {
   const double *x; const double *y; double *z; double a, b;
   __device__ void operator()(size_t i) const { z[i] = __dadd_rn(__dmul_rn(a, x[i]), __dmul_rn(b, y[i])); }
};
struct FDivpy {
   const double *x; const double *b; double *y; const int *m; int mv;
   __device__ void operator()(size_t i) const
   {
      if (m == nullptr || m[i] == mv) y[i] = __dadd_rn(y[i], x[i] / b[i]);
   }
};
struct FScaleDiv {   // zero-guess Jacobi sweep: u = 0 + (w*f)/d on marked points, else u_old (or 0)
   const double *f; const double *d; double *u; const int *m; int mv; const double *uold; double w;
   int core;   // 1: WeightedJacobi_core semantics (skip when d == 0)
   __device__ void operator()(size_t i) const
   {
      const double keep = uold ? uold[i] : 0.0;
      if (m == nullptr || m[i] == mv) {
         const double di = d[i];
         if (core && di == 0.0) { u[i] = keep; }
         else { u[i] = keep + ((w == 1.0) ? f[i] : __dmul_rn(w, f[i])) / di; }
      } else {
         u[i] = keep;
      }
   }
};
struct FDiagScale {   // HYPRE_ParCSRDiagScale: y = x ./ diag(A)
   const double *dg; const double *x; double *y;
   __device__ void operator()(size_t i) const { y[i] = x[i] / dg[i]; }
};

int vec_set(double *y, double v, size_t n, cudaStream_t st)
{
   if (n == 0) return 0;
   if (v == 0.0) { HB_CUDA(cudaMemsetAsync(y, 0, n * sizeof(double), st)); return 0; }
   FSet f{y, v};
   HB_EW(f, n, st);
   return 0;
}
int vec_copy(const double *x, double *y, size_t n, cudaStream_t st)
{
   if (n == 0 || x == y) return 0;
   HB_CUDA(cudaMemcpyAsync(y, x, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
   return 0;
}
int vec_scale(double a, double *y, size_t n, cudaStream_t st)
{
   // hypre_SeqVectorScaleHost: alpha == 1 is a no-op, alpha == 0 fills zeros (vector.c:922-930)
   if (a == 1.0) return 0;
   if (a == 0.0) return vec_set(y, 0.0, n, st);
   FScale f{y, a};
   HB_EW(f, n, st);
   return 0;
}
int vec_axpy(double a, const double *x, double *y, size_t n, cudaStream_t st)
{
   FAxpy f{x, y, a};
   HB_EW(f, n, st);
   return 0;
}
int vec_axpby_out(double a, const double *x, double b, const double *y, double *z, size_t n,
                  cudaStream_t st)
{
   FAxpbyOut f{x, y, z, a, b};
   HB_EW(f, n, st);
   return 0;
}
int vec_divpy(const double *x, const double *b, double *y, const int *marker, int mval, size_t n,
              cudaStream_t st)
{
   FDivpy f{x, b, y, marker, mval};
   HB_EW(f, n, st);
   return 0;
}
int vec_scale_div(double w, const double *fv, const double *d, double *u, const int *marker,
                  int mval, const double *u_old, size_t n, cudaStream_t st)
{
   FScaleDiv f{fv, d, u, marker, mval, u_old, w, 0};
   HB_EW(f, n, st);
   return 0;
}
int vec_diag_scale(const double *diag, const double *x, double *y, size_t n, cudaStream_t st)
{
   FDiagScale f{diag, x, y};
   HB_EW(f, n, st);
   return 0;
}

// ---------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------
constexpr int kRedThreads = 256;

static inline int red_grid(size_t n)
{
   size_t g = (n + (size_t) kRedThreads * 8 - 1) / ((size_t) kRedThreads * 8);
   if (g > (size_t) kRedBlocksMax) g = kRedBlocksMax;
   if (g < 1) g = 1;
   return (int) g;
}

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double *partials,
                                                   unsigned int *counter, double *out0,
                                                   const int *slots)
{
   __shared__ double sm[NV][kRedThreads / 32];
   __shared__ bool   is_last;
   const int tid = threadIdx.x;
#pragma unroll
   for (int k = 0; k < NV; k++) {
      double s = v[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if ((tid & 31) == 0) sm[k][tid >> 5] = s;
   }
   __syncthreads();
   if (tid == 0) {
#pragma unroll
      for (int k = 0; k < NV; k++) {
         double t = 0.0;
#pragma unroll
         for (int w = 0; w < kRedThreads / 32; w++) t += sm[k][w];
         partials[(size_t) k * kRedBlocksMax + blockIdx.x] = t;
      }
      __threadfence();
      const unsigned int ticket = atomicInc(counter, gridDim.x - 1);
      is_last = (ticket == gridDim.x - 1);
   }
   __syncthreads();
   if (is_last) {
      // fixed-order final sum: thread t adds partials t, t+256, ...; then the same tree
      __threadfence();
      double w[NV];
#pragma unroll
      for (int k = 0; k < NV; k++) {
         double t = 0.0;
         for (int b = tid; b < (int) gridDim.x; b += kRedThreads) {
            t += ((volatile double *) partials)[(size_t) k * kRedBlocksMax + b];
         }
         w[k] = t;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < NV; k++) {
         double s = w[k];
#pragma unroll
         for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
         if ((tid & 31) == 0) sm[k][tid >> 5] = s;
      }
      __syncthreads();
      if (tid == 0) {
#pragma unroll
         for (int k = 0; k < NV; k++) {
            double t = 0.0;
#pragma unroll
            for (int ww = 0; ww < kRedThreads / 32; ww++) t += sm[k][ww];
            out0[slots[k]] = t;
         }
      }
   }
}

__global__ void __launch_bounds__(kRedThreads)
dot_kernel(const double *__restrict__ x, const double *__restrict__ y, size_t n, double *partials,
           unsigned int *counter, double *scalars, int slot)
{
   const size_t stride = (size_t) gridDim.x * kRedThreads;
   double s[1] = {0.0};
   double s1 = 0.0, s2 = 0.0, s3 = 0.0;
   size_t i = (size_t) blockIdx.x * kRedThreads + threadIdx.x;
   for (; i + 3 * stride < n; i += 4 * stride) {
      s[0] += x[i] * y[i];
      s1 += x[i + stride] * y[i + stride];
      s2 += x[i + 2 * stride] * y[i + 2 * stride];
      s3 += x[i + 3 * stride] * y[i + 3 * stride];
   }
   for (; i < n; i += stride) s[0] += x[i] * y[i];
   s[0] = (s[0] + s1) + (s2 + s3);
   __shared__ int slots[1];
   if (threadIdx.x == 0) slots[0] = slot;
   __syncthreads();
   block_reduce_store<1>(s, partials, counter, scalars, slots);
}

__global__ void __launch_bounds__(kRedThreads)
dot2_kernel(const double *__restrict__ x, const double *__restrict__ y,
            const double *__restrict__ z, const double *__restrict__ z2, size_t n,
            double *partials, unsigned int *counter, double *scalars, int slot0, int slot1)
{
   const size_t stride = (size_t) gridDim.x * kRedThreads;
   double s[2] = {0.0, 0.0};
   for (size_t i = (size_t) blockIdx.x * kRedThreads + threadIdx.x; i < n; i += stride) {
      s[0] += x[i] * y[i];
      s[1] += z[i] * z2[i];
   }
   __shared__ int slots[2];
   if (threadIdx.x == 0) { slots[0] = slot0; slots[1] = slot1; }
   __syncthreads();
   block_reduce_store<2>(s, partials, counter, scalars, slots);
}

// scalars[slot0 + k] = <Z_k, x> for k < nv <= kMassNV, Z_k = Z + k * zstride: the inner products of one
// vector with a slab of basis vectors in ONE pass over x (hypre_SeqVectorMassInnerProd,
// src/seq_mv/vector_batched.c:251-330; the device twin loops over cuBLAS dots).  Columns >= nv of the
// reduction land in the scratch slot `dump`.
__global__ void __launch_bounds__(kRedThreads)
mass_dot_kernel(const double *__restrict__ x, const double *__restrict__ Z, size_t zstride, int nv, size_t n,
                double *partials, unsigned int *counter, double *scalars, int slot0, int dump)
{
   const size_t stride = (size_t) gridDim.x * kRedThreads;
   double s[kMassNV];
#pragma unroll
   for (int k = 0; k < kMassNV; k++) s[k] = 0.0;
   for (size_t i = (size_t) blockIdx.x * kRedThreads + threadIdx.x; i < n; i += stride) {
      const double xi = x[i];
#pragma unroll
      for (int k = 0; k < kMassNV; k++) {
         if (k < nv) s[k] += Z[(size_t) k * zstride + i] * xi;
      }
   }
   __shared__ int slots[kMassNV];
   if (threadIdx.x < kMassNV) slots[threadIdx.x] = ((int) threadIdx.x < nv) ? slot0 + (int) threadIdx.x : dump;
   __syncthreads();
   block_reduce_store<kMassNV>(s, partials, counter, scalars, slots);
}

int vec_mass_dot_dev(const double *x, const double *Z, size_t zstride, int nv, size_t n, int slot0, int dump,
                     cudaStream_t st)
{
   Ctx &c = ctx();
   HB_REQUIRE(nv >= 1 && nv <= kMassNV, HB200_ERROR_ARG, "vec_mass_dot_dev: 1..4 vectors per pass");
   HB_LAUNCH(mass_dot_kernel, red_grid(n), kRedThreads, 0, st, x, Z, zstride, nv, n, c.d_partials, c.d_counter,
             c.d_scalars, slot0, dump);
   HB_LAUNCH_CHECK();
   return 0;
}

int vec_dot_dev(const double *x, const double *y, size_t n, int slot, cudaStream_t st)
{
   Ctx &c = ctx();
   HB_LAUNCH(dot_kernel, red_grid(n), kRedThreads, 0, st, x, y, n, c.d_partials, c.d_counter,
             c.d_scalars, slot);
   HB_LAUNCH_CHECK();
   return 0;
}

int vec_dot2_dev(const double *x, const double *y, const double *z, const double *z2, size_t n,
                 int slot0, int slot1, cudaStream_t st)
{
   Ctx &c = ctx();
   HB_LAUNCH(dot2_kernel, red_grid(n), kRedThreads, 0, st, x, y, z, z2, n, c.d_partials,
             c.d_counter, c.d_scalars, slot0, slot1);
   HB_LAUNCH_CHECK();
   return 0;
}

int scalars_allreduce(int slot, int count, cudaStream_t st)
{
   Ctx &c = ctx();
   if (c.nranks <= 1) return 0;
#ifdef HB200_WITH_NCCL
   timer_tick(T_ALLREDUCE);
   HB_NCCL(nccl_api().AllReduce(c.d_scalars + slot, c.d_scalars + slot, count, ncclDouble, ncclSum, c.nccl, st));
   timer_tick(T_OTHER);
   return 0;
#else
   return set_error(HB200_ERROR_GENERIC, "libhb200 built without NCCL but nranks > 1");
#endif
}

int scalars_fetch(int slot, int count, double *out, cudaStream_t st)
{
   Ctx &c = ctx();
   timer_tick(T_HOST_SYNC);
   HB_CUDA(cudaMemcpyAsync(c.h_scalars + slot, c.d_scalars + slot, sizeof(double) * count,
                           cudaMemcpyDeviceToHost, st));
   HB_CUDA(cudaStreamSynchronize(st));
   timer_tick(T_OTHER);
   for (int k = 0; k < count; k++) out[k] = c.h_scalars[slot + k];
   return halo_check_error();   // a polling halo kernel gave up: the numbers above mean nothing
}

// ---------------------------------------------------------------------------------------
// PCG fused updates (pcg.c:584-651, 977-978)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRedThreads)
pcg_update_xr_kernel(const double *__restrict__ p, const double *__restrict__ s,
                     double *__restrict__ x, double *__restrict__ r, size_t n, double *partials,
                     unsigned int *counter, double *scalars, int slot_gamma, int slot_sdotp,
                     int slot_rr, int slot_flag, int skip_break)
{
   // the reference's breakdown tests on alpha (pcg.c:588-636), evaluated on the device so
   // that the host never has to wait before x and r are updated
   const double gamma = scalars[slot_gamma];
   const double sdotp = scalars[slot_sdotp];
   int flag = 0;
   double alpha = 0.0;
   if (sdotp == 0.0) { flag = 1; }
   else {
      alpha = gamma / sdotp;
      if (alpha <= 0.0)               { flag = (skip_break < 3) ? 2 : -2; }
      else if (!(alpha >= 4.9406564584124654e-324)) { flag = (skip_break < 2) ? 3 : -3; }
      else if (!(alpha >= DBL_MIN))   { flag = (skip_break < 1) ? 4 : -4; }
   }
   double acc[1] = {0.0};
   if (flag <= 0) {
      const size_t stride = (size_t) gridDim.x * kRedThreads;
      for (size_t i = (size_t) blockIdx.x * kRedThreads + threadIdx.x; i < n; i += stride) {
         x[i] = __dadd_rn(x[i], __dmul_rn(alpha, p[i]));
         const double rn = __dadd_rn(r[i], __dmul_rn(-alpha, s[i]));
         r[i] = rn;
         acc[0] += rn * rn;
      }
   }
   __shared__ int slots[1];
   if (threadIdx.x == 0) slots[0] = slot_rr;
   if (blockIdx.x == 0 && threadIdx.x == 0) {
      scalars[slot_flag]     = (double) flag;
      scalars[slot_flag + 1] = alpha;
   }
   __syncthreads();
   block_reduce_store<1>(acc, partials, counter, scalars, slots);
}

int pcg_update_xr(const double *p, const double *s, double *x, double *r, size_t n,
                  int slot_gamma, int slot_sdotp, int slot_rr, int slot_flag, int skip_break,
                  cudaStream_t st)
{
   Ctx &c = ctx();
   HB_LAUNCH(pcg_update_xr_kernel, red_grid(n), kRedThreads, 0, st, p, s, x, r, n, c.d_partials,
             c.d_counter, c.d_scalars, slot_gamma, slot_sdotp, slot_rr, slot_flag, skip_break);
   HB_LAUNCH_CHECK();
   return 0;
}

struct FUpdateP {
   const double *s; double *p; const double *scalars; int sn, sd;
   __device__ void operator()(size_t i) const
   {
      const double beta = scalars[sn] / scalars[sd];
      // ScaleVector(beta, p); Axpy(1.0, s, p)
      p[i] = __dadd_rn(__dmul_rn(p[i], beta), s[i]);
   }
};

int pcg_update_p(const double *s, double *p, size_t n, int slot_num, int slot_den, cudaStream_t st)
{
   FUpdateP f{s, p, ctx().d_scalars, slot_num, slot_den};
   HB_EW(f, n, st);
   return 0;
}

}  // namespace hb
