// hb_ew.cuh — element-wise kernel template shared by the vector / relaxation translation units.
#pragma once
#include "hb_internal.cuh"

namespace hb {

constexpr int kEwThreads = 256;
constexpr int kEwUnroll  = 4;

static inline int ew_grid(size_t n)
{
   size_t g = (n + (size_t) kEwThreads * kEwUnroll - 1) / ((size_t) kEwThreads * kEwUnroll);
   const size_t cap = (size_t) kNumSMs * 16;
   if (g > cap) g = cap;
   if (g < 1) g = 1;
   return (int) g;
}

template <class F>
__global__ void __launch_bounds__(kEwThreads) ew_kernel(size_t n, F f)
{
   const size_t stride = (size_t) gridDim.x * kEwThreads;
   size_t i = (size_t) blockIdx.x * kEwThreads + threadIdx.x;
   // 4 independent iterations in flight per thread
   for (; i + 3 * stride < n; i += 4 * stride) {
      f(i);
      f(i + stride);
      f(i + 2 * stride);
      f(i + 3 * stride);
   }
   for (; i < n; i += stride) f(i);
}

#define HB_EW(functor, n, st)                                                            \
   do {                                                                                  \
      if ((n) > 0) {                                                                     \
         HB_LAUNCH((ew_kernel), ew_grid(n), kEwThreads, 0, st, (size_t) (n), functor);   \
         HB_LAUNCH_CHECK();                                                              \
      }                                                                                  \
   } while (0)


}  // namespace hb
