// hb_peer.cuh — device-side primitives of the NVLink peer-put halo (parcsr_peer.cu, kernels_offd.cu):
// system-scope acquire / release accesses and the BOUNDED spin.
//
// A kernel that polls a flag written by another GPU must not be able to hold the device for ever: if a
// peer died, left the collective sequence or never launched its put, the poll gives up after
// `timeout_ns` (HB200_HALO_TIMEOUT_S, default 30 s), records what it was waiting for in a host-mapped
// error word and lets the kernel finish; every later poll sees the word and leaves at once, and the
// host turns it into an error at its next synchronisation point (halo_check_error).
#pragma once
#include "hb_internal.cuh"
#ifdef HB200_EMU
#include <sched.h>
#include <chrono>
#endif

namespace hb {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
   unsigned long long v;
#ifndef HB200_EMU
   asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
#else
   v = *(const volatile unsigned long long *) p;
#endif
   return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
#ifndef HB200_EMU
   asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
   *(volatile unsigned long long *) p = v;
#endif
}
__device__ __forceinline__ unsigned long long peer_now_ns()
{
#ifndef HB200_EMU
   unsigned long long t;
   asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
   return t;
#else
   return (unsigned long long) std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
#endif
}

// what a polling kernel needs to give up cleanly
struct SpinGuard {
   unsigned long long *err = nullptr;        // host-mapped error word (0 = fine)
   unsigned long long  timeout_ns = 0;       // 0 = wait for ever
};

// error word: [63:56] what (1 = put waits for an ack, 2 = wait polls an arrival flag), [55:40] segment,
// [39:0] the epoch waited for
__device__ __forceinline__ unsigned long long spin_err_word(int what, int seg, unsigned long long epoch)
{
   return ((unsigned long long) what << 56) | ((unsigned long long) (seg & 0xffff) << 40) | (epoch & 0xffffffffffull);
}

// poll *p until it reaches `target`; false = gave up (the data of this exchange is then undefined and
// the solve is reported as failed by the host)
__device__ __forceinline__ bool spin_until_ge(const unsigned long long *p, unsigned long long target,
                                              const SpinGuard &g, int what, int seg)
{
   if (ld_acquire_sys(p) >= target) return true;
   unsigned long long t0 = 0;
   unsigned int spins = 0;
   while (ld_acquire_sys(p) < target) {
#ifndef HB200_EMU
      __nanosleep(40);                       // a system-scope poll is not free for the memory system: back off
#endif
      if ((++spins & 63u) != 0u) continue;
#ifdef HB200_EMU
      sched_yield();                         // the peer is another host process
#endif
      if (g.err == nullptr) continue;
      if (*(volatile unsigned long long *) g.err != 0ull) return false;        // somebody gave up already
      if (g.timeout_ns == 0ull) continue;
      const unsigned long long now = peer_now_ns();
      if (t0 == 0ull) { t0 = now; continue; }
      if (now - t0 > g.timeout_ns) {
         *(volatile unsigned long long *) g.err = spin_err_word(what, seg, target);
#ifndef HB200_EMU
         __threadfence_system();
#endif
         return false;
      }
   }
   return true;
}

// device-scope relay of an arrival: ONE block polls the system-scope flags the peers write over NVLink and
// republishes the epoch in local memory; every other block of the kernel waits on that local word (a
// thousand blocks polling system-scope flags slowed the whole device down by 5x: profiles/r2_session_log.md)
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p)
{
   unsigned long long v;
#ifndef HB200_EMU
   asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
#else
   v = *(const volatile unsigned long long *) p;
#endif
   return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v)
{
#ifndef HB200_EMU
   asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
   *(volatile unsigned long long *) p = v;
#endif
}
__device__ __forceinline__ bool spin_local_until_ge(const unsigned long long *p, unsigned long long target, const SpinGuard &g)
{
   unsigned int spins = 0;
   unsigned long long t0 = 0;
   while (ld_acquire_gpu(p) < target) {
#ifndef HB200_EMU
      __nanosleep(100);
#else
      sched_yield();
#endif
      if ((++spins & 255u) != 0u || g.err == nullptr) continue;
      if (*(volatile unsigned long long *) g.err != 0ull) return false;
      if (g.timeout_ns == 0ull) continue;
      const unsigned long long now = peer_now_ns();
      if (t0 == 0ull) { t0 = now; continue; }
      if (now - t0 > 2 * g.timeout_ns) return false;   // (the relay block reports the reason)
   }
   return true;
}

}  // namespace hb
