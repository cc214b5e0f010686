// oracle/emu/cuda_runtime.h — TEST INFRASTRUCTURE, not product code.
//
// A host-only stand-in for the slice of the CUDA runtime and device language that
// hypre_b200/csrc/*.cu uses, so that the very same kernel sources compile with g++ into
// oracle/_ref/libhb200_emu.so and their LOGIC (indexing, format decoding, epilogues, reductions,
// control flow of the cycle and the Krylov drivers) can be checked on a machine without a GPU.
// It says nothing about performance and is never loaded by the hypre_b200 package: only tests
// (tests/test_emu_kernels.py) build and load it.
//
// Execution model: one kernel launch runs its blocks one after the other; the threads of a block
// are fibers (ucontext) on the calling OS thread, switched only at the points where CUDA threads
// can observe each other: __syncthreads() and the warp shuffles.  Everything is deterministic.
// Stream capture records launches and async copies with their arguments by value; a graph launch
// replays the list.  Multi-rank runs: one host process per rank, NCCL calls over oracle/minimpi
// (nccl_emu.cpp), exported "device" memory in POSIX shared memory.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <chrono>
#include <functional>
#include <tuple>

#ifndef HB200_EMU
#error "oracle/emu/cuda_runtime.h is the emulation header: compile with -DHB200_EMU"
#endif

// ---------------------------------------------------------------------------------------
// language
// ---------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static

struct hb_emu_dim3 { unsigned x = 1, y = 1, z = 1; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) int4 { int x, y, z, w; };

namespace hb_emu {
struct ThreadCtx { hb_emu_dim3 tid, bid, bdim, gdim; };
extern ThreadCtx *g_cur;             // the running fiber's coordinates
void  syncthreads();
void  syncwarp();
unsigned long long shfl_down_bits(unsigned long long bits, unsigned delta, int width);
int   vote_all(int pred);
void *dyn_smem();
void  launch(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body);
long long launches();
// stream capture: while a capture is open, launches and async copies are recorded (arguments by
// value, as CUDA copies kernel parameters at launch) instead of run; a graph launch replays them
bool  capturing();
void  record(std::function<void()> op);
// kernel launch whose body already holds its arguments by value
inline void launch_bound(unsigned grid, unsigned block, size_t smem, std::function<void()> body)
{
   if (capturing()) record([=]() { launch(grid, block, smem, body); });
   else launch(grid, block, smem, body);
}
}  // namespace hb_emu

#define threadIdx (hb_emu::g_cur->tid)
#define blockIdx  (hb_emu::g_cur->bid)
#define blockDim  (hb_emu::g_cur->bdim)
#define gridDim   (hb_emu::g_cur->gdim)

inline void __syncthreads() { hb_emu::syncthreads(); }
inline void __syncwarp(unsigned = 0xffffffffu) { hb_emu::syncwarp(); }
inline void __threadfence() {}
inline void __threadfence_system() {}

template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32)
{
   static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
   unsigned long long bits = 0;
   memcpy(&bits, &v, sizeof(T));
   bits = hb_emu::shfl_down_bits(bits, delta, width);
   T out;
   memcpy(&out, &bits, sizeof(T));
   return out;
}

inline int __all_sync(unsigned, int pred) { return hb_emu::vote_all(pred); }

template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcs(const T *p) { return *p; }
template <class T> inline T __ldcg(const T *p) { return *p; }
template <class T> inline void __stcs(T *p, T v) { *p = v; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline unsigned int atomicInc(unsigned int *p, unsigned int val)
{
   const unsigned int old = *p;
   *p = (old >= val) ? 0u : old + 1u;
   return old;
}
using std::max;
using std::min;

// ---------------------------------------------------------------------------------------
// runtime API (host memory stands in for device memory; streams and events are ordinals)
// ---------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNotSupported = 801, cudaErrorInvalidValue = 1 };
typedef struct hb_emu_stream *cudaStream_t;
typedef struct hb_emu_event { double t_ms; } *cudaEvent_t;
typedef struct hb_emu_graph *cudaGraph_t;
typedef struct hb_emu_graph_exec *cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal, cudaStreamCaptureModeThreadLocal, cudaStreamCaptureModeRelaxed };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; size_t totalGlobalMem; };

inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 16; return cudaSuccess; }   // one emulated device per host process
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
   memset(p, 0, sizeof(*p));
   snprintf(p->name, sizeof(p->name), "emulated sm_100 (host fibers)");
   p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
   return cudaSuccess;
}
namespace hb_emu {
void *dev_alloc(size_t bytes);       // large blocks are page-mapped so that they can be exported (IPC)
void  dev_free(void *p);
int   ipc_export(void *handle64, void *p);
int   ipc_open(void **p, const void *handle64);
int   ipc_close(void *p);
}
template <class T> inline cudaError_t cudaMalloc(T **p, size_t bytes)
{
   *p = (T *) hb_emu::dev_alloc(bytes ? bytes : 8);
   return *p ? cudaSuccess : 2;
}
template <class T> inline cudaError_t cudaMallocHost(T **p, size_t bytes) { return cudaMalloc(p, bytes); }
enum { cudaHostAllocMapped = 2, cudaHostAllocPortable = 1 };
template <class T> inline cudaError_t cudaHostAlloc(T **p, size_t bytes, unsigned) { return cudaMalloc(p, bytes); }
template <class T> inline cudaError_t cudaHostGetDevicePointer(T **d, void *h, unsigned) { *d = (T *) h; return cudaSuccess; }
inline cudaError_t cudaFree(void *p) { hb_emu::dev_free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void *p) { hb_emu::dev_free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr)
{
   if (hb_emu::capturing()) { hb_emu::record([=]() { if (n) memmove(d, s, n); }); return cudaSuccess; }
   if (n) memmove(d, s, n);
   return cudaSuccess;
}
inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr)
{
   if (hb_emu::capturing()) { hb_emu::record([=]() { if (n) memset(d, v, n); }); return cudaSuccess; }
   if (n) memset(d, v, n);
   return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t) malloc(8); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline double hb_emu_now_ms()
{
   return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t) malloc(sizeof(hb_emu_event)); (*e)->t_ms = 0; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t_ms = hb_emu_now_ms(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float) (b->t_ms - a->t_ms); return cudaSuccess; }
// graphs: the recorded operations of one capture, replayed in order
namespace hb_emu {
int  capture_begin();
int  capture_end(void **graph);
int  graph_instantiate(void **exec, void *graph);
int  graph_launch(void *exec);
void graph_destroy(void *graph_or_exec);
}
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return hb_emu::capture_begin() ? cudaErrorNotSupported : cudaSuccess; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *g) { return hb_emu::capture_end((void **) g) ? cudaErrorNotSupported : cudaSuccess; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t *e, cudaGraph_t g, unsigned long long = 0) { return hb_emu::graph_instantiate((void **) e, g) ? cudaErrorNotSupported : cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t) { return hb_emu::graph_launch(e) ? cudaErrorNotSupported : cudaSuccess; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { hb_emu::graph_destroy(g); return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { hb_emu::graph_destroy(e); return cudaSuccess; }
// "device" memory exported to the other host processes of a multi-rank run: POSIX shared memory
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { return hb_emu::ipc_export(h->reserved, p) ? cudaErrorNotSupported : cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { return hb_emu::ipc_open(p, h.reserved) ? cudaErrorNotSupported : cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void *p) { hb_emu::ipc_close(p); return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int threads, size_t)
{
   *n = std::max(1, 1536 / std::max(threads, 1));
   return cudaSuccess;
}
template <class T> inline cudaError_t cudaMemcpyToSymbol(T &symbol, const void *src, size_t n) { memcpy(&symbol, src, n); return cudaSuccess; }
