// runtime.cu — process context (device, streams, NCCL communicator), memory and error
// plumbing of libhb200.  Replaces, for the solve path only, hypre_Handle / hypre_TAlloc /
// hypre_TMemcpy (src/utilities/handle.h:36-83, memory.c:956-990) and the hot-path MPI
// wrappers (src/utilities/mpistubs.c:940+).
#include "hb_internal.cuh"
#include <stdarg.h>
#include <string.h>
#include <strings.h>
#include <mutex>
#include <dlfcn.h>
#ifndef HB200_EMU
#include <nvtx3/nvToolsExt.h>
#endif

namespace hb {

static std::string g_err;
static Ctx         g_ctx;

Ctx &ctx() { return g_ctx; }

int set_error(int flag, const char *fmt, ...)
{
   char buf[1024];
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(buf, sizeof(buf), fmt, ap);
   va_end(ap);
   g_err = buf;
   if (getenv("HB200_VERBOSE")) fprintf(stderr, "[hb200] error %d: %s\n", flag, buf);
   return flag;
}

void prof_push(const char *name)
{
#ifndef HB200_EMU
   nvtxRangePushA(name);
#else
   (void) name;
#endif
}
void prof_pop()
{
#ifndef HB200_EMU
   nvtxRangePop();
#endif
}

bool env_flag(const char *name, bool dflt)
{
   const char *e = getenv(name);
   if (!e) return dflt;
   return !(e[0] == 0 || !strcmp(e, "0") || !strcasecmp(e, "off") || !strcasecmp(e, "no"));
}

int require_ready()
{
   if (!g_ctx.ready) {
      return set_error(HB200_ERROR_GENERIC,
                       "hb200_init() has not succeeded: no CUDA device bound (there is no CPU fallback)");
   }
   return 0;
}

#ifdef HB200_WITH_NCCL
static NcclApi g_nccl;
NcclApi &nccl_api() { return g_nccl; }

int nccl_load()
{
   if (g_nccl.loaded) return 0;
#ifndef HB200_EMU
   // RTLD_NOLOAD first: reuse whatever NCCL the process already has (e.g. PyTorch's)
   void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
   if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
   if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
#else
   void *h = dlopen("libnccl_emu.so", RTLD_NOW | RTLD_GLOBAL);   // g++ test build: NCCL calls over host message passing
#endif
   if (!h) return set_error(HB200_ERROR_GENERIC, "cannot load libnccl.so.2: %s", dlerror());
#define HB_SYM(field, name)                                                              \
   *(void **) (&g_nccl.field) = dlsym(h, name);                                          \
   if (!g_nccl.field) return set_error(HB200_ERROR_GENERIC, "libnccl lacks %s", name);
   HB_SYM(GetUniqueId, "ncclGetUniqueId")
   HB_SYM(CommInitRank, "ncclCommInitRank")
   HB_SYM(CommDestroy, "ncclCommDestroy")
   HB_SYM(AllReduce, "ncclAllReduce")
   HB_SYM(AllGather, "ncclAllGather")
   HB_SYM(Send, "ncclSend")
   HB_SYM(Recv, "ncclRecv")
   HB_SYM(GroupStart, "ncclGroupStart")
   HB_SYM(GroupEnd, "ncclGroupEnd")
   HB_SYM(GetErrorString, "ncclGetErrorString")
#undef HB_SYM
   g_nccl.loaded = true;
   return 0;
}
#endif

// ---- stream-time attribution --------------------------------------------------------------------
bool g_timers_on = false;
bool g_trace_on = false;

void trace_line(const char *fmt, ...)
{
   char buf[512];
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(buf, sizeof(buf), fmt, ap);
   va_end(ap);
   fprintf(stderr, "[hb200 trace rank %d] %s\n", ctx().rank, buf);
   fflush(stderr);
}

static std::vector<cudaEvent_t> g_tev;
static std::vector<int> g_tcat;
static size_t g_tn = 0;

void timers_begin()
{
   static bool checked = false;
   if (!checked) { checked = true; g_timers_on = env_flag("HB200_TIMERS", false); }
   if (!g_timers_on) return;
   g_tn = 0;
   timer_tick_impl(T_OTHER);
}

void timer_tick_impl(int cat)
{
   if (g_tn >= g_tev.size()) {
      const size_t old = g_tev.size();
      g_tev.resize(old + 4096);
      g_tcat.resize(old + 4096);
      for (size_t k = old; k < g_tev.size(); k++) cudaEventCreate(&g_tev[k]);
   }
   cudaEventRecord(g_tev[g_tn], g_ctx.s_comp);
   g_tcat[g_tn] = cat;
   g_tn++;
}

void timers_report(const char *what)
{
   if (!g_timers_on || g_tn < 2) return;
   timer_tick_impl(T_OTHER);
   cudaStreamSynchronize(g_ctx.s_comp);
   double acc[T_NUM] = {0};
   long cnt[T_NUM] = {0};
   for (size_t k = 0; k + 1 < g_tn; k++) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, g_tev[k], g_tev[k + 1]);
      acc[g_tcat[k]] += ms;
      cnt[g_tcat[k]]++;
   }
   static const char *names[T_NUM] = {"other", "matvec_diag(+relax)", "matvec_offd", "halo_start(pack/put)", "halo_wait(exposed)",
                                      "blas1", "allreduce", "host_sync(fetch)", "ge_solve", "relax_zero_guess"};
   double tot = 0.0;
   for (int c = 0; c < T_NUM; c++) tot += acc[c];
   fprintf(stderr, "[hb200 timers rank %d] %s: %.3f ms on the compute stream\n", g_ctx.rank, what, tot);
   for (int c = 0; c < T_NUM; c++) {
      if (cnt[c]) fprintf(stderr, "   %-24s %9.3f ms %5.1f%%  (%ld intervals)\n", names[c], acc[c], 100.0 * acc[c] / tot, cnt[c]);
   }
   g_tn = 0;
}

int ws_get(int slot, size_t bytes, double **out)
{
   Ctx &c = g_ctx;
   if (slot < 0 || slot >= 16) return set_error(HB200_ERROR_ARG, "ws_get: bad slot");
   if (bytes == 0) bytes = 8;
   if (c.ws_bytes[slot] < bytes) {
      if (c.ws_ptr[slot]) { cudaStreamSynchronize(c.s_comp); cudaFree(c.ws_ptr[slot]); c.ws_ptr[slot] = nullptr; c.ws_bytes[slot] = 0; }
      cudaError_t e = cudaMalloc(&c.ws_ptr[slot], bytes);
      if (e != cudaSuccess) return set_error(HB200_ERROR_MEMORY, "workspace cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
      c.ws_bytes[slot] = bytes;
   }
   *out = (double *) c.ws_ptr[slot];
   return 0;
}

}  // namespace hb

using namespace hb;

extern "C" {

const char *hb200_last_error(void) { return hb::g_err.c_str(); }
const char *hb200_version(void) { return "hb200 0.1 (sm_100a)"; }

int hb200_init(int device)
{
   Ctx &c = ctx();
   if (c.ready) return 0;
   g_trace_on = env_flag("HB200_TRACE", false);
   int ndev = 0;
   cudaError_t e = cudaGetDeviceCount(&ndev);
   if (e != cudaSuccess || ndev == 0) {
      return set_error(HB200_ERROR_GENERIC,
                       "hb200_init: no CUDA device available (%s); libhb200 has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
   }
   HB_REQUIRE(device >= 0 && device < ndev, HB200_ERROR_ARG, "hb200_init: bad device index");
   HB_CUDA(cudaSetDevice(device));
   cudaDeviceProp prop;
   HB_CUDA(cudaGetDeviceProperties(&prop, device));
   if (prop.major < 10) {
      return set_error(HB200_ERROR_GENERIC,
                       "hb200_init: device %s is sm_%d%d; libhb200 is built for sm_100a only",
                       prop.name, prop.major, prop.minor);
   }
   c.device = device;
   HB_CUDA(cudaStreamCreateWithFlags(&c.s_comp, cudaStreamNonBlocking));
   HB_CUDA(cudaStreamCreateWithFlags(&c.s_comm, cudaStreamNonBlocking));
   HB_CUDA(cudaEventCreateWithFlags(&c.ev_a, cudaEventDisableTiming));
   HB_CUDA(cudaEventCreateWithFlags(&c.ev_b, cudaEventDisableTiming));
   HB_CUDA(cudaEventCreate(&c.ev_c));
   HB_CUDA(cudaEventCreate(&c.ev_d));
   HB_CUDA(cudaMalloc(&c.d_partials, sizeof(double) * kRedBlocksMax * 4));
   HB_CUDA(cudaMalloc(&c.d_counter, sizeof(unsigned int) * 4));
   HB_CUDA(cudaMemset(c.d_counter, 0, sizeof(unsigned int) * 4));
   HB_CUDA(cudaMalloc(&c.d_scalars, sizeof(double) * kScalarSlots));
   HB_CUDA(cudaMemset(c.d_scalars, 0, sizeof(double) * kScalarSlots));
   HB_CUDA(cudaMallocHost(&c.h_scalars, sizeof(double) * kScalarSlots));
   c.rank = 0;
   c.nranks = 1;
   c.ready = true;
   return 0;
}

int hb200_finalize(void)
{
   Ctx &c = ctx();
   if (!c.ready) return 0;
   cudaDeviceSynchronize();
#ifdef HB200_WITH_NCCL
   if (c.nccl) { nccl_api().CommDestroy(c.nccl); c.nccl = nullptr; }
#endif
   for (int k = 0; k < 16; k++) if (c.ws_ptr[k]) cudaFree(c.ws_ptr[k]);
   for (int r = 0; r < (int) c.peer_arena.size(); r++) {
      if (r != c.rank && c.peer_arena[r]) cudaIpcCloseMemHandle(c.peer_arena[r]);
   }
   if (c.arena) cudaFree(c.arena);
   if (c.h_halo_err) cudaFreeHost(c.h_halo_err);
   cudaFree(c.d_partials);
   cudaFree(c.d_counter);
   cudaFree(c.d_scalars);
   cudaFreeHost(c.h_scalars);
   cudaEventDestroy(c.ev_a);
   cudaEventDestroy(c.ev_b);
   cudaEventDestroy(c.ev_c);
   cudaEventDestroy(c.ev_d);
   cudaStreamDestroy(c.s_comp);
   cudaStreamDestroy(c.s_comm);
   c = Ctx();
   return 0;
}

int hb200_comm_get_unique_id(void *id128)
{
#ifdef HB200_WITH_NCCL
   HB_REQUIRE(id128 != nullptr, HB200_ERROR_ARG, "null id buffer");
   HB_CHECK(nccl_load());
   ncclUniqueId id;
   HB_NCCL(nccl_api().GetUniqueId(&id));
   static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
   memcpy(id128, &id, 128);
   return 0;
#else
   (void) id128;
   return set_error(HB200_ERROR_GENERIC, "libhb200 built without NCCL");
#endif
}

int hb200_comm_init(int rank, int nranks, const void *id128)
{
   HB_CHECK(require_ready());
   Ctx &c = ctx();
   HB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, HB200_ERROR_ARG, "bad rank/nranks");
   if (nranks == 1) { c.rank = 0; c.nranks = 1; return 0; }
#ifdef HB200_WITH_NCCL
   HB_REQUIRE(id128 != nullptr, HB200_ERROR_ARG, "null id buffer");
   HB_CHECK(nccl_load());
   ncclUniqueId id;
   memcpy(&id, id128, 128);
   HB_TRACE("comm_init: ncclCommInitRank %d/%d ...", rank, nranks);
   HB_NCCL(nccl_api().CommInitRank(&c.nccl, nranks, id, rank));
   HB_TRACE("comm_init: done");
   c.rank = rank;
   c.nranks = nranks;
   return 0;
#else
   return set_error(HB200_ERROR_GENERIC, "libhb200 built without NCCL");
#endif
}

int hb200_comm_rank(void) { return ctx().rank; }
int hb200_comm_size(void) { return ctx().nranks; }

int hb200_comm_barrier(void)
{
   HB_CHECK(require_ready());
   Ctx &c = ctx();
   if (c.nranks > 1) {
      HB_CHECK(scalars_allreduce(kScalarSlots - 1, 1, c.s_comp));
   }
   HB_CUDA(cudaStreamSynchronize(c.s_comp));
   return 0;
}

int hb200_set_halo_mode(int mode)
{
   HB_REQUIRE(mode >= 0 && mode <= 2, HB200_ERROR_ARG, "halo mode must be 0 (NCCL), 1 (peer put) or 2 (peer put if possible)");
   Ctx &c = ctx();
   if (mode == 0 || c.nranks <= 1) { c.halo_mode = mode == 2 ? 0 : mode; return 0; }
   HB_CHECK(require_ready());
   // collective: all ranks must make the same call
   int ok = 0;
   HB_TRACE("set_halo_mode(%d): mapping the peer arenas ...", mode);
   HB_CHECK(arena_setup_collective(&ok));
   HB_TRACE("set_halo_mode(%d): peer arenas %s", mode, ok ? "mapped on every rank" : "unavailable, NCCL halo");
   if (!ok && mode == 1) {
      return set_error(HB200_ERROR_GENERIC, "peer halo unavailable on at least one rank: %s", std::string(hb200_last_error()).c_str());
   }
   c.halo_mode = ok ? 1 : 0;
   return 0;
}

int hb200_halo_mode(void) { return ctx().halo_mode; }

int hb200_malloc(void **dev, size_t bytes)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(dev != nullptr, HB200_ERROR_ARG, "null out pointer");
   *dev = nullptr;
   if (bytes == 0) bytes = 8;
   cudaError_t e = cudaMalloc(dev, bytes);
   if (e != cudaSuccess) {
      return set_error(HB200_ERROR_MEMORY, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
   }
   return 0;
}

int hb200_free(void *dev)
{
   if (dev) HB_CUDA(cudaFree(dev));
   return 0;
}

int hb200_memcpy_h2d(void *dev, const void *host, size_t bytes)
{
   HB_CHECK(require_ready());
   if (bytes == 0) return 0;
   HB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx().s_comp));
   HB_CUDA(cudaStreamSynchronize(ctx().s_comp));
   return 0;
}

int hb200_memcpy_d2h(void *host, const void *dev, size_t bytes)
{
   HB_CHECK(require_ready());
   if (bytes == 0) return 0;
   HB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx().s_comp));
   HB_CUDA(cudaStreamSynchronize(ctx().s_comp));
   return 0;
}

int hb200_memcpy_d2d(void *dst, const void *src, size_t bytes)
{
   HB_CHECK(require_ready());
   if (bytes == 0) return 0;
   HB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx().s_comp));
   return 0;
}

int hb200_sync(void)
{
   HB_CHECK(require_ready());
   HB_CUDA(cudaStreamSynchronize(ctx().s_comm));
   HB_CUDA(cudaStreamSynchronize(ctx().s_comp));
   return halo_check_error();
}

void *hb200_compute_stream(void) { return (void *) ctx().s_comp; }

long long hb200_launch_count(int reset)
{
   long long v = ctx().launches;
   if (reset) ctx().launches = 0;
   return v;
}

/* ---- vector C-ABI ---- */
int hb200_vec_set(double *y, double value, size_t n)
{
   HB_CHECK(require_ready());
   return vec_set(y, value, n, ctx().s_comp);
}
int hb200_vec_copy(const double *x, double *y, size_t n)
{
   HB_CHECK(require_ready());
   return vec_copy(x, y, n, ctx().s_comp);
}
int hb200_vec_scale(double alpha, double *y, size_t n)
{
   HB_CHECK(require_ready());
   return vec_scale(alpha, y, n, ctx().s_comp);
}
int hb200_vec_axpy(double alpha, const double *x, double *y, size_t n)
{
   HB_CHECK(require_ready());
   return vec_axpy(alpha, x, y, n, ctx().s_comp);
}
int hb200_vec_inner_prod(const double *x, const double *y, size_t n, double *result)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(result != nullptr, HB200_ERROR_ARG, "null result");
   Ctx &c = ctx();
   const int slot = kScalarSlots - 2;
   HB_CHECK(vec_dot_dev(x, y, n, slot, c.s_comp));
   HB_CHECK(scalars_allreduce(slot, 1, c.s_comp));
   return scalars_fetch(slot, 1, result, c.s_comp);
}
int hb200_vec_pointwise_divpy(const double *x, const double *b, double *y, const int *marker,
                              int marker_val, size_t n)
{
   HB_CHECK(require_ready());
   return vec_divpy(x, b, y, marker, marker_val, n, ctx().s_comp);
}

}  // extern "C"
