// parcsr_peer.cu — halo exchange by direct NVLink peer stores (halo mode 1).
//
// Replaces the pack kernel + grouped ncclSend/ncclRecv of the default path (and, in the
// reference, thrust::gather + D2H + MPI_Isend/Irecv/Waitall + H2D,
// src/parcsr_mv/par_csr_matvec_device.c:201-205, par_csr_communication.c:459-715) by ONE kernel on
// the sending side and ONE on the receiving side, both on the compute stream:
//
//   put  : buf_q[parity][k] = x[send_map_elmts[k]]   written straight into the receiver's HBM over
//          NVLink (IPC-mapped arena), __threadfence_system, last CTA stores the arrival flags;
//   wait : spin on this rank's arrival flags (acquire, system scope), copy the arena buffer into
//          the matrix's private x_ext, store "consumed" flags back to the senders.
//
// The transfer itself overlaps with the diag-block SpMV that sits between the two kernels in the
// stream.  No NCCL launch, no second stream, no events: a coarse-level exchange costs two ~3 us
// kernels instead of a ~25 us NCCL group, and the whole V-cycle stays capturable in a CUDA graph.
// Receive buffers are double-buffered by exchange parity and protected by the consumed flags, so a
// sender can never overwrite data that has not been copied out (also for non-symmetric neighbour
// sets, e.g. interpolation matrices).  Epoch counters live in device memory, so a replayed graph
// keeps counting.
#include "hb_internal.cuh"
#include "hb_peer.cuh"
#include <stdlib.h>
#include <string.h>

namespace hb {

struct PeerPlan {
   int n_out = 0, n_in = 0, total_out = 0, total_in = 0;
   // arena carve-outs of this plan (returned to the free list by peer_plan_free)
   size_t region_off[4] = {0, 0, 0, 0}, region_bytes[4] = {0, 0, 0, 0};
   int    n_regions = 0;
   // outgoing segments
   int                 *d_out_starts = nullptr;   // n_out + 1
   const int           *d_gather = nullptr;       // gather map (borrowed) or NULL = contiguous
   double             **d_out_dst = nullptr;      // [2][n_out] remote data pointers
   unsigned long long **d_out_flag = nullptr;     // [2][n_out] remote arrival flags
   unsigned long long  *acks = nullptr;           // local (arena): consumed-epoch per outgoing segment
   // incoming segments
   double              *in_buf[2] = {nullptr, nullptr};   // local (arena), total_in doubles each
   unsigned long long  *in_flags = nullptr;       // local (arena): [2][n_in] arrival epochs
   unsigned long long **d_in_ack = nullptr;       // [n_in] remote consumed flags (at the senders)
   double              *dst = nullptr;            // private destination buffer (borrowed)
   // device state
   unsigned long long  *d_epoch = nullptr;        // [0] = out epoch, [1] = in epoch, [2] = local relay of the arrival epoch
   unsigned int        *d_ticket = nullptr;       // [0] put, [1] wait
};

// ---------------------------------------------------------------------------------------
// arena
// ---------------------------------------------------------------------------------------
// Collective over all ranks.  Every rank takes part in every collective of this function even
// when its local part failed, so that one rank's failure cannot hang the others; the verdict is
// the sum of the local failure counts.  *all_ok = 1 when every rank mapped every peer arena.
int arena_setup_collective(int *all_ok)
{
   Ctx &c = ctx();
   *all_ok = 0;
   if (c.nranks <= 1) return 0;
   if (c.peer_ok) { *all_ok = 1; return 0; }
#ifdef HB200_WITH_NCCL
   const char *e = getenv("HB200_ARENA_MB");
   const size_t hs = sizeof(cudaIpcMemHandle_t);
   int bad = 0;
   char msg[256] = "";
   c.arena_bytes = (size_t) (e ? atoi(e) : 256) << 20;
   c.arena_used = 0;
   cudaIpcMemHandle_t mine;
   memset(&mine, 0, sizeof(mine));
   if (cudaMalloc((void **) &c.arena, c.arena_bytes) != cudaSuccess ||
       cudaMemset(c.arena, 0, c.arena_bytes) != cudaSuccess ||
       cudaIpcGetMemHandle(&mine, c.arena) != cudaSuccess) {
      snprintf(msg, sizeof(msg), "peer arena allocation / IPC export failed: %s", cudaGetErrorString(cudaGetLastError()));
      bad = 1;
   }
   char *d_h = nullptr;
   HB_CUDA(cudaMalloc((void **) &d_h, hs * (size_t) c.nranks));
   HB_CUDA(cudaMemcpy(d_h + hs * (size_t) c.rank, &mine, hs, cudaMemcpyHostToDevice));
   HB_NCCL(nccl_api().AllGather(d_h + hs * (size_t) c.rank, d_h, hs, ncclChar, c.nccl, c.s_comp));
   HB_CUDA(cudaStreamSynchronize(c.s_comp));
   std::vector<cudaIpcMemHandle_t> all((size_t) c.nranks);
   HB_CUDA(cudaMemcpy(all.data(), d_h, hs * (size_t) c.nranks, cudaMemcpyDeviceToHost));
   cudaFree(d_h);
   // did everybody export?  (opening the handle of a rank whose export failed is undefined)
   double v = (double) bad;
   HB_CUDA(cudaMemcpyAsync(c.d_scalars + kScalarSlots - 1, &v, sizeof(double), cudaMemcpyHostToDevice, c.s_comp));
   HB_CHECK(scalars_allreduce(kScalarSlots - 1, 1, c.s_comp));
   HB_CHECK(scalars_fetch(kScalarSlots - 1, 1, &v, c.s_comp));
   if (v == 0.0) {
      c.peer_arena.assign((size_t) c.nranks, nullptr);
      for (int r = 0; r < c.nranks && !bad; r++) {
         if (r == c.rank) { c.peer_arena[r] = c.arena; continue; }
         void *p = nullptr;
         cudaError_t er = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
         if (er != cudaSuccess) {
            snprintf(msg, sizeof(msg), "cudaIpcOpenMemHandle(rank %d) failed: %s (the peer halo needs all ranks on one "
                     "NVLink domain)", r, cudaGetErrorString(er));
            cudaGetLastError();
            bad = 1;
         }
         c.peer_arena[r] = (char *) p;
      }
      v = (double) bad;
      HB_CUDA(cudaMemcpyAsync(c.d_scalars + kScalarSlots - 1, &v, sizeof(double), cudaMemcpyHostToDevice, c.s_comp));
      HB_CHECK(scalars_allreduce(kScalarSlots - 1, 1, c.s_comp));
      HB_CHECK(scalars_fetch(kScalarSlots - 1, 1, &v, c.s_comp));
   }
   if (v == 0.0) {
      // watchdog of the polling kernels: one host-mapped word the kernels write when they give up
      if (!c.h_halo_err) {
         HB_CUDA(cudaHostAlloc((void **) &c.h_halo_err, sizeof(unsigned long long), cudaHostAllocMapped));
         *c.h_halo_err = 0ull;
         HB_CUDA(cudaHostGetDevicePointer((void **) &c.d_halo_err, (void *) c.h_halo_err, 0));
         const char *ts = getenv("HB200_HALO_TIMEOUT_S");
         const double secs = ts ? atof(ts) : 30.0;
         c.halo_timeout_ns = secs > 0.0 ? (unsigned long long) (secs * 1e9) : 0ull;
      }
      c.peer_ok = true;
      *all_ok = 1;
      return 0;
   }
   // somebody failed: nobody uses the arena
   for (int r = 0; r < c.nranks && r < (int) c.peer_arena.size(); r++) {
      if (r != c.rank && c.peer_arena[r]) cudaIpcCloseMemHandle(c.peer_arena[r]);
   }
   c.peer_arena.clear();
   if (c.arena) { cudaFree(c.arena); c.arena = nullptr; }
   cudaGetLastError();
   if (msg[0]) set_error(HB200_ERROR_GENERIC, "%s", msg);   // kept for hb200_last_error()
   return 0;
#else
   return set_error(HB200_ERROR_GENERIC, "libhb200 built without NCCL");
#endif
}

int arena_setup()
{
   int ok = 0;
   HB_CHECK(arena_setup_collective(&ok));
   if (!ok && ctx().nranks > 1) {
      return set_error(HB200_ERROR_GENERIC, "peer halo unavailable on at least one rank (see the ranks' hb200_last_error)");
   }
   return 0;
}

int arena_alloc(size_t bytes, size_t *offset)
{
   Ctx &c = ctx();
   bytes = (bytes + 255) & ~(size_t) 255;
   // first fit in the regions that destroyed plans gave back
   for (size_t k = 0; k < c.arena_free.size(); k++) {
      if (c.arena_free[k].second >= bytes) {
         *offset = c.arena_free[k].first;
         c.arena_free[k].first += bytes;
         c.arena_free[k].second -= bytes;
         if (c.arena_free[k].second == 0) c.arena_free.erase(c.arena_free.begin() + (long) k);
         return 0;
      }
   }
   const size_t off = (c.arena_used + 255) & ~(size_t) 255;
   if (off + bytes > c.arena_bytes) {
      return set_error(HB200_ERROR_MEMORY, "peer arena exhausted (%zu MB); raise HB200_ARENA_MB", c.arena_bytes >> 20);
   }
   *offset = off;
   c.arena_used = off + bytes;
   return 0;
}

void arena_release(size_t offset, size_t bytes)
{
   Ctx &c = ctx();
   bytes = (bytes + 255) & ~(size_t) 255;
   if (bytes == 0 || !c.arena) return;
   auto &fl = c.arena_free;
   size_t k = 0;
   while (k < fl.size() && fl[k].first < offset) k++;
   fl.insert(fl.begin() + (long) k, std::make_pair(offset, bytes));
   // merge with the neighbours
   if (k + 1 < fl.size() && fl[k].first + fl[k].second == fl[k + 1].first) { fl[k].second += fl[k + 1].second; fl.erase(fl.begin() + (long) k + 1); }
   if (k > 0 && fl[k - 1].first + fl[k - 1].second == fl[k].first) { fl[k - 1].second += fl[k].second; fl.erase(fl.begin() + (long) k); }
   // a free region at the top of the bump pointer lowers it again
   if (!fl.empty() && fl.back().first + fl.back().second == ((c.arena_used + 255) & ~(size_t) 255)) {
      c.arena_used = fl.back().first;
      fl.pop_back();
   }
}

// a polling halo kernel gave up (hb_peer.cuh): report it once per occurrence as an error of the call
int halo_check_error()
{
   Ctx &c = ctx();
   if (!c.h_halo_err) return 0;
   const unsigned long long w = *(volatile unsigned long long *) c.h_halo_err;
   if (w == 0ull) return 0;
   const int what = (int) (w >> 56), seg = (int) ((w >> 40) & 0xffff);
   return set_error(HB200_ERROR_GENERIC,
                    "halo exchange timed out on rank %d: the %s kernel waited more than %.0f s for %s %d to reach epoch %llu "
                    "(a peer died or left the collective sequence); results of this call are undefined",
                    c.rank, what == 1 ? "put" : "wait", (double) c.halo_timeout_ns * 1e-9,
                    what == 1 ? "the consumed-flag of outgoing segment" : "the arrival flag of incoming segment", seg,
                    (unsigned long long) (w & 0xffffffffffull));
}

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
constexpr int kHaloBlock = 512;

__global__ void __launch_bounds__(kHaloBlock)
halo_put_kernel(int total, int n_out, const int *__restrict__ out_starts, const int *__restrict__ gather,
                const double *__restrict__ src, double *const *__restrict__ dst2,
                unsigned long long *const *__restrict__ flag2, const unsigned long long *__restrict__ acks,
                unsigned long long *epoch_ctr, unsigned int *ticket, SpinGuard guard)
{
   __shared__ bool is_last;
   const unsigned long long epoch = epoch_ctr[0] + 1;
   const int par = (int) (epoch & 1ull);
   // the receivers must have copied exchange (epoch - 2) out of this parity's buffers
   if (epoch > 2) {
      for (int i = threadIdx.x; i < n_out; i += blockDim.x) spin_until_ge(acks + i, epoch - 2, guard, 1, i);
      __syncthreads();
   }
   for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
      // segment of entry k (n_out <= a few dozen)
      int lo = 0, hi = n_out - 1;
      while (lo < hi) {
         const int mid = (lo + hi + 1) >> 1;
         if (out_starts[mid] <= k) lo = mid; else hi = mid - 1;
      }
      const double v = gather ? src[gather[k]] : src[k];
      dst2[par * n_out + lo][k - out_starts[lo]] = v;
   }
   // one system-scope fence per CTA (cumulative over the CTA's stores through the barrier); with
   // several CTAs the last one to arrive publishes the arrival flags
   __syncthreads();
   if (threadIdx.x == 0) {
      __threadfence_system();
      if (gridDim.x == 1) {
         is_last = true;
      } else {
         const unsigned int t = atomicInc(ticket, gridDim.x - 1);
         is_last = (t == gridDim.x - 1);
         if (is_last) __threadfence_system();
      }
   }
   __syncthreads();
   if (is_last) {
      for (int i = threadIdx.x; i < n_out; i += blockDim.x) st_release_sys(flag2[par * n_out + i], epoch);
      if (threadIdx.x == 0) epoch_ctr[0] = epoch;
   }
}

__global__ void __launch_bounds__(kHaloBlock)
halo_wait_kernel(int total, int n_in, const double *__restrict__ buf0, const double *__restrict__ buf1,
                 const unsigned long long *__restrict__ flags, unsigned long long *const *__restrict__ in_ack,
                 double *__restrict__ dst, unsigned long long *epoch_ctr, unsigned int *ticket, SpinGuard guard)
{
   __shared__ bool is_last;
   const unsigned long long epoch = epoch_ctr[1] + 1;
   const int par = (int) (epoch & 1ull);
   for (int j = threadIdx.x; j < n_in; j += blockDim.x) spin_until_ge(flags + par * n_in + j, epoch, guard, 2, j);
   __syncthreads();
   const double *buf = par ? buf1 : buf0;
   for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
      dst[k] = __ldcg(buf + k);
   }
   __syncthreads();
   if (threadIdx.x == 0) {
      if (gridDim.x == 1) {
         is_last = true;
      } else {
         __threadfence();
         const unsigned int t = atomicInc(ticket + 1, gridDim.x - 1);
         is_last = (t == gridDim.x - 1);
      }
   }
   __syncthreads();
   if (is_last) {
      // everything of this exchange has been copied out: tell the senders, advance the epoch
      for (int j = threadIdx.x; j < n_in; j += blockDim.x) st_release_sys(in_ack[j], epoch);
      if (threadIdx.x == 0) epoch_ctr[1] = epoch;
   }
}

// small exchanges (the coarse levels) run as one CTA: no ticket, one fence
static inline int halo_grid(int total)
{
   if (total <= kHaloBlock * 8) return 1;
   int grid = (total + kHaloBlock * 4 - 1) / (kHaloBlock * 4);
   return grid > 128 ? 128 : grid;
}

static SpinGuard spin_guard()
{
   SpinGuard g;
   g.err = ctx().d_halo_err;
   g.timeout_ns = ctx().halo_timeout_ns;
   return g;
}

int peer_put(PeerPlan *pl, const double *src, cudaStream_t st)
{
   if (pl->n_out == 0) return 0;
   HB_LAUNCH(halo_put_kernel, halo_grid(pl->total_out), kHaloBlock, 0, st, pl->total_out, pl->n_out, pl->d_out_starts,
             pl->d_gather, src, pl->d_out_dst, pl->d_out_flag, pl->acks, pl->d_epoch, pl->d_ticket, spin_guard());
   HB_LAUNCH_CHECK();
   return 0;
}

int peer_wait(PeerPlan *pl, cudaStream_t st)
{
   if (pl->n_in == 0) return 0;
   HB_LAUNCH(halo_wait_kernel, halo_grid(pl->total_in), kHaloBlock, 0, st, pl->total_in, pl->n_in, pl->in_buf[0],
             pl->in_buf[1], pl->in_flags, pl->d_in_ack, pl->dst, pl->d_epoch, pl->d_ticket, spin_guard());
   HB_LAUNCH_CHECK();
   return 0;
}

bool peer_has_out(const PeerPlan *pl) { return pl && pl->n_out > 0; }

bool peer_wait_args(const PeerPlan *pl, PeerWaitArgs *out)
{
   if (!pl || pl->n_in == 0) return false;
   out->buf0 = pl->in_buf[0]; out->buf1 = pl->in_buf[1];
   out->flags = pl->in_flags; out->in_ack = pl->d_in_ack;
   out->epoch_ctr = pl->d_epoch; out->ticket = pl->d_ticket;
   out->n_in = pl->n_in;
   out->err = ctx().d_halo_err;
   out->timeout_ns = ctx().halo_timeout_ns;
   return true;
}

void peer_fused_args(const PeerPlan *pl, PeerFusedArgs *out)
{
   *out = PeerFusedArgs();
   out->w.err = ctx().d_halo_err;
   out->w.timeout_ns = ctx().halo_timeout_ns;
   if (!pl) return;
   out->w.buf0 = pl->in_buf[0]; out->w.buf1 = pl->in_buf[1];
   out->w.flags = pl->in_flags; out->w.in_ack = pl->d_in_ack;
   out->w.epoch_ctr = pl->d_epoch; out->w.ticket = pl->d_ticket;
   out->w.n_in = pl->n_in;
   out->n_out = pl->n_out; out->total_out = pl->total_out;
   out->out_starts = pl->d_out_starts; out->gather = pl->d_gather;
   out->dst2 = pl->d_out_dst; out->flag2 = pl->d_out_flag; out->acks = pl->acks;
}

void peer_plan_free(PeerPlan *pl)
{
   if (!pl) return;
   // (the caller has synchronised the device; the regions are zeroed again, after a collective step,
   //  by the next plan that takes them: build_plan)
   for (int k = 0; k < pl->n_regions; k++) arena_release(pl->region_off[k], pl->region_bytes[k]);
   if (pl->d_out_starts) cudaFree(pl->d_out_starts);
   if (pl->d_out_dst) cudaFree(pl->d_out_dst);
   if (pl->d_out_flag) cudaFree(pl->d_out_flag);
   if (pl->d_in_ack) cudaFree(pl->d_in_ack);
   if (pl->d_epoch) cudaFree(pl->d_epoch);
   if (pl->d_ticket) cudaFree(pl->d_ticket);
   delete pl;
}

// ---------------------------------------------------------------------------------------
// plan construction (collective over all ranks, same call order everywhere)
// ---------------------------------------------------------------------------------------
// out: segments this rank sends (peer, start offsets);  in: segments it receives.
// Every rank runs every collective of this function even when a local step failed (arena exhausted,
// cudaMalloc failure): the failure count is agreed on at the end, and a plan exists afterwards either
// on every rank or on none (*out_plan = NULL: the matrix keeps the NCCL halo).
static int build_plan(PeerPlan **out_plan, int n_out, const int *out_procs, const int *out_starts,
                      const int *d_gather, int n_in, const int *in_procs, const int *in_starts, double *dst)
{
#ifdef HB200_WITH_NCCL
   Ctx &c = ctx();
   const int nr = c.nranks;
   int bad = 0;
   *out_plan = nullptr;
   PeerPlan *pl = new PeerPlan();
   pl->n_out = n_out; pl->n_in = n_in;
   pl->total_out = n_out ? out_starts[n_out] : 0;
   pl->total_in = n_in ? in_starts[n_in] : 0;
   pl->d_gather = d_gather;
   pl->dst = dst;
   // ---- carve local arena space: 2 receive buffers, arrival flags, consumed flags
   const size_t want[4] = {sizeof(double) * (size_t) (pl->total_in ? pl->total_in : 1),
                           sizeof(double) * (size_t) (pl->total_in ? pl->total_in : 1),
                           sizeof(unsigned long long) * (size_t) (2 * (n_in ? n_in : 1)),
                           sizeof(unsigned long long) * (size_t) (n_out ? n_out : 1)};
   size_t off[4] = {0, 0, 0, 0};
   for (int k = 0; k < 4 && !bad; k++) {
      if (arena_alloc(want[k], &off[k])) { bad = 1; break; }
      pl->region_off[k] = off[k]; pl->region_bytes[k] = want[k]; pl->n_regions = k + 1;
   }
   const size_t off_buf0 = off[0], off_buf1 = off[1], off_flags = off[2], off_acks = off[3];
   pl->in_buf[0] = (double *) (c.arena + off_buf0);
   pl->in_buf[1] = (double *) (c.arena + off_buf1);
   pl->in_flags = (unsigned long long *) (c.arena + off_flags);
   pl->acks = (unsigned long long *) (c.arena + off_acks);
   // ---- publish: row [p] of my table = {buf0, buf1, flag0, flag1 of the segment I receive from
   //      rank p; ack slot of the segment I send to rank p}; -1 = none
   const int W = 5;
   std::vector<long long> mine((size_t) nr * W, -1), all((size_t) nr * nr * W, -1);
   if (!bad) {
      for (int j = 0; j < n_in; j++) {
         const int p = in_procs[j];
         long long *row = &mine[(size_t) p * W];
         row[0] = (long long) (off_buf0 + sizeof(double) * (size_t) in_starts[j]);
         row[1] = (long long) (off_buf1 + sizeof(double) * (size_t) in_starts[j]);
         row[2] = (long long) (off_flags + sizeof(unsigned long long) * (size_t) j);
         row[3] = (long long) (off_flags + sizeof(unsigned long long) * (size_t) (n_in + j));
      }
      for (int i = 0; i < n_out; i++) {
         mine[(size_t) out_procs[i] * W + 4] = (long long) (off_acks + sizeof(unsigned long long) * (size_t) i);
      }
   }
   long long *d_t = nullptr;
   const size_t row_bytes = sizeof(long long) * (size_t) nr * W;
   HB_CUDA(cudaMalloc((void **) &d_t, row_bytes * (size_t) nr));   // (a failure here is fatal for the process anyway)
   HB_CUDA(cudaMemcpy((char *) d_t + row_bytes * (size_t) c.rank, mine.data(), row_bytes, cudaMemcpyHostToDevice));
   HB_NCCL(nccl_api().AllGather((char *) d_t + row_bytes * (size_t) c.rank, d_t, row_bytes, ncclChar, c.nccl, c.s_comp));
   HB_CUDA(cudaStreamSynchronize(c.s_comp));
   HB_CUDA(cudaMemcpy(all.data(), d_t, row_bytes * (size_t) nr, cudaMemcpyDeviceToHost));
   cudaFree(d_t);
   // every rank is past the destruction of whatever owned these regions before (same call order on
   // all ranks, destruction synchronises the device): no late store of an old plan can land any more.
   // Flags and consumed-epochs start from zero.
   if (!bad) {
      if (cudaMemsetAsync(c.arena + off_flags, 0, want[2], c.s_comp) != cudaSuccess ||
          cudaMemsetAsync(c.arena + off_acks, 0, want[3], c.s_comp) != cudaSuccess) bad = 1;
   }
   // ---- resolve remote pointers
   std::vector<double *> dst2((size_t) (2 * (n_out ? n_out : 1)), nullptr);
   std::vector<unsigned long long *> flag2((size_t) (2 * (n_out ? n_out : 1)), nullptr);
   std::vector<unsigned long long *> in_ack((size_t) (n_in ? n_in : 1), nullptr);
   char mismatch[160] = "";
   for (int i = 0; i < n_out && !bad; i++) {
      const int q = out_procs[i];
      const long long *row = &all[((size_t) q * nr + (size_t) c.rank) * W];   // q's segment from me
      if (row[0] < 0) { snprintf(mismatch, sizeof(mismatch), "rank %d offers no receive segment for rank %d", q, c.rank); bad = 1; break; }
      // parity 0 buffers are used by even epochs, parity 1 by odd ones: tables ordered [par][i]
      dst2[(size_t) i] = (double *) (c.peer_arena[q] + row[0]);
      dst2[(size_t) (n_out + i)] = (double *) (c.peer_arena[q] + row[1]);
      flag2[(size_t) i] = (unsigned long long *) (c.peer_arena[q] + row[2]);
      flag2[(size_t) (n_out + i)] = (unsigned long long *) (c.peer_arena[q] + row[3]);
   }
   for (int j = 0; j < n_in && !bad; j++) {
      const int p = in_procs[j];
      const long long a = all[((size_t) p * nr + (size_t) c.rank) * W + 4];         // p's ack slot for me
      if (a < 0) { snprintf(mismatch, sizeof(mismatch), "rank %d offers no consumed-flag for rank %d", p, c.rank); bad = 1; break; }
      in_ack[(size_t) j] = (unsigned long long *) (c.peer_arena[p] + a);
   }
   if (!bad) {
      bool ok = cudaMalloc((void **) &pl->d_out_starts, sizeof(int) * (size_t) (n_out + 1)) == cudaSuccess;
      if (ok && n_out) ok = cudaMemcpy(pl->d_out_starts, out_starts, sizeof(int) * (size_t) (n_out + 1), cudaMemcpyHostToDevice) == cudaSuccess;
      ok = ok && cudaMalloc((void **) &pl->d_out_dst, sizeof(double *) * dst2.size()) == cudaSuccess;
      ok = ok && cudaMemcpy(pl->d_out_dst, dst2.data(), sizeof(double *) * dst2.size(), cudaMemcpyHostToDevice) == cudaSuccess;
      ok = ok && cudaMalloc((void **) &pl->d_out_flag, sizeof(unsigned long long *) * flag2.size()) == cudaSuccess;
      ok = ok && cudaMemcpy(pl->d_out_flag, flag2.data(), sizeof(unsigned long long *) * flag2.size(), cudaMemcpyHostToDevice) == cudaSuccess;
      ok = ok && cudaMalloc((void **) &pl->d_in_ack, sizeof(unsigned long long *) * in_ack.size()) == cudaSuccess;
      ok = ok && cudaMemcpy(pl->d_in_ack, in_ack.data(), sizeof(unsigned long long *) * in_ack.size(), cudaMemcpyHostToDevice) == cudaSuccess;
      ok = ok && cudaMalloc((void **) &pl->d_epoch, sizeof(unsigned long long) * 4) == cudaSuccess;
      ok = ok && cudaMemset(pl->d_epoch, 0, sizeof(unsigned long long) * 4) == cudaSuccess;
      ok = ok && cudaMalloc((void **) &pl->d_ticket, sizeof(unsigned int) * 2) == cudaSuccess;
      ok = ok && cudaMemset(pl->d_ticket, 0, sizeof(unsigned int) * 2) == cudaSuccess;
      if (!ok) { cudaGetLastError(); bad = 1; }
   }
   // ---- agree: peers must not start writing into this arena region before everybody has built the
   //      plan, and nobody may use a plan that some rank could not build
   double v = (double) bad;
   HB_CUDA(cudaMemcpyAsync(c.d_scalars + kScalarSlots - 1, &v, sizeof(double), cudaMemcpyHostToDevice, c.s_comp));
   HB_CUDA(cudaStreamSynchronize(c.s_comp));
   HB_CHECK(scalars_allreduce(kScalarSlots - 1, 1, c.s_comp));
   HB_CHECK(scalars_fetch(kScalarSlots - 1, 1, &v, c.s_comp));
   if (v != 0.0) {
      peer_plan_free(pl);
      if (mismatch[0]) set_error(HB200_ERROR_GENERIC, "halo plan mismatch: %s", mismatch);
      HB_TRACE("peer plan not built (%d rank(s) failed): this matrix keeps the NCCL halo", (int) v);
      return 0;
   }
   *out_plan = pl;
   return 0;
#else
   return set_error(HB200_ERROR_GENERIC, "libhb200 built without NCCL");
#endif
}

int peer_plans_ensure(hb200_parcsr *A, bool reverse)
{
   Ctx &c = ctx();
   CommPkgD &pk = A->pkg;
   if (c.nranks <= 1) return 0;
   bool &tried = reverse ? pk.peer_tried_rev : pk.peer_tried;
   if (tried) return 0;
   tried = true;
   HB_TRACE("peer plan (%s) for a %d x %d block: %d sends, %d recvs ...", reverse ? "reverse" : "forward", A->num_rows,
            A->num_cols, pk.num_sends, pk.num_recvs);
   HB_CHECK(arena_setup());
   PeerPlan *pl = nullptr;
   if (!reverse) {
      // forward (job 1): out = sends (gather through send_map_elmts), in = recvs -> x_ext
      HB_CHECK(build_plan(&pl, pk.num_sends, pk.send_procs.data(), pk.send_map_starts.data(), pk.d_send_map_elmts,
                          pk.num_recvs, pk.recv_procs.data(), pk.recv_vec_starts.data(), pk.d_recv_buf));
      pk.fwd = pl;
   } else {
      // reverse (job 2): out = recv segments of y_tmp (contiguous), in = send segments -> send_buf
      HB_CHECK(build_plan(&pl, pk.num_recvs, pk.recv_procs.data(), pk.recv_vec_starts.data(), nullptr,
                          pk.num_sends, pk.send_procs.data(), pk.send_map_starts.data(), pk.d_send_buf));
      pk.rev = pl;
   }
   // no plan (the ranks agreed that somebody could not build it): this direction of this matrix goes
   // over NCCL, on every rank alike
   if (!pl) { if (reverse) pk.peer_off_rev = true; else pk.peer_off = true; }
   HB_TRACE("peer plan %s", pl ? "built" : "unavailable");
   return 0;
}

}  // namespace hb
