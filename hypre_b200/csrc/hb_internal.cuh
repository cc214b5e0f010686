// hb_internal.cuh — shared declarations of libhb200 (not part of the C-ABI).
//
// One process = one GPU = one Ctx: a compute stream, a comm stream, an optional NCCL
// communicator, and small pinned/device scratch used by the deterministic reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/hb200.h"

#ifdef HB200_WITH_NCCL
#include "nccl_dyn.cuh"
#endif

namespace hb {

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
int set_error(int flag, const char *fmt, ...);
// run-time switch from the environment: unset -> dflt; "", "0", "off", "no" -> false; anything else -> true
bool env_flag(const char *name, bool dflt);

#define HB_CUDA(call)                                                                    \
   do {                                                                                  \
      cudaError_t e_ = (call);                                                           \
      if (e_ != cudaSuccess)                                                             \
         return hb::set_error(HB200_ERROR_GENERIC, "CUDA error %s at %s:%d: %s", #call,  \
                              __FILE__, __LINE__, cudaGetErrorString(e_));               \
   } while (0)

#define HB_CHECK(expr)                                                                   \
   do {                                                                                  \
      int f_ = (expr);                                                                   \
      if (f_) return f_;                                                                 \
   } while (0)

#define HB_REQUIRE(cond, flag, msg)                                                      \
   do {                                                                                  \
      if (!(cond)) return hb::set_error((flag), "%s (%s:%d)", (msg), __FILE__, __LINE__);\
   } while (0)

#ifdef HB200_WITH_NCCL
#define HB_NCCL(call)                                                                    \
   do {                                                                                  \
      ncclResult_t r_ = (call);                                                          \
      if (r_ != ncclSuccess)                                                             \
         return hb::set_error(HB200_ERROR_GENERIC, "NCCL error %s at %s:%d: %s", #call,  \
                              __FILE__, __LINE__, hb::nccl_api().GetErrorString(r_));    \
   } while (0)
#endif

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
constexpr int kNumSMs        = 148;     // B200
constexpr int kRedBlocksMax  = 1184;    // 148 * 8 partials for two-stage reductions
constexpr int kScalarSlots   = 128;     // device scalar slots (dot results, alpha, beta ...)

struct Ctx {
   bool          ready = false;
   int           device = 0;
   cudaStream_t  s_comp = nullptr;   // all kernels
   cudaStream_t  s_comm = nullptr;   // halo pack / NCCL / peer puts
   cudaEvent_t   ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr;
   int           rank = 0, nranks = 1;
#ifdef HB200_WITH_NCCL
   ncclComm_t    nccl = nullptr;
#endif
   int           halo_mode = 0;
   // reduction scratch
   double       *d_partials = nullptr;   // kRedBlocksMax * 4 doubles
   unsigned int *d_counter  = nullptr;   // last-block-done tickets (one per slot)
   double       *d_scalars  = nullptr;   // kScalarSlots doubles
   double       *h_scalars  = nullptr;   // pinned mirror
   long long     launches = 0;
   bool          capturing = false;      // inside CUDA graph capture
   // fused-dot request: the caller arms it right before the operation whose LAST SpMV kernel should
   // also produce <y, dot_req_w> in scalar slot dot_req_slot; the launch site consumes it when the
   // format supports the fused epilogue (last_dot_fused = true), else the caller runs dot_kernel
   bool          dot_req_armed = false;
   const double *dot_req_w = nullptr;
   int           dot_req_slot = -1;
   bool          last_dot_fused = false;
   // NVLink peer arena (halo mode 1): one IPC-shared allocation per rank; halo receive buffers
   // and arrival / consumed flags of every matrix are carved out of it, peers write into it directly
   char         *arena = nullptr;
   size_t        arena_bytes = 0, arena_used = 0;
   std::vector<char *> peer_arena;       // IPC-mapped base of every rank's arena (self = arena)
   bool          peer_ok = false;
   // watchdog of the polling halo kernels (hb_peer.cuh): host-mapped error word + its device alias
   unsigned long long *h_halo_err = nullptr, *d_halo_err = nullptr;
   unsigned long long  halo_timeout_ns = 0;
   // free list of the arena: regions returned by destroyed halo plans, reused first-fit
   std::vector<std::pair<size_t, size_t>> arena_free;   // (offset, bytes)
   // persistent workspace (Krylov work vectors): stable addresses across solves keep the
   // captured V-cycle graphs valid and take cudaMalloc out of the solve
   void         *ws_ptr[16] = {nullptr};
   size_t        ws_bytes[16] = {0};
};
int ws_get(int slot, size_t bytes, double **out);

// optional stream-time attribution (env HB200_TIMERS=1): tick(cat) charges the compute-stream time
// from now on to `cat`; categories mirror the reference's HYPRE_TIMER_IDs (seq_mv/HYPRE_seq_mv.h:85-97)
enum TimerCat { T_OTHER = 0, T_MATVEC_DIAG, T_MATVEC_OFFD, T_HALO_START, T_HALO_WAIT, T_BLAS1, T_ALLREDUCE,
                T_HOST_SYNC, T_GE_SOLVE, T_RELAX_ZERO, T_NUM };
extern bool g_timers_on;
// HB200_TRACE=1: one stderr line per collective set-up step and per solve (finding the rank and the
// step a multi-rank run stops in); off by default, never on the timed path
extern bool g_trace_on;
void trace_line(const char *fmt, ...);
#define HB_TRACE(...) do { if (hb::g_trace_on) hb::trace_line(__VA_ARGS__); } while (0)
void timer_tick_impl(int cat);
inline void timer_tick(int cat) { if (g_timers_on) timer_tick_impl(cat); }
void timers_begin();
void timers_report(const char *what);
// profiler ranges with the reference's own names (hypre_GpuProfilingPushRange, src/utilities/
// device_markers.c:70-123; "AMGCycle", "AMG Level-k", "Relaxation", "Residual", "Restriction",
// "Interpolation", "Coarse solve" in par_cycle.c:119-869, "PCG-Solve" in pcg.c:379): NVTX v3, header-only,
// a no-op unless a tool is attached.  Ranges around captured launches mark the capture, not the replays.
void prof_push(const char *name);
void prof_pop();
struct ProfRange { explicit ProfRange(const char *n) { prof_push(n); } ~ProfRange() { prof_pop(); } };
int arena_setup();                                         // collective, after the NCCL communicator exists
int arena_alloc(size_t bytes, size_t *offset);             // 256-byte aligned carve-out
void arena_release(size_t offset, size_t bytes);           // give a carve-out back (plan destruction)
int halo_check_error();                                    // non-zero when a polling halo kernel gave up
struct PeerPlan;
Ctx &ctx();
int  require_ready();

#ifndef HB200_EMU
#define HB_LAUNCH(kern, grid, block, smem, stream, ...)                                  \
   do {                                                                                  \
      kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                          \
      hb::ctx().launches++;                                                              \
   } while (0)
// dynamic shared memory of a kernel
#define HB_DYN_SHARED(type, name) extern __shared__ type name[]
#else
// -DHB200_EMU: g++ build of these sources for the CPU-only logic tests (never the product build)
#define HB_LAUNCH(kern, grid, block, smem, stream, ...)                                  \
   do {                                                                                  \
      (void) (stream);                                                                   \
      auto hb_args_ = std::make_tuple(__VA_ARGS__);   /* evaluated and copied now */    \
      hb_emu::launch_bound((unsigned) (grid), (unsigned) (block), (size_t) (smem),       \
                           [=]() { std::apply([](auto... a_) { kern(a_...); }, hb_args_); }); \
      hb::ctx().launches++;                                                              \
   } while (0)
#define HB_DYN_SHARED(type, name) type *name = reinterpret_cast<type *>(hb_emu::dyn_smem())
#endif

#define HB_LAUNCH_CHECK()                                                                \
   do {                                                                                  \
      cudaError_t e_ = cudaPeekAtLastError();                                            \
      if (e_ != cudaSuccess)                                                             \
         return hb::set_error(HB200_ERROR_GENERIC, "kernel launch failed at %s:%d: %s",  \
                              __FILE__, __LINE__, cudaGetErrorString(e_));               \
   } while (0)

// ---------------------------------------------------------------------------------------
// device CSR block
// ---------------------------------------------------------------------------------------
// (3, 4, 5 were the stream-v4 / unrolled vector variants of round 1: removed, they lost on every level measured)
enum SpmvKind { SPMV_AUTO = 0, SPMV_VECTOR = 1, SPMV_STREAM = 2, SPMV_SELL = 6, SPMV_PAT = 7, SPMV_VECTOR16 = 8,
                SPMV_BOX = 9 };   // row-pattern format whose patterns are compact 3 x 3 x 3 stencils: register-window kernel

struct DCsr {
   int        nrows = 0, ncols = 0;
   long long  nnz = 0;
   int       *i = nullptr;        // nrows+1
   int       *j = nullptr;        // nnz
   double    *a = nullptr;        // nnz
   short     *j16 = nullptr;      // nnz: column - row, when every entry fits 16 bits (square blocks), else NULL
   // list of rows with at least one entry (hypre_CSRMatrixRownnz, csr_matrix.c:381-397);
   // built for every block, used when it is sparse in rows (offd blocks)
   int       *rownnz = nullptr;
   int        num_rownnz = 0;
   // nnz-balanced partition for the stream kernel: blk_row[b] .. blk_row[b+1]
   int       *blk_row = nullptr;
   int        nblks = 0;
   int        kind = SPMV_VECTOR;
   int        lanes = 1;          // lanes per row (vector kernel: K; stream kernel: phase-2 L)
   // dictionary-packed sliced-ELL copy (kernels_sell.cu), present when the block qualifies:
   // column index = row + off_dict[code] with <= 256 distinct offsets (structured grids), values
   // either coded the same way (<= 256 distinct, constant-coefficient stencils) or raw fp64
   bool       has_sell = false;
   int        sell_nslices = 0;
   long long *sell_ptr = nullptr;      // nslices+1, in entries (multiples of 32)
   unsigned char *sell_cidx = nullptr; // offset codes, [slice][k][lane]
   unsigned char *sell_vidx = nullptr; // value codes, or NULL
   double    *sell_val = nullptr;      // raw values when not coded, or NULL
   int       *sell_offdict = nullptr;  // 256 entries
   double    *sell_valdict = nullptr;  // 256 entries
   int        sell_nd = 0, sell_nv = 0;
   // row-pattern copy (kernels_pat.cu), present when the block has <= 256 distinct rows as lists of
   // (column - row, value): one byte per row + a table of the patterns (constant-coefficient stencils)
   bool       has_pat = false;
   unsigned char *pat_code = nullptr;  // nrows
   int       *pat_base = nullptr;      // nrows: first column of each row (rectangular blocks), NULL = the row itself
   int       *pat_ptr = nullptr;       // pat_npat + 1
   int       *pat_off = nullptr;       // pat_nent
   double    *pat_val = nullptr;       // pat_nent
   int        pat_npat = 0, pat_nent = 0;
   bool       pat_wide = false;        // 16-bit codes in pat_code, table read from global memory
   bool       pat_skips_boundary = false;   // the pattern kernels leave the rows with offd entries to the boundary kernel
   int       *pat_irr = nullptr;       // rows outside the table (code 255), swept by the CSR kernel
   int        pat_nirr = 0;
   long long  pat_irr_nnz = 0;
   // box view of the row-pattern table (kernels_pat.cu, spmv_box): every pattern is a subset of the
   // 27 offsets {dz*sz + dy*sy + dx}, stored diagonal first and then in ascending order
   bool       has_box = false;
   int        box_sy = 0, box_sz = 0, box_p0 = -1;   // strides; the full (27-entry) pattern, -1 = none
   unsigned int *box_mask = nullptr;   // pat_npat presence masks (bit t = slot (dz+1)*9 + (dy+1)*3 + (dx+1))
   double    *box_val = nullptr;       // pat_npat x 27 slot values
   double     box_p0_val[27] = {0};    // the reference pattern's values (kernel argument: constant bank)
   bool       box_uniform = false;     // every pattern = the reference pattern with slots missing
   unsigned int box_ref_mask = 0;      // slots of the reference pattern
   bool       box_geo = false;         // full 27-slot reference pattern and every row misses exactly the neighbours
                                       // outside the box sy x (sz / sy) x (nrows / sz): no row codes needed
   int        max_row_nnz = 0;
   double     avg_row_nnz = 0.0;
   // formats the automatic choice cannot pick for this block are built on first request
   // (dcsr_ensure_formats), not at upload
   bool       defer_j16 = false, defer_sell = false;
};

int  dcsr_upload(DCsr &M, int nrows, int ncols, const int *hi, const int *hj, const double *ha);
int  dcsr_analyze(DCsr &M, const int *hi, const int *hj, const double *ha);   // formats + kernel choice of a block whose CSR is on the device
int  dcsr_free(DCsr &M);
void dcsr_choose_kernel(DCsr &M, int kind, int lanes);
int  dcsr_ensure_formats(DCsr &M, int kind);
int  dcsr_build_partition(DCsr &M, const int *hi);
int  dcsr_build_sell(DCsr &M, const int *hi, const int *hj, const double *ha);   // kernels_sell.cu
int  dcsr_free_sell(DCsr &M);
// host-side result of the row-pattern analysis of one CSR block (kernels_pat.cu)
struct PatHost {
   bool ok = false, square = true, wide = false;
   bool skips_boundary = false;        // rows with offd entries carry the "outside" code and are in no list
   std::vector<unsigned char> code;    // per row: pattern id, 255 = outside the table
   std::vector<unsigned short> code16; // wide variant: 65535 = outside the table
   std::vector<int> base;              // per row first column (rectangular blocks only)
   std::vector<int> ptr, off;          // table: ptr[npat+1], off[nent]
   std::vector<double> val;            // table: val[nent]
   std::vector<int> irr;               // rows with code 255
   long long irr_nnz = 0;
};
int  pat_analyze_host(int nrows, int ncols, const int *hi, const int *hj, const double *ha, PatHost &out, bool wide);
extern const int *g_pat_boundary_offd_i;   // set around the upload of a diag block whose ParCSR matrix has an offd block
int  dcsr_build_pat(DCsr &M, const int *hi, const int *hj, const double *ha);    // kernels_pat.cu

int  dcsr_free_pat(DCsr &M);
// host-side transpose (stable: entries of each output row in ascending source-row order,
// the order hypre_CSRMatrixMatvecTHost accumulates in, csr_matvec.c:1095-1110)
void host_csr_transpose(int nrows, int ncols, const int *ai, const int *aj, const double *aa,
                        std::vector<int> &ti, std::vector<int> &tj, std::vector<double> &ta);

// ---------------------------------------------------------------------------------------
// SpMV epilogues.  One struct, a compile-time kind: keeps the template count small.
// ---------------------------------------------------------------------------------------
enum EpiKind {
   EPI_AXPBY = 0,     // y = beta*b + alpha*sum            (beta == 0: b not read)
   EPI_ACC,           // y += alpha*sum                    (offd pass of EPI_AXPBY)
   EPI_JACOBI7,       // y = u + (w*f - w*sum)/d  [marked]  (par_relax.c:1216-1244)
   EPI_JACOBI7_ACC,   // y -= (w*sum)/d           [marked]  (offd pass)
   EPI_JACOBI_CORE,   // y = (1-w*skip)*u + w*(f - sum)/d, only if d != 0 [marked] (par_relax.c:258-295)
   EPI_JACOBI_CORE_ACC,
   // Chebyshev smoothing fused into its SpMVs (par_cheby_solve.c:284-341; blocks without an offd part):
   EPI_CHEBY_FIRST,   // r = [ds*](f - sum); u1 = r*c; [t = ds*u1]; last: y = orig_u + [ds*]u1
   EPI_CHEBY_STEP     // u' = mult*r + [ds*]sum; [t = ds*u']; last: y = orig_u + [ds*]u'
};

struct EpiArgs {
   double        alpha = 1.0, beta = 0.0, w = 1.0;
   const double *b = nullptr;      // AXPBY: b;  JACOBI: f
   const double *u = nullptr;      // JACOBI: old u
   const double *d = nullptr;      // JACOBI: l1 norms or NULL (use diagonal)
   const int    *cf = nullptr;     // CF marker or NULL
   int           relax_points = 0;
   int           skip_diag = 0;
   double       *y = nullptr;      // output
   double       *y2 = nullptr;     // second output (Chebyshev)
   const double *r = nullptr;      // Chebyshev: r (read by the steps)
   double       *r_out = nullptr;  // Chebyshev first step: r is written here
   int           cheby_last = 0;   // this SpMV is the last of the sweep: y = orig_u (ea.u) + [ds*]u'
   // optional fused dot: partial[block] = sum_rows y[row]*dotw[row] (deterministic 2-stage)
   const double *dotw = nullptr;
   int           dot_slot = -1;
};

// y-type epilogue launchers (kernels_spmv.cu).  rows_list: optional compressed row list.
int spmv_launch(const DCsr &M, const double *x, int epi_kind, const EpiArgs &ea,
                bool use_rownnz, cudaStream_t st);
int spmv_sell_launch(const DCsr &M, const double *x, int epi_kind, const EpiArgs &ea, cudaStream_t st);
int spmv_pat_launch(const DCsr &M, const double *x, int epi_kind, const EpiArgs &ea, cudaStream_t st);
int spmv_box_launch(const DCsr &M, const double *x, int epi_kind, const EpiArgs &ea, cudaStream_t st);
bool spmv_box_supports(int epi_kind);
bool spmv_can_fuse_dot(const DCsr &M, int epi_kind);   // row-pattern format, no rows outside the table
bool fused_dots_enabled();                              // HB200_FUSED_DOTS=1 turns them on

// ---------------------------------------------------------------------------------------
// BLAS-1 (kernels_blas1.cu).  Scalars live in ctx().d_scalars[slot]; dots are two-stage,
// last-block-reduces, fixed order => bitwise reproducible run to run.
// ---------------------------------------------------------------------------------------
int vec_set(double *y, double v, size_t n, cudaStream_t st);
int vec_copy(const double *x, double *y, size_t n, cudaStream_t st);
int vec_scale(double a, double *y, size_t n, cudaStream_t st);
int vec_axpy(double a, const double *x, double *y, size_t n, cudaStream_t st);
int vec_axpby_out(double a, const double *x, double b, const double *y, double *z, size_t n, cudaStream_t st);
int vec_divpy(const double *x, const double *b, double *y, const int *marker, int mval, size_t n, cudaStream_t st);
int vec_scale_div(double w, const double *f, const double *d, double *u, const int *marker, int mval,
                  const double *u_old, size_t n, cudaStream_t st);   // u = (w*f)/d (zero-guess Jacobi)
int vec_diag_scale(const double *diag, const double *x, double *y, size_t n, cudaStream_t st);
// dot into device slot (local part only)
int vec_dot_dev(const double *x, const double *y, size_t n, int slot, cudaStream_t st);
// <Z_k, x> for k < nv <= kMassNV vectors of a slab in one pass (vector_batched.c); the reduction scratch holds 4 columns
constexpr int kMassNV = 4;
int vec_mass_dot_dev(const double *x, const double *Z, size_t zstride, int nv, size_t n, int slot0, int dump,
                     cudaStream_t st);
// two dots at once: slot0 = <x,y>, slot1 = <z,z2>
int vec_dot2_dev(const double *x, const double *y, const double *z, const double *z2, size_t n,
                 int slot0, int slot1, cudaStream_t st);
// finalize a fused dot whose per-block partials were written by another kernel
int dot_finalize(int nblocks, int slot, cudaStream_t st);
// allreduce (sum) of `count` consecutive device scalar slots over all ranks (NCCL); no-op for 1 rank
int scalars_allreduce(int slot, int count, cudaStream_t st);
// copy `count` slots to pinned host and wait
int scalars_fetch(int slot, int count, double *out, cudaStream_t st);

// PCG fused updates (krylov.cu uses them)
//   x += alpha*p; r -= alpha*s; slot_rr = <r,r>   with alpha = S[slot_gamma]/S[slot_sdotp]
int pcg_update_xr(const double *p, const double *s, double *x, double *r, size_t n,
                  int slot_gamma, int slot_sdotp, int slot_rr, int slot_flag, int skip_break,
                  cudaStream_t st);
//   p = s + beta*p with beta = S[slot_num]/S[slot_den]
int pcg_update_p(const double *s, double *p, size_t n, int slot_num, int slot_den, cudaStream_t st);

// ---------------------------------------------------------------------------------------
// ParCSR
// ---------------------------------------------------------------------------------------
struct CommPkgD {
   int num_sends = 0, num_recvs = 0;
   std::vector<int> send_procs, send_map_starts, send_map_elmts, recv_procs, recv_vec_starts;
   int    *d_send_map_elmts = nullptr;
   double *d_send_buf = nullptr;      // send_map_starts[num_sends]
   double *d_recv_buf = nullptr;      // = x_ext, num_cols_offd
   // MatvecT unpack: CSR-of-E grouping send entries by target row (deterministic)
   int    *d_unpack_rows = nullptr, *d_unpack_ptr = nullptr, *d_unpack_idx = nullptr;
   int     n_unpack_rows = 0;
   // peer-put halo (halo mode 1): one plan per direction, built collectively at first use
   struct PeerPlan *fwd = nullptr, *rev = nullptr;
   bool    peer_tried = false, peer_tried_rev = false;
   bool    peer_off = false, peer_off_rev = false;   // the ranks agreed that the plan cannot be built: NCCL for this matrix
};

}  // namespace hb

struct hb200_parcsr {
   int      num_rows = 0, num_cols = 0, num_cols_offd = 0;
   int64_t  first_row = 0, first_col = 0, global_rows = 0, global_cols = 0;
   hb::DCsr diag, offd;
   bool     has_T = false;
   hb::DCsr diagT, offdT;            // built lazily for MatvecT (restriction)
   std::vector<int64_t> col_map_offd;
   int64_t *d_col_map_offd = nullptr;
   hb::CommPkgD pkg;
   double  *d_ytmp = nullptr;        // MatvecT: offd^T x (num_cols_offd)
   int     *d_unpack_slot = nullptr; // MatvecT, fused kernel: coarse row -> position in the unpack plan, -1 = none
   double  *d_diaginv = nullptr;     // lazily extracted diagonal (DIAGSCALE precond, Jacobi type 0)
   // kept host copies of the CSR (needed to build transposes / level schedules lazily)
   std::vector<int> h_diag_i, h_diag_j, h_offd_i, h_offd_j;
   std::vector<double> h_diag_a, h_offd_a;
   bool     keep_host = true;
   // level schedule of the diag block for hybrid Gauss-Seidel (relax.cu), built lazily
   void    *gs_sched = nullptr;
   int      gs_chunks = 0;            // hybrid GS: the reference's thread count (0 / 1: sequential sweep)
};

namespace hb {
int parcsr_read_binary_exact(hb200_parcsr **A, const char *prefix);   // ij.cu: a matrix written by hb200_parcsr_print_ij_binary, entry order kept
int parcsr_ensure_T(hb200_parcsr *A);
int parcsr_diag(hb200_parcsr *A, const double **diag);   // lazily extracted diagonal of the diag block
int parcsr_halo_begin(hb200_parcsr *A, const double *x, cudaStream_t st_comp);   // pack + exchange on s_comm
int parcsr_halo_end(hb200_parcsr *A, cudaStream_t st_comp);                     // make s_comp wait
int parcsr_matvec(hb200_parcsr *A, double alpha, const double *x, double beta, const double *b,
                  double *y);
int parcsr_matvecT(hb200_parcsr *A, double alpha, const double *x, double beta, double *y);
int gs_sched_free(void *p);
// peer-put halo (parcsr_peer.cu)
int  arena_setup_collective(int *all_ok);                    // collective; never hangs on a local failure
bool peer_has_out(const PeerPlan *pl);
int  peer_plans_ensure(hb200_parcsr *A, bool reverse);      // collective, lazy: builds A->pkg.fwd or .rev
int  peer_put(PeerPlan *pl, const double *src, cudaStream_t st);
int  peer_wait(PeerPlan *pl, cudaStream_t st);
// what a kernel needs to take over the job of halo_wait_kernel (the fused wait + offd pass,
// HB200_FUSE_WAIT=1): arrival flags, the two receive buffers, the senders' ack slots, epoch, ticket
struct PeerWaitArgs {
   const double *buf0 = nullptr, *buf1 = nullptr;
   const unsigned long long *flags = nullptr;
   unsigned long long *const *in_ack = nullptr;
   unsigned long long *epoch_ctr = nullptr;
   unsigned int *ticket = nullptr;
   int n_in = 0;
   unsigned long long *err = nullptr;      // watchdog (hb_peer.cuh SpinGuard)
   unsigned long long timeout_ns = 0;
};
bool peer_wait_args(const PeerPlan *pl, PeerWaitArgs *out);   // false when the plan receives nothing
// everything one kernel needs to run a complete halo exchange itself (kernels_offd.cu, parcsr_fused):
// the put half (gather map, remote buffers and arrival flags, consumed-flags of the receivers) and
// the wait half (PeerWaitArgs)
struct PeerFusedArgs {
   PeerWaitArgs w;
   int n_out = 0, total_out = 0;
   const int *out_starts = nullptr, *gather = nullptr;
   double *const *dst2 = nullptr;
   unsigned long long *const *flag2 = nullptr;
   const unsigned long long *acks = nullptr;
};
void peer_fused_args(const PeerPlan *pl, PeerFusedArgs *out);
// one launch = put + diag pass + wait + offd pass + epilogue of a ParCSR operation on a latency-bound
// level; *done = false when the block does not qualify (the caller then runs the separate kernels)
int  parcsr_fused_try(hb200_parcsr *A, const double *x, int epi_kind, const EpiArgs &ea, bool *done);
int  parcsr_fusedT_try(hb200_parcsr *A, double alpha, const double *x, double beta, double *y, bool *done);
// the boundary half of a split ParCSR operation: complete rows (diag part, offd part, one epilogue) over the
// non-empty-row list of the offd block.  peer = true: with the put in front and the flag poll inside (side
// stream, beside the main kernel); false: x_ext has already arrived in the matrix's receive buffer
int  parcsr_boundary_launch(hb200_parcsr *A, const double *x, int epi_kind, const EpiArgs &ea, bool peer, cudaStream_t st);
bool parcsr_main_skips_boundary(const hb200_parcsr *A);
int  spmv_offd_wait_launch(const DCsr &M, const PeerWaitArgs &w, int epi_kind, const EpiArgs &ea, cudaStream_t st);
int  parcsr_offd_pass(hb200_parcsr *A, int epi_kind, const EpiArgs &ea);   // halo_end + offd SpMV (fused or not)
void peer_plan_free(PeerPlan *pl);
}  // namespace hb
