// amg.cu — device-resident BoomerAMG hierarchy and the multigrid cycle.
//
// Reference: hypre_BoomerAMGCycle (src/parcsr_ls/par_cycle.c:23-889; level-counter state
// machine :225-243, 694-865), hypre_BoomerAMGSolve (src/parcsr_ls/par_amg_solve.c:22-424),
// hypre_ParAMGData (src/parcsr_ls/par_amg.h:18-310).
//
// Differences from the reference that do not change results:
//  * work vectors are per-level arenas (F, U, U_alt, Vtemp) instead of fine-sized temporaries
//    that are logically resized per level (par_cycle.c:307-311);
//  * Jacobi-type sweeps are out of place (one fused SpMV pass); U[l] ping-pongs between two
//    buffers and at level 0 the starting buffer is chosen so that the last sweep lands in the
//    caller's vector (no copy back);
//  * hypre_ParVectorSetZeros(U[l+1]) (par_cycle.c:742) only sets the all-zeros flag: the first
//    sweep on the coarse level is the SpMV-free element-wise form (par_relax.c:1221-1228), so
//    the memset would be dead traffic;
//  * restriction uses a stored transpose of P (deterministic row-parallel SpMV).
#include "hb_internal.cuh"
#include <stdlib.h>
#include "relax.cuh"
#include <errno.h>
#include <math.h>
#include <sys/stat.h>

struct hb200_amg_level {
   hb200_parcsr *A = nullptr, *P = nullptr;
   int     n = 0;
   double *l1 = nullptr;
   int    *cf = nullptr;
   double  relax_weight = 1.0, omega = 1.0;
   double *F = nullptr, *U = nullptr, *Ualt = nullptr, *Vtemp = nullptr;
   double *cheby_ds = nullptr;
   std::vector<double> cheby_coefs;
   int     cheby_order_set = 0;
   double *cheby_work = nullptr;   // 4*n, allocated on demand
   // run-time state
   double *u_cur = nullptr;
   bool    u_zero = false;
};

struct hb200_amg {
   int num_levels = 0;
   std::vector<hb200_amg_level> lev;
   int num_grid_sweeps[4] = {1, 1, 1, 1};
   int grid_relax_type[4] = {18, 18, 18, 9};
   int relax_order = 0, cycle_type = 1, fcycle = 0;
   int cheby_order = 2, cheby_scale = 1, cheby_variant = 0;
   int user_relax_type = -1;
   double tol = 0.0;
   int min_iter = 0, max_iter = 1, converge_type = 0;
   hb::GEData ge;
   bool has_ge = false;
   std::vector<double> ge_A_mat;      // the dense coarse matrix as it was handed over (hb200_amg_save)
   bool owns_matrices = false;        // a loaded hierarchy (hb200_amg_load) owns its level matrices
   bool use_graph = false;
   // fused-dot request of the caller (Krylov preconditioner use): <u, f> of the cycle's result in
   // this scalar slot, produced by the last level-0 sweep when it can; dot_fused reports it
   int  dot_req_slot = -1;
   bool dot_fused = false;
   // graph cache: one executable graph per (f, u, zero flag) the cycle has been called with
   struct GraphEntry {
      const double *f; double *u; int zero;
      cudaGraphExec_t exec; long long launches;
      int dot_slot; bool dot_fused;
   };
   std::vector<GraphEntry> graphs;
   long long cycles_run = 0, cycles_run_dot = 0;
};

namespace hb {

void amg_drop_graphs(hb200_amg *amg);

static int level_alloc(hb200_amg_level &L, bool level0)
{
   const size_t n = (size_t) (L.n ? L.n : 1);
   if (!level0) {
      if (!L.F) HB_CUDA(cudaMalloc(&L.F, sizeof(double) * n));
      if (!L.U) { HB_CUDA(cudaMalloc(&L.U, sizeof(double) * n)); HB_CUDA(cudaMemset(L.U, 0, sizeof(double) * n)); }
   }
   if (!L.Ualt) { HB_CUDA(cudaMalloc(&L.Ualt, sizeof(double) * n)); HB_CUDA(cudaMemset(L.Ualt, 0, sizeof(double) * n)); }
   if (!L.Vtemp) HB_CUDA(cudaMalloc(&L.Vtemp, sizeof(double) * n));
   return 0;
}

// number of out-of-place (buffer-swapping) sweeps level 0 will see in one cycle
static int level0_swaps(const hb200_amg *amg, bool zero_guess)
{
   auto sweeps_for = [&](int cycle_param, int relax_type, int nsweep, bool zero) {
      int swaps = 0;
      if (!relax_is_jacobi(relax_type)) return 0;
      for (int j = 0; j < nsweep; j++) {
         const int npts = (amg->relax_order == 1 && cycle_param < 3) ? 2 : 1;
         for (int q = 0; q < npts; q++) {
            const int rp = npts == 2 ? 1 : 0;   // only whether it is non-zero matters
            const bool form7 = (relax_type == 7) || (relax_type == 18 && rp == 0);
            if (zero && form7) { zero = false; continue; }   // element-wise, writes in place
            zero = false;
            swaps++;
         }
      }
      return swaps;
   };
   if (amg->num_levels == 1) {
      int rt = amg->user_relax_type == -1 ? 6 : amg->user_relax_type;
      return sweeps_for(3 /* relax_points forced 0 */, rt, amg->num_grid_sweeps[0], zero_guess);
   }
   int s = sweeps_for(1, amg->grid_relax_type[1], amg->num_grid_sweeps[1], zero_guess);
   s += sweeps_for(2, amg->grid_relax_type[2], amg->num_grid_sweeps[2], false);
   return s;
}

static int do_relax(hb200_amg *amg, int level, int relax_type, int relax_order, int cycle_param,
                    bool force_no_cf, bool last_of_cycle = false)
{
   hb200_amg_level &L = amg->lev[level];
   Ctx &c = ctx();
   const double *f = L.F;
   if (relax_type == 9 || relax_type == 19) {
      HB_REQUIRE(amg->has_ge, HB200_ERROR_GENERIC, "coarse relax type 9 but no GE data was set");
      ProfRange pr("Coarse solve");
      HB_CHECK(ge_solve(amg->ge, f, L.u_cur));
      L.u_zero = false;
      return 0;
   }
   ProfRange pr_relax("Relaxation");
   if (relax_type == 16) {
      HB_REQUIRE(!L.cheby_coefs.empty(), HB200_ERROR_GENERIC, "Chebyshev relaxation but no coefficients set");
      if (!L.cheby_work) HB_CUDA(cudaMalloc(&L.cheby_work, sizeof(double) * 4 * (size_t) (L.n ? L.n : 1)));
      if (L.u_zero) HB_CHECK(vec_set(L.u_cur, 0.0, (size_t) L.n, c.s_comp));
      const size_t n = (size_t) (L.n ? L.n : 1);
      HB_CHECK(cheby_solve(L.A, f, L.cheby_ds, L.cheby_coefs.data(), amg->cheby_order,
                           amg->cheby_scale, L.u_cur, L.cheby_work, L.cheby_work + n,
                           L.cheby_work + 2 * n, L.cheby_work + 3 * n));
      L.u_zero = false;
      return 0;
   }
   // hypre_BoomerAMGRelaxIF (par_relax_interface.c:19-65)
   int pts[2] = {0, 0};
   int npts = 1;
   if (!force_no_cf && relax_order == 1 && cycle_param < 3) {
      npts = 2;
      pts[0] = cycle_param < 2 ? 1 : -1;
      pts[1] = -pts[0];
   }
   for (int q = 0; q < npts; q++) {
      if (relax_is_jacobi(relax_type)) {
         double *other = (L.u_cur == L.Ualt) ? L.U : L.Ualt;
         bool shortcut = false;
         const bool form7 = (relax_type == 7) || (relax_type == 18 && pts[q] == 0);
         if (last_of_cycle && q == npts - 1 && amg->dot_req_slot >= 0) {
            // the sweep that writes the cycle's result: ask it for <u, f> as well
            c.dot_req_armed = true; c.dot_req_w = f; c.dot_req_slot = amg->dot_req_slot;
         }
         if (L.u_zero && form7) {
            // element-wise zero-guess sweep writes straight into the current buffer
            HB_CHECK(relax_jacobi_oop(L.A, f, L.cf, relax_type, pts[q], L.relax_weight, L.l1,
                                      nullptr, L.u_cur, true, &shortcut));
         } else {
            HB_CHECK(relax_jacobi_oop(L.A, f, L.cf, relax_type, pts[q], L.relax_weight, L.l1,
                                      L.u_cur, other, L.u_zero, &shortcut));
            L.u_cur = other;
         }
         if (last_of_cycle && q == npts - 1) amg->dot_fused = c.last_dot_fused;
      } else if (relax_is_gs(relax_type)) {
         if (L.u_zero) HB_CHECK(vec_set(L.u_cur, 0.0, (size_t) L.n, c.s_comp));
         HB_CHECK(relax_hybrid_gs(L.A, f, L.cf, relax_type, pts[q], L.relax_weight, L.omega, L.l1,
                                  L.u_cur, L.Vtemp));
      } else {
         return set_error(HB200_ERROR_ARG, "relax type %d is not on the B200 path", relax_type);
      }
      L.u_zero = false;   // par_relax.c:170
   }
   return 0;
}

static int cycle_body(hb200_amg *amg, const double *f_dev, double *u_dev, bool u_all_zeros)
{
   Ctx &c = ctx();
   const int nl = amg->num_levels;
   amg->dot_fused = false;
   hb200_amg_level &L0 = amg->lev[0];
   // level-0 aliases (par_amg_solve.c:108-110)
   L0.F = (double *) f_dev;
   L0.U = u_dev;
   const int swaps = level0_swaps(amg, u_all_zeros);
   if (u_all_zeros && (swaps & 1)) L0.u_cur = L0.Ualt; else L0.u_cur = u_dev;
   L0.u_zero = u_all_zeros;
   if (u_all_zeros && L0.u_cur != u_dev) {
      // nothing to initialise: the first sweep overwrites the buffer (zero flag)
   }
   for (int l = 1; l < nl; l++) { amg->lev[l].u_cur = amg->lev[l].U; amg->lev[l].u_zero = false; }

   ProfRange pr_cycle("AMGCycle");
   char lvl_name[32];
   // ---- the reference's cycling state machine (par_cycle.c:225-243, 300-865) ----
   std::vector<int> lev_counter(nl);
   lev_counter[0] = 1;
   for (int k = 1; k < nl; k++) lev_counter[k] = amg->fcycle ? 1 : amg->cycle_type;
   int fcycle_lev = nl - 2;
   int level = 0, cycle_param = 1;
   bool not_finished = true;
   while (not_finished) {
      int num_sweep, relax_type;
      bool one_level = false;
      snprintf(lvl_name, sizeof(lvl_name), "AMG Level-%d", level);
      ProfRange pr_level(lvl_name);
      if (nl > 1) {
         num_sweep = amg->num_grid_sweeps[cycle_param];
         relax_type = amg->grid_relax_type[cycle_param];
      } else {
         num_sweep = amg->num_grid_sweeps[0];
         relax_type = amg->user_relax_type == -1 ? 6 : amg->user_relax_type;
         one_level = true;
      }
      for (int j = 0; j < num_sweep; j++) {
         // the cycle ends with the post-smoothing of level 0 (cycle_param 2): its last sweep writes the result
         const bool last = (nl > 1 && level == 0 && cycle_param == 2 && j == num_sweep - 1);
         HB_CHECK(do_relax(amg, level, relax_type, amg->relax_order, cycle_param, one_level, last));
      }
      --lev_counter[level];
      if (lev_counter[level] >= 0 && level != nl - 1) {
         // go down: residual + restriction (par_cycle.c:742-790)
         hb200_amg_level &Lf = amg->lev[level];
         hb200_amg_level &Lc = amg->lev[level + 1];
         Lc.u_cur = Lc.U;
         Lc.u_zero = true;                     // hypre_ParVectorSetZeros(U_array[coarse])
         {
            ProfRange pr("Residual");
            if (Lf.u_zero) {
               // r = f - A*0 : no relaxation happened on this level (num_sweeps == 0)
               HB_CHECK(vec_copy(Lf.F, Lf.Vtemp, (size_t) Lf.n, c.s_comp));
            } else {
               HB_CHECK(parcsr_matvec(Lf.A, -1.0, Lf.u_cur, 1.0, Lf.F, Lf.Vtemp));
            }
         }
         {
            ProfRange pr("Restriction");
            HB_CHECK(parcsr_matvecT(Lf.P, 1.0, Lf.Vtemp, 0.0, Lc.F));
         }
         ++level;
         lev_counter[level] = lev_counter[level] > amg->cycle_type ? lev_counter[level] : amg->cycle_type;
         cycle_param = (level == nl - 1) ? 3 : 1;
      } else if (level != 0) {
         // go up: interpolation u_f += P u_c (par_cycle.c:815-843)
         hb200_amg_level &Lf = amg->lev[level - 1];
         hb200_amg_level &Lc = amg->lev[level];
         ProfRange pr("Interpolation");
         if (Lc.u_zero) HB_CHECK(vec_set(Lc.u_cur, 0.0, (size_t) Lc.n, c.s_comp));
         if (Lf.u_zero) { HB_CHECK(vec_set(Lf.u_cur, 0.0, (size_t) Lf.n, c.s_comp)); }
         HB_CHECK(parcsr_matvec(Lf.P, 1.0, Lc.u_cur, 1.0, Lf.u_cur, Lf.u_cur));
         Lf.u_zero = false;
         --level;
         cycle_param = 2;
         if (amg->fcycle && fcycle_lev == level) {
            lev_counter[level] = lev_counter[level] > 1 ? lev_counter[level] : 1;
            fcycle_lev--;
         }
      } else {
         not_finished = false;
      }
   }
   if (L0.u_zero) {
      // no sweep touched u (all sweep counts 0 and one level): materialise the zeros
      HB_CHECK(vec_set(L0.u_cur, 0.0, (size_t) L0.n, c.s_comp));
      L0.u_zero = false;
   }
   if (L0.u_cur != u_dev) {
      HB_CHECK(vec_copy(L0.u_cur, u_dev, (size_t) L0.n, c.s_comp));
      L0.u_cur = u_dev;
   }
   return 0;
}

int amg_cycle(hb200_amg *amg, const double *f_dev, double *u_dev, bool u_all_zeros)
{
   Ctx &c = ctx();
   // multi-rank: only the peer-put halo keeps every step of the cycle a plain kernel on one stream
   // (HB200_GRAPH_NCCL=1, opt-in until it has run on hardware: capture the cycle with the NCCL halo as well —
   // grouped ncclSend/ncclRecv on the comm stream are capturable through the same event fork / join)
   static const bool graph_nccl = env_flag("HB200_GRAPH_NCCL", false);
   if (!amg->use_graph || g_timers_on || (c.nranks > 1 && c.halo_mode != 1 && !graph_nccl)) {
      return cycle_body(amg, f_dev, u_dev, u_all_zeros);
   }
#ifdef HB200_EMU
   // g++ test build: graphs are recorded and replayed on one rank only (the emulated NCCL is not stream-aware)
   if (c.nranks > 1) return cycle_body(amg, f_dev, u_dev, u_all_zeros);
#endif
   // CUDA-graph path: the topology of a cycle is fixed by the hierarchy, so capture once per
   // (f, u, zero flag) and replay; removes the launch latency of the ~60 small coarse-level
   // kernels.
   for (auto &g : amg->graphs) {
      if (g.f == f_dev && g.u == u_dev && g.zero == (int) u_all_zeros && g.dot_slot == amg->dot_req_slot) {
         HB_CUDA(cudaGraphLaunch(g.exec, c.s_comp));
         c.launches += g.launches;
         amg->dot_fused = g.dot_fused;
         return 0;
      }
   }
   // the first cycle of a hierarchy runs eagerly: it builds the lazily allocated scratch
   // (Chebyshev work vectors, GS wavefront schedules), which cannot happen under capture
   // (the same holds for the first cycle that carries a fused-dot request: its kernel variants bind
   // their scratch pointers and launch attributes on first use)
   long long &runs = amg->dot_req_slot >= 0 ? amg->cycles_run_dot : amg->cycles_run;
   if (runs++ == 0) return cycle_body(amg, f_dev, u_dev, u_all_zeros);
   for (int l = 0; l + 1 < amg->num_levels; l++) HB_CHECK(parcsr_ensure_T(amg->lev[l].P));
   if (amg->graphs.size() >= 128) {   // (GMRES(k) presents k + 1 different (f, u) pairs)
      cudaGraphExecDestroy(amg->graphs.front().exec);
      amg->graphs.erase(amg->graphs.begin());
   }
   const long long before = c.launches;
   HB_TRACE("amg_cycle: capturing the V-cycle graph (%d levels, zero guess %d)", amg->num_levels, (int) u_all_zeros);
   cudaGraph_t graph = nullptr;
   HB_CUDA(cudaStreamBeginCapture(c.s_comp, cudaStreamCaptureModeThreadLocal));
   c.capturing = true;
   int f = cycle_body(amg, f_dev, u_dev, u_all_zeros);
   c.capturing = false;
   cudaError_t e = cudaStreamEndCapture(c.s_comp, &graph);
   if (f) { if (graph) cudaGraphDestroy(graph); return f; }
   if (e != cudaSuccess) return set_error(HB200_ERROR_GENERIC, "graph capture failed: %s", cudaGetErrorString(e));
   hb200_amg::GraphEntry ge;
   ge.f = f_dev; ge.u = u_dev; ge.zero = (int) u_all_zeros;
   ge.dot_slot = amg->dot_req_slot; ge.dot_fused = amg->dot_fused;
   ge.launches = c.launches - before;
   c.launches = before;
   e = cudaGraphInstantiate(&ge.exec, graph, 0);
   cudaGraphDestroy(graph);
   if (e != cudaSuccess) return set_error(HB200_ERROR_GENERIC, "graph instantiate failed: %s", cudaGetErrorString(e));
   amg->graphs.push_back(ge);
   HB_CUDA(cudaGraphLaunch(ge.exec, c.s_comp));
   c.launches += ge.launches;
   return 0;
}

// relaxation sweeps one cycle makes on every level: the cycling state machine of cycle_body without the work
// (the reference's "cycle complexity" adds the nonzeros of a level once per sweep, par_cycle.c:455-474)
static void amg_cycle_sweeps(const hb200_amg *amg, int *sweeps)
{
   const int nl = amg->num_levels;
   for (int l = 0; l < nl; l++) sweeps[l] = 0;
   std::vector<int> lev_counter(nl);
   lev_counter[0] = 1;
   for (int k = 1; k < nl; k++) lev_counter[k] = amg->fcycle ? 1 : amg->cycle_type;
   int fcycle_lev = nl - 2, level = 0, cycle_param = 1;
   bool not_finished = true;
   while (not_finished) {
      sweeps[level] += (nl > 1) ? amg->num_grid_sweeps[cycle_param] : amg->num_grid_sweeps[0];
      --lev_counter[level];
      if (lev_counter[level] >= 0 && level != nl - 1) {
         ++level;
         lev_counter[level] = lev_counter[level] > amg->cycle_type ? lev_counter[level] : amg->cycle_type;
         cycle_param = (level == nl - 1) ? 3 : 1;
      } else if (level != 0) {
         --level;
         cycle_param = 2;
         if (amg->fcycle && fcycle_lev == level) {
            lev_counter[level] = lev_counter[level] > 1 ? lev_counter[level] : 1;
            fcycle_lev--;
         }
      } else {
         not_finished = false;
      }
   }
}

// resid_norms != NULL: the residual norm before the first and after every cycle (max_iter + 1 values) and the
// norm of f are handed back for the caller's per-cycle table — computed even when tol == 0, as the reference does
// with print_level > 1 / logging > 1 (par_amg_solve.c:130, 228)
int amg_solve(hb200_amg *amg, hb200_parcsr *A, const double *f, double *u, bool u_all_zeros,
              int *num_iterations, double *rel_resid_norm, double *resid_norms, double *rhs_norm_out)
{
   // hypre_BoomerAMGSolve (par_amg_solve.c:22-424); the printing is the caller's
   Ctx &c = ctx();
   (void) A;
   hb200_amg_level &L0 = amg->lev[0];
   const size_t n = (size_t) L0.n;
   const double tol = amg->tol;
   const bool norms = tol > 0.0 || resid_norms != nullptr;
   double resid_nrm = 1.0, resid_nrm_init = 0.0, rhs_norm = 0.0, relative_resid = 1.0;
   const int S = kScalarSlots - 4;
   int flag = 0;
   if (norms) {
      // Vtemp = f ; Vtemp = A u - f ; ||Vtemp||   (:170-182)
      if (u_all_zeros) { HB_CHECK(vec_set(u, 0.0, n, c.s_comp)); u_all_zeros = false; }
      HB_CHECK(parcsr_matvec(L0.A, 1.0, u, -1.0, f, L0.Vtemp));
      HB_CHECK(vec_dot2_dev(L0.Vtemp, L0.Vtemp, f, f, n, S, S + 1, c.s_comp));
      HB_CHECK(scalars_allreduce(S, 2, c.s_comp));
      double v[2];
      HB_CHECK(scalars_fetch(S, 2, v, c.s_comp));
      resid_nrm = sqrt(v[0]);
      if (resid_nrm != 0.0 && !(resid_nrm / resid_nrm == resid_nrm / resid_nrm)) {
         return set_error(HB200_ERROR_GENERIC, "hb200_amg_solve: INFs and/or NaNs detected in input");
      }
      resid_nrm_init = resid_nrm;
      if (amg->converge_type == 0) {
         rhs_norm = sqrt(v[1]);
         relative_resid = rhs_norm != 0.0 ? resid_nrm_init / rhs_norm : resid_nrm_init;
      } else {
         relative_resid = 1.0;
      }
      if (resid_norms) resid_norms[0] = resid_nrm_init;
   }
   int cycle_count = 0;
   while ((relative_resid >= tol || cycle_count < amg->min_iter) && cycle_count < amg->max_iter) {
      HB_CHECK(amg_cycle(amg, f, u, u_all_zeros));
      u_all_zeros = false;
      if (norms) {
         HB_CHECK(parcsr_matvec(L0.A, 1.0, u, -1.0, f, L0.Vtemp));
         HB_CHECK(vec_dot_dev(L0.Vtemp, L0.Vtemp, n, S, c.s_comp));
         HB_CHECK(scalars_allreduce(S, 1, c.s_comp));
         double v;
         HB_CHECK(scalars_fetch(S, 1, &v, c.s_comp));
         resid_nrm = sqrt(v);
         if (amg->converge_type == 0) relative_resid = rhs_norm != 0.0 ? resid_nrm / rhs_norm : resid_nrm;
         else                          relative_resid = resid_nrm / resid_nrm_init;
         if (resid_norms) resid_norms[cycle_count + 1] = resid_nrm;
      }
      ++cycle_count;
   }
   if (rhs_norm_out) *rhs_norm_out = rhs_norm;
   if (cycle_count == amg->max_iter && tol > 0.0) flag |= HB200_ERROR_CONV;
   if (num_iterations) *num_iterations = cycle_count;
   if (rel_resid_norm) *rel_resid_norm = relative_resid;
   return flag;
}

void amg_drop_graphs(hb200_amg *amg)
{
   if (amg->graphs.empty()) return;
   cudaStreamSynchronize(ctx().s_comp);
   for (auto &g : amg->graphs) cudaGraphExecDestroy(g.exec);
   amg->graphs.clear();
}

void amg_set_dot_request(hb200_amg *amg, int slot) { if (amg) amg->dot_req_slot = slot; }
bool amg_dot_fused(const hb200_amg *amg) { return amg && amg->dot_fused; }

}  // namespace hb

using namespace hb;

// record of a saved hierarchy (hb200_amg_save / hb200_amg_load below)
static const unsigned long long kAmgMagic = 0x31474d4130304248ULL;   // "HB00AMG1"

template <class T>
static bool wr(FILE *f, const T *p, size_t n) { return n == 0 || fwrite(p, sizeof(T), n, f) == n; }
template <class T>
static bool rd(FILE *f, T *p, size_t n) { return n == 0 || fread(p, sizeof(T), n, f) == n; }

extern "C" {

int hb200_amg_create(hb200_amg **out, int num_levels)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(out && num_levels >= 1, HB200_ERROR_ARG, "bad arguments");
   hb200_amg *a = new hb200_amg();
   a->num_levels = num_levels;
   a->lev.resize(num_levels);
   *out = a;
   return 0;
}

int hb200_amg_destroy(hb200_amg *amg)
{
   if (!amg) return 0;
   cudaDeviceSynchronize();
   for (int l = 0; l < amg->num_levels; l++) {
      hb200_amg_level &L = amg->lev[l];
      if (L.l1) cudaFree(L.l1);
      if (L.cf) cudaFree(L.cf);
      if (l > 0) { if (L.F) cudaFree(L.F); if (L.U) cudaFree(L.U); }
      if (L.Ualt) cudaFree(L.Ualt);
      if (L.Vtemp) cudaFree(L.Vtemp);
      if (L.cheby_ds) cudaFree(L.cheby_ds);
      if (L.cheby_work) cudaFree(L.cheby_work);
   }
   if (amg->ge.d_LfT) cudaFree(amg->ge.d_LfT);
   if (amg->ge.d_UT) cudaFree(amg->ge.d_UT);
   if (amg->ge.d_Udiag) cudaFree(amg->ge.d_Udiag);
   if (amg->ge.d_b) cudaFree(amg->ge.d_b);
   for (auto &g : amg->graphs) cudaGraphExecDestroy(g.exec);
   if (amg->owns_matrices) {
      for (int l = 0; l < amg->num_levels; l++) {
         if (amg->lev[l].A) hb200_parcsr_destroy(amg->lev[l].A);
         if (amg->lev[l].P) hb200_parcsr_destroy(amg->lev[l].P);
      }
   }
   delete amg;
   return 0;
}

int hb200_amg_set_level(hb200_amg *amg, int level, hb200_parcsr *A, hb200_parcsr *P,
                        const double *l1_norms, const int *cf_marker, double relax_weight,
                        double omega)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(amg && A && level >= 0 && level < amg->num_levels, HB200_ERROR_ARG, "bad arguments");
   HB_REQUIRE(P != nullptr || level == amg->num_levels - 1, HB200_ERROR_ARG, "P missing on a non-coarsest level");
   hb200_amg_level &L = amg->lev[level];
   L.A = A; L.P = P; L.n = A->num_rows;
   L.relax_weight = relax_weight; L.omega = omega;
   const size_t n = (size_t) (L.n ? L.n : 1);
   if (l1_norms) {
      if (!L.l1) HB_CUDA(cudaMalloc(&L.l1, sizeof(double) * n));
      HB_CUDA(cudaMemcpy(L.l1, l1_norms, sizeof(double) * (size_t) L.n, cudaMemcpyHostToDevice));
   }
   if (cf_marker) {
      if (!L.cf) HB_CUDA(cudaMalloc(&L.cf, sizeof(int) * n));
      HB_CUDA(cudaMemcpy(L.cf, cf_marker, sizeof(int) * (size_t) L.n, cudaMemcpyHostToDevice));
   }
   HB_CHECK(level_alloc(L, level == 0));
   if (P) HB_CHECK(parcsr_ensure_T(P));
   return 0;
}

int hb200_amg_set_level_cheby(hb200_amg *amg, int level, const double *ds, const double *coefs, int order)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(amg && coefs && level >= 0 && level < amg->num_levels, HB200_ERROR_ARG, "bad arguments");
   hb200_amg_level &L = amg->lev[level];
   HB_REQUIRE(L.A != nullptr, HB200_ERROR_ARG, "set_level must precede set_level_cheby");
   L.cheby_coefs.assign(coefs, coefs + order + 1);
   L.cheby_order_set = order;
   if (ds) {
      if (!L.cheby_ds) HB_CUDA(cudaMalloc(&L.cheby_ds, sizeof(double) * (size_t) (L.n ? L.n : 1)));
      HB_CUDA(cudaMemcpy(L.cheby_ds, ds, sizeof(double) * (size_t) L.n, cudaMemcpyHostToDevice));
   }
   return 0;
}

int hb200_amg_set_cycle(hb200_amg *amg, const int *num_grid_sweeps4, const int *grid_relax_type4,
                        int relax_order, int cycle_type, int fcycle, int cheby_order,
                        int cheby_scale, int cheby_variant, int user_relax_type)
{
   HB_REQUIRE(amg && num_grid_sweeps4 && grid_relax_type4, HB200_ERROR_ARG, "bad arguments");
   bool changed = amg->relax_order != relax_order || amg->cycle_type != cycle_type || amg->fcycle != fcycle ||
                  amg->cheby_order != cheby_order || amg->cheby_scale != cheby_scale ||
                  amg->cheby_variant != cheby_variant || amg->user_relax_type != user_relax_type;
   for (int k = 0; k < 4; k++) {
      changed = changed || amg->num_grid_sweeps[k] != num_grid_sweeps4[k] || amg->grid_relax_type[k] != grid_relax_type4[k];
      amg->num_grid_sweeps[k] = num_grid_sweeps4[k];
      amg->grid_relax_type[k] = grid_relax_type4[k];
   }
   amg->relax_order = relax_order; amg->cycle_type = cycle_type; amg->fcycle = fcycle;
   amg->cheby_order = cheby_order; amg->cheby_scale = cheby_scale; amg->cheby_variant = cheby_variant;
   amg->user_relax_type = user_relax_type;
   if (changed) amg_drop_graphs(amg);   // the captured cycles have the old topology
   return 0;
}

int hb200_amg_set_level_weights(hb200_amg *amg, int level, double relax_weight, double omega)
{
   HB_REQUIRE(amg && level >= 0 && level < amg->num_levels, HB200_ERROR_ARG, "bad arguments");
   hb200_amg_level &L = amg->lev[level];
   if (L.relax_weight != relax_weight || L.omega != omega) {
      L.relax_weight = relax_weight; L.omega = omega;
      amg_drop_graphs(amg);              // kernel arguments are captured by value
   }
   return 0;
}

int hb200_amg_set_solve(hb200_amg *amg, double tol, int min_iter, int max_iter, int converge_type)
{
   HB_REQUIRE(amg != nullptr, HB200_ERROR_ARG, "null amg");
   amg->tol = tol; amg->min_iter = min_iter; amg->max_iter = max_iter; amg->converge_type = converge_type;
   return 0;
}

int hb200_amg_set_coarse_ge(hb200_amg *amg, const double *A_mat, int n, int first_row, int num_local)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(amg && A_mat && n >= 1 && first_row >= 0 && num_local >= 0 && first_row + num_local <= n,
              HB200_ERROR_ARG, "bad arguments");
   std::vector<double> LfT, UT, Ud;
   HB_CHECK(ge_factor_host(A_mat, n, LfT, UT, Ud));
   GEData &g = amg->ge;
   g.n = n; g.first_row = first_row; g.num_local = num_local;
   const size_t nn = (size_t) n * n;
   amg->ge_A_mat.assign(A_mat, A_mat + nn);
   HB_CUDA(cudaMalloc(&g.d_LfT, sizeof(double) * nn));
   HB_CUDA(cudaMalloc(&g.d_UT, sizeof(double) * nn));
   HB_CUDA(cudaMalloc(&g.d_Udiag, sizeof(double) * n));
   HB_CUDA(cudaMalloc(&g.d_b, sizeof(double) * n));
   HB_CUDA(cudaMemcpy(g.d_LfT, LfT.data(), sizeof(double) * nn, cudaMemcpyHostToDevice));
   HB_CUDA(cudaMemcpy(g.d_UT, UT.data(), sizeof(double) * nn, cudaMemcpyHostToDevice));
   HB_CUDA(cudaMemcpy(g.d_Udiag, Ud.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
   amg->has_ge = true;
   return 0;
}

int hb200_amg_set_use_graph(hb200_amg *amg, int enable)
{
   HB_REQUIRE(amg != nullptr, HB200_ERROR_ARG, "null amg");
   amg->use_graph = enable != 0;
   return 0;
}

int hb200_amg_cycle(hb200_amg *amg, const double *f, double *u, int u_all_zeros)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(amg && ((f && u) || amg->lev[0].n == 0), HB200_ERROR_ARG, "null argument");
   for (int l = 0; l < amg->num_levels; l++) HB_REQUIRE(amg->lev[l].A, HB200_ERROR_ARG, "level not set");
   return amg_cycle(amg, f, u, u_all_zeros != 0);
}

int hb200_amg_solve(hb200_amg *amg, const double *f, double *u, int u_all_zeros, int *num_iterations,
                    double *rel_resid_norm)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(amg && ((f && u) || amg->lev[0].n == 0), HB200_ERROR_ARG, "null argument");
   for (int l = 0; l < amg->num_levels; l++) HB_REQUIRE(amg->lev[l].A, HB200_ERROR_ARG, "level not set");
   return amg_solve(amg, amg->lev[0].A, f, u, u_all_zeros != 0, num_iterations, rel_resid_norm, nullptr, nullptr);
}

int hb200_amg_solve_logged(hb200_amg *amg, const double *f, double *u, int u_all_zeros, int *num_iterations,
                           double *rel_resid_norm, double *resid_norms, double *rhs_norm)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(amg && resid_norms && ((f && u) || amg->lev[0].n == 0), HB200_ERROR_ARG, "null argument");
   for (int l = 0; l < amg->num_levels; l++) HB_REQUIRE(amg->lev[l].A, HB200_ERROR_ARG, "level not set");
   return amg_solve(amg, amg->lev[0].A, f, u, u_all_zeros != 0, num_iterations, rel_resid_norm, resid_norms, rhs_norm);
}

int hb200_amg_cycle_sweeps(const hb200_amg *amg, int *sweeps_per_level)
{
   HB_REQUIRE(amg && sweeps_per_level, HB200_ERROR_ARG, "null argument");
   amg_cycle_sweeps(amg, sweeps_per_level);
   return 0;
}

// ---- a hierarchy on disk: the level matrices as binary IJ files, everything else in one record per rank ----------
int hb200_amg_save(const hb200_amg *amg, const char *dirname)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(amg && dirname, HB200_ERROR_ARG, "hb200_amg_save: null argument");
   Ctx &c = ctx();
   for (int l = 0; l < amg->num_levels; l++) HB_REQUIRE(amg->lev[l].A, HB200_ERROR_ARG, "hb200_amg_save: level not set");
   if (mkdir(dirname, 0777) != 0 && errno != EEXIST) return set_error(HB200_ERROR_GENERIC, "hb200_amg_save: cannot create %s", dirname);
   char path[1024];
   // level matrices (collective calls: every rank goes through the same sequence)
   for (int l = 0; l < amg->num_levels; l++) {
      snprintf(path, sizeof(path), "%s/A%d", dirname, l);
      HB_CHECK(hb200_parcsr_print_ij_binary(amg->lev[l].A, path));
      if (amg->lev[l].P) {
         snprintf(path, sizeof(path), "%s/P%d", dirname, l);
         HB_CHECK(hb200_parcsr_print_ij_binary(amg->lev[l].P, path));
      }
   }
   snprintf(path, sizeof(path), "%s/amg.%05d.bin", dirname, c.rank);
   FILE *f = fopen(path, "wb");
   if (!f) return set_error(HB200_ERROR_GENERIC, "hb200_amg_save: cannot open %s", path);
   HB_CUDA(cudaStreamSynchronize(c.s_comp));
   const int ints[20] = {amg->num_levels, c.nranks, amg->num_grid_sweeps[0], amg->num_grid_sweeps[1], amg->num_grid_sweeps[2],
                         amg->num_grid_sweeps[3], amg->grid_relax_type[0], amg->grid_relax_type[1], amg->grid_relax_type[2],
                         amg->grid_relax_type[3], amg->relax_order, amg->cycle_type, amg->fcycle, amg->cheby_order, amg->cheby_scale,
                         amg->cheby_variant, amg->user_relax_type, amg->min_iter, amg->max_iter, amg->converge_type};
   bool ok = wr(f, &kAmgMagic, 1) && wr(f, ints, 20) && wr(f, &amg->tol, 1);
   for (int l = 0; ok && l < amg->num_levels; l++) {
      const hb200_amg_level &L = amg->lev[l];
      const size_t n = (size_t) L.n;
      const int head[6] = {L.n, L.P ? 1 : 0, L.l1 ? 1 : 0, L.cf ? 1 : 0, L.cheby_ds ? 1 : 0, (int) L.cheby_coefs.size()};
      const int chunks = L.A->gs_chunks;
      ok = wr(f, head, 6) && wr(f, &L.relax_weight, 1) && wr(f, &L.omega, 1) && wr(f, &L.cheby_order_set, 1) && wr(f, &chunks, 1);
      std::vector<double> hv(n);
      std::vector<int> hi(n);
      if (ok && L.l1) { if (n) HB_CUDA(cudaMemcpy(hv.data(), L.l1, sizeof(double) * n, cudaMemcpyDeviceToHost)); ok = wr(f, hv.data(), n); }
      if (ok && L.cf) { if (n) HB_CUDA(cudaMemcpy(hi.data(), L.cf, sizeof(int) * n, cudaMemcpyDeviceToHost)); ok = wr(f, hi.data(), n); }
      if (ok && L.cheby_ds) { if (n) HB_CUDA(cudaMemcpy(hv.data(), L.cheby_ds, sizeof(double) * n, cudaMemcpyDeviceToHost)); ok = wr(f, hv.data(), n); }
      ok = ok && wr(f, L.cheby_coefs.data(), L.cheby_coefs.size());
   }
   const int ge[4] = {amg->has_ge ? 1 : 0, amg->ge.n, amg->ge.first_row, amg->ge.num_local};
   ok = ok && wr(f, ge, 4);
   if (ok && amg->has_ge) ok = wr(f, amg->ge_A_mat.data(), amg->ge_A_mat.size());
   fclose(f);
   if (!ok) return set_error(HB200_ERROR_GENERIC, "hb200_amg_save: write to %s failed", path);
   return 0;
}

int hb200_amg_load(hb200_amg **out, const char *dirname)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(out && dirname, HB200_ERROR_ARG, "hb200_amg_load: null argument");
   Ctx &c = ctx();
   char path[1024];
   snprintf(path, sizeof(path), "%s/amg.%05d.bin", dirname, c.rank);
   FILE *f = fopen(path, "rb");
   if (!f) return set_error(HB200_ERROR_ARG, "hb200_amg_load: cannot open %s", path);
   unsigned long long magic = 0;
   int ints[20];
   double tol = 0.0;
   bool ok = rd(f, &magic, 1) && magic == kAmgMagic && rd(f, ints, 20) && rd(f, &tol, 1);
   if (!ok || ints[0] < 1 || ints[0] > 100) { fclose(f); return set_error(HB200_ERROR_GENERIC, "%s is not a hierarchy record of this library", path); }
   if (ints[1] != c.nranks) { fclose(f); return set_error(HB200_ERROR_ARG, "hb200_amg_load: the hierarchy was saved on %d ranks, this run has %d", ints[1], c.nranks); }
   hb200_amg *amg = nullptr;
   int fl = hb200_amg_create(&amg, ints[0]);
   if (fl) { fclose(f); return fl; }
   amg->owns_matrices = true;
   auto fail = [&](int flag) { fclose(f); hb200_amg_destroy(amg); return flag; };
   for (int l = 0; l < amg->num_levels; l++) {
      int head[6], cheby_order_set = 0, chunks = 0;
      double rw = 1.0, om = 1.0;
      if (!(rd(f, head, 6) && rd(f, &rw, 1) && rd(f, &om, 1) && rd(f, &cheby_order_set, 1) && rd(f, &chunks, 1)) || head[0] < 0) {
         return fail(set_error(HB200_ERROR_GENERIC, "hb200_amg_load: %s is truncated", path));
      }
      const size_t n = (size_t) head[0];
      std::vector<double> l1(n), ds(n), coefs((size_t) (head[5] > 0 ? head[5] : 0));
      std::vector<int> cf(n);
      ok = (!head[2] || rd(f, l1.data(), n)) && (!head[3] || rd(f, cf.data(), n)) && (!head[4] || rd(f, ds.data(), n)) &&
           rd(f, coefs.data(), coefs.size());
      if (!ok) return fail(set_error(HB200_ERROR_GENERIC, "hb200_amg_load: %s is truncated", path));
      hb200_parcsr *A = nullptr, *P = nullptr;
      char mp[1024];
      snprintf(mp, sizeof(mp), "%s/A%d", dirname, l);
      fl = parcsr_read_binary_exact(&A, mp);
      if (!fl && head[1]) { snprintf(mp, sizeof(mp), "%s/P%d", dirname, l); fl = parcsr_read_binary_exact(&P, mp); }
      // the matrices belong to the hierarchy from here on (destroyed with it, also on the failure paths below)
      amg->lev[l].A = A; amg->lev[l].P = P;
      if (fl) return fail(fl);
      if (A->num_rows != head[0]) return fail(set_error(HB200_ERROR_GENERIC, "hb200_amg_load: level %d has %d rows, the record says %d", l, A->num_rows, head[0]));
      fl = hb200_amg_set_level(amg, l, A, P, head[2] ? l1.data() : nullptr, head[3] ? cf.data() : nullptr, rw, om);
      if (!fl && !coefs.empty()) fl = hb200_amg_set_level_cheby(amg, l, head[4] ? ds.data() : nullptr, coefs.data(), (int) coefs.size() - 1);
      if (!fl && chunks > 1) fl = hb200_parcsr_set_gs_chunks(A, chunks);
      if (fl) return fail(fl);
      amg->lev[l].cheby_order_set = cheby_order_set;
   }
   fl = hb200_amg_set_cycle(amg, ints + 2, ints + 6, ints[10], ints[11], ints[12], ints[13], ints[14], ints[15], ints[16]);
   if (!fl) fl = hb200_amg_set_solve(amg, tol, ints[17], ints[18], ints[19]);
   int ge[4];
   if (!fl && !rd(f, ge, 4)) fl = set_error(HB200_ERROR_GENERIC, "hb200_amg_load: %s is truncated", path);
   if (!fl && ge[0]) {
      std::vector<double> A_mat((size_t) ge[1] * (size_t) ge[1]);
      if (!rd(f, A_mat.data(), A_mat.size())) fl = set_error(HB200_ERROR_GENERIC, "hb200_amg_load: %s is truncated", path);
      else fl = hb200_amg_set_coarse_ge(amg, A_mat.data(), ge[1], ge[2], ge[3]);
   }
   if (fl) return fail(fl);
   fclose(f);
   *out = amg;
   return 0;
}

int hb200_amg_level_matrix(hb200_amg *amg, int level, int which, hb200_parcsr **M)
{
   HB_REQUIRE(amg && M && level >= 0 && level < amg->num_levels && (which == 0 || which == 1), HB200_ERROR_ARG, "bad arguments");
   *M = which == 0 ? amg->lev[level].A : amg->lev[level].P;
   return 0;
}

int hb200_amg_num_levels(const hb200_amg *amg)
{
   return amg ? amg->num_levels : 0;
}

int hb200_amg_level_vector(hb200_amg *amg, int level, int which, double **dev, int *n)
{
   HB_REQUIRE(amg && dev && level >= 0 && level < amg->num_levels, HB200_ERROR_ARG, "bad arguments");
   hb200_amg_level &L = amg->lev[level];
   *dev = which == 0 ? L.F : L.u_cur;
   if (n) *n = L.n;
   return 0;
}

}  // extern "C"
