// ij.cu — the callers' entry into ParCSR without hypre in the process: coordinate
// triplets -> ParCSR (diag / offd split, col_map_offd, CommPkg), and hypre's on-disk IJ formats.
//
// Reference: HYPRE_IJMatrixCreate / SetValues / AddToValues / Assemble (src/IJ_mv/HYPRE_IJMatrix.c),
// hypre_IJMatrixAssembleParCSR (src/IJ_mv/IJMatrix_parcsr.c:2641-3115: the order of a row's entries, the
// diagonal moved to the front, col_map_offd), hypre_MatvecCommPkgCreate (src/parcsr_mv/par_csr_communication.c:
// neighbours ascending, the entries a neighbour needs in ascending global order), hypre_IJMatrixRead
// (src/IJ_mv/IJMatrix.c:110-249: "<name>.<5-digit rank>", header `ilower iupper jlower jupper`, then `i j value`;
// Matrix Market with is_mm), hypre_ParCSRMatrixPrintIJ (src/parcsr_mv/par_csr_matrix.c: `%b %b %.14e`),
// hypre_ParVectorPrintIJ / hypre_IJVectorRead (`jlower jupper`, then `j value`).
//
// The assembly is host code (one pass per rank over its triplets; the arrays then go through the same
// upload as a hierarchy level, hb200_parcsr_create).  On N ranks the CommPkg needs what the other ranks ask
// for: two NCCL all-gathers (the ownership ranges, then every rank's col_map_offd padded to the longest).
#include "hb_internal.cuh"
#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

namespace hb {

struct IjHost {
   int64_t ilower = 0, iupper = -1, jlower = 0, jupper = -1;
   std::vector<int> diag_i, diag_j, offd_i, offd_j;
   std::vector<double> diag_a, offd_a;
   std::vector<int64_t> col_map_offd;
};

// rows of [ilower, iupper]: entries in insertion order, a repeated (i, j) lands on its first occurrence
// (add: summed, else the last value wins — HYPRE_IJMatrixAddToValues / SetValues), the diagonal entry of a
// square block first (IJMatrix_parcsr.c:2830-2848), off-range columns compressed through the ascending
// col_map_offd (:2954-3001)
static int ij_assemble_host(int64_t ilower, int64_t iupper, int64_t jlower, int64_t jupper, int64_t nnz,
                            const int64_t *rows, const int64_t *cols, const double *vals, int add, IjHost &out,
                            bool move_diag = true, int64_t add_from = -1)
{
   // entries k >= add_from (contributions received from other ranks) are always added, whatever `add` says
   if (add_from < 0) add_from = nnz;
   const int64_t nr64 = iupper - ilower + 1, nc64 = jupper - jlower + 1;
   HB_REQUIRE(nr64 >= 0 && nc64 >= 0 && nr64 < 0x7fffffffLL && nc64 < 0x7fffffffLL, HB200_ERROR_ARG, "IJ ranges out of the 32-bit local index space");
   HB_REQUIRE(nnz >= 0 && nnz < 0x7fffffffLL, HB200_ERROR_ARG, "IJ triplet count out of range");
   const int nr = (int) nr64;
   out = IjHost();
   out.ilower = ilower; out.iupper = iupper; out.jlower = jlower; out.jupper = jupper;
   // counting sort of the triplets by row, stable (insertion order inside a row)
   std::vector<int> start((size_t) nr + 1, 0);
   for (int64_t k = 0; k < nnz; k++) {
      HB_REQUIRE(rows[k] >= ilower && rows[k] <= iupper, HB200_ERROR_ARG, "IJ triplet outside the rows this rank owns");
      start[(size_t) (rows[k] - ilower) + 1]++;
   }
   for (int r = 0; r < nr; r++) start[(size_t) r + 1] += start[(size_t) r];
   std::vector<int> order((size_t) nnz), fill(start.begin(), start.end() - 1);
   for (int64_t k = 0; k < nnz; k++) order[(size_t) fill[(size_t) (rows[k] - ilower)]++] = (int) k;
   out.diag_i.assign((size_t) nr + 1, 0);
   out.offd_i.assign((size_t) nr + 1, 0);
   std::vector<int64_t> big_offd_j;
   std::vector<int64_t> rj;
   std::vector<double> ra;
   std::unordered_map<int64_t, int> seen;
   for (int r = 0; r < nr; r++) {
      rj.clear(); ra.clear(); seen.clear();
      for (int p = start[(size_t) r]; p < start[(size_t) r + 1]; p++) {
         const int k = order[(size_t) p];
         const int64_t j = cols[k];
         auto it = seen.find(j);
         if (it == seen.end()) { seen.emplace(j, (int) rj.size()); rj.push_back(j); ra.push_back(vals[k]); }
         else if (add || (int64_t) k >= add_from) ra[(size_t) it->second] += vals[k];
         else ra[(size_t) it->second] = vals[k];
      }
      int dpos = -1;
      for (size_t q = 0; move_diag && q < rj.size(); q++) {
         if (rj[q] >= jlower && rj[q] <= jupper && rj[q] - jlower == (int64_t) r) { dpos = (int) q; break; }
      }
      if (dpos >= 0) { out.diag_j.push_back(r); out.diag_a.push_back(ra[(size_t) dpos]); }
      for (size_t q = 0; q < rj.size(); q++) {
         if (rj[q] < jlower || rj[q] > jupper) { big_offd_j.push_back(rj[q]); out.offd_a.push_back(ra[q]); }
         else if ((int) q != dpos) { out.diag_j.push_back((int) (rj[q] - jlower)); out.diag_a.push_back(ra[q]); }
      }
      out.diag_i[(size_t) r + 1] = (int) out.diag_j.size();
      out.offd_i[(size_t) r + 1] = (int) big_offd_j.size();
   }
   out.col_map_offd = big_offd_j;
   std::sort(out.col_map_offd.begin(), out.col_map_offd.end());
   out.col_map_offd.erase(std::unique(out.col_map_offd.begin(), out.col_map_offd.end()), out.col_map_offd.end());
   out.offd_j.resize(big_offd_j.size());
   for (size_t q = 0; q < big_offd_j.size(); q++) {
      out.offd_j[q] = (int) (std::lower_bound(out.col_map_offd.begin(), out.col_map_offd.end(), big_offd_j[q]) - out.col_map_offd.begin());
   }
   return 0;
}

// all-gather of `count` int64 per rank through device buffers (NCCL moves device memory)
static int allgather_i64(const int64_t *mine, size_t count, std::vector<int64_t> &all)
{
   Ctx &c = ctx();
   all.assign(count * (size_t) c.nranks, 0);
   if (c.nranks <= 1) { std::copy(mine, mine + count, all.begin()); return 0; }
#ifdef HB200_WITH_NCCL
   if (count == 0) return 0;
   int64_t *d_in = nullptr, *d_out = nullptr;
   HB_CUDA(cudaMalloc(&d_in, sizeof(int64_t) * count));
   HB_CUDA(cudaMalloc(&d_out, sizeof(int64_t) * count * (size_t) c.nranks));
   HB_CUDA(cudaMemcpyAsync(d_in, mine, sizeof(int64_t) * count, cudaMemcpyHostToDevice, c.s_comp));
   ncclResult_t r = nccl_api().AllGather(d_in, d_out, count, ncclInt64, c.nccl, c.s_comp);
   cudaError_t e = cudaMemcpyAsync(all.data(), d_out, sizeof(int64_t) * all.size(), cudaMemcpyDeviceToHost, c.s_comp);
   if (e == cudaSuccess) e = cudaStreamSynchronize(c.s_comp);
   cudaFree(d_in); cudaFree(d_out);
   if (r != ncclSuccess) return set_error(HB200_ERROR_GENERIC, "IJ assembly: ncclAllGather failed: %s", nccl_api().GetErrorString(r));
   if (e != cudaSuccess) return set_error(HB200_ERROR_GENERIC, "IJ assembly: %s", cudaGetErrorString(e));
   return 0;
#else
   return set_error(HB200_ERROR_GENERIC, "libhb200 built without NCCL but nranks > 1");
#endif
}

// The CommPkg of rank `me` from what every rank owns (own[5 q .. 5 q + 4] = ilower, iupper, jlower, jupper, number
// of off-range columns) and needs (need[q * max_offd ..]: rank q's col_map_offd, ascending): pure host code.
//  * receive side: the owners of my off-range columns.  The column ranges are disjoint and my list ascends, so one
//    owner's columns are contiguous in it;
//  * send side: what every other rank needs of my columns, ranks ascending, in its order (ascending global index).
static int ij_commpkg_host(int np, int me, const int64_t *own, const int64_t *need, int64_t max_offd,
                           std::vector<int> &send_procs, std::vector<int> &send_map_starts, std::vector<int> &send_map_elmts,
                           std::vector<int> &recv_procs, std::vector<int> &recv_vec_starts)
{
   send_procs.clear(); send_map_elmts.clear(); recv_procs.clear();
   send_map_starts.assign(1, 0); recv_vec_starts.assign(1, 0);
   const int64_t jlower = own[(size_t) me * 5 + 2], jupper = own[(size_t) me * 5 + 3];
   const int n_offd = (int) own[(size_t) me * 5 + 4];
   const int64_t *mine = need + (size_t) me * (size_t) max_offd;
   for (int k = 0; k < n_offd; ) {
      const int64_t g = mine[k];
      int owner = -1;
      for (int q = 0; q < np; q++) if (q != me && g >= own[(size_t) q * 5 + 2] && g <= own[(size_t) q * 5 + 3]) { owner = q; break; }
      HB_REQUIRE(owner >= 0, HB200_ERROR_ARG, "IJ assembly: a column index is owned by no rank");
      int e = k;
      while (e < n_offd && mine[e] >= own[(size_t) owner * 5 + 2] && mine[e] <= own[(size_t) owner * 5 + 3]) e++;
      recv_procs.push_back(owner);
      recv_vec_starts.push_back(e);
      k = e;
   }
   for (int q = 0; q < np; q++) {
      if (q == me) continue;
      const int64_t nq = own[(size_t) q * 5 + 4];
      const int64_t *lst = need + (size_t) q * (size_t) max_offd;
      const size_t before = send_map_elmts.size();
      for (int64_t t = 0; t < nq; t++) if (lst[t] >= jlower && lst[t] <= jupper) send_map_elmts.push_back((int) (lst[t] - jlower));
      if (send_map_elmts.size() > before) { send_procs.push_back(q); send_map_starts.push_back((int) send_map_elmts.size()); }
   }
   return 0;
}

// the ParCSR object of this rank's assembled rows: global sizes and the CommPkg from what every rank owns and needs
static int ij_create_parcsr(const IjHost &h, hb200_parcsr **A)
{
   Ctx &c = ctx();
   const int np = c.nranks, me = c.rank;
   const int nr = (int) (h.iupper - h.ilower + 1), nc = (int) (h.jupper - h.jlower + 1);
   const int n_offd = (int) h.col_map_offd.size();
   int64_t mine[5] = {h.ilower, h.iupper, h.jlower, h.jupper, (int64_t) n_offd};
   std::vector<int64_t> own;
   HB_CHECK(allgather_i64(mine, 5, own));
   int64_t gr0 = own[0], gr1 = own[1], gc0 = own[2], gc1 = own[3], max_offd = 0;
   for (int q = 0; q < np; q++) {
      gr0 = std::min(gr0, own[(size_t) q * 5 + 0]); gr1 = std::max(gr1, own[(size_t) q * 5 + 1]);
      gc0 = std::min(gc0, own[(size_t) q * 5 + 2]); gc1 = std::max(gc1, own[(size_t) q * 5 + 3]);
      max_offd = std::max(max_offd, own[(size_t) q * 5 + 4]);
   }
   std::vector<int> send_procs, send_map_starts(1, 0), send_map_elmts, recv_procs, recv_vec_starts(1, 0);
   if (np > 1) {
      // every rank's col_map_offd (ascending), padded to the longest
      std::vector<int64_t> pad((size_t) max_offd, -1), need;
      std::copy(h.col_map_offd.begin(), h.col_map_offd.end(), pad.begin());
      HB_CHECK(allgather_i64(pad.data(), (size_t) max_offd, need));
      HB_CHECK(ij_commpkg_host(np, me, own.data(), need.data(), max_offd, send_procs, send_map_starts, send_map_elmts, recv_procs, recv_vec_starts));
   }
   std::vector<int64_t> cmap(h.col_map_offd);
   for (auto &g : cmap) g -= gc0;                        // (IJMatrix_parcsr.c:3003-3009: global indices without the base)
   const int zero = 0;
   return hb200_parcsr_create(A, nr, nc, n_offd, h.diag_i.data(), h.diag_j.empty() ? &zero : h.diag_j.data(),
                              h.diag_a.data(), n_offd ? h.offd_i.data() : nullptr, h.offd_j.data(), h.offd_a.data(),
                              n_offd ? cmap.data() : nullptr, h.ilower - gr0, h.jlower - gc0, gr1 - gr0 + 1, gc1 - gc0 + 1,
                              (int) send_procs.size(), send_procs.data(), send_map_starts.data(), send_map_elmts.data(),
                              (int) recv_procs.size(), recv_procs.data(), recv_vec_starts.data());
}

// one rank's part file of hypre_IJMatrixRead, or a whole Matrix Market file (coordinate, real / integer,
// general / symmetric; 1 rank)
static int ij_parse_file(const char *filename, int is_mm, int rank, int64_t *range4, std::vector<int64_t> &rows,
                         std::vector<int64_t> &cols, std::vector<double> &vals)
{
   char path[1024];
   if (is_mm) snprintf(path, sizeof(path), "%s", filename);
   else snprintf(path, sizeof(path), "%s.%05d", filename, rank);
   FILE *f = fopen(path, "r");
   if (!f) return set_error(HB200_ERROR_ARG, "cannot open %s", path);
   bool sym = false;
   char line[512];
   if (is_mm) {
      if (!fgets(line, sizeof(line), f) || strncmp(line, "%%MatrixMarket", 14) != 0) { fclose(f); return set_error(HB200_ERROR_GENERIC, "%s: no Matrix Market banner", path); }
      std::string b(line);
      for (auto &ch : b) ch = (char) tolower(ch);
      if (b.find("coordinate") == std::string::npos || (b.find("real") == std::string::npos && b.find("integer") == std::string::npos)) {
         fclose(f);
         return set_error(HB200_ERROR_GENERIC, "%s: only sparse real-valued/integer coordinate matrices are supported", path);
      }
      sym = b.find("symmetric") != std::string::npos;
      long long nrow = 0, ncol = 0, nz = 0;
      while (fgets(line, sizeof(line), f)) { if (line[0] != '%') break; }
      if (sscanf(line, "%lld %lld %lld", &nrow, &ncol, &nz) != 3) { fclose(f); return set_error(HB200_ERROR_GENERIC, "%s: MM read size error", path); }
      range4[0] = 0; range4[1] = nrow - 1; range4[2] = 0; range4[3] = ncol - 1;
   } else {
      long long a, b, c2, d;
      if (fscanf(f, "%lld %lld %lld %lld", &a, &b, &c2, &d) != 4) { fclose(f); return set_error(HB200_ERROR_GENERIC, "%s: no IJ header", path); }
      range4[0] = a; range4[1] = b; range4[2] = c2; range4[3] = d;
   }
   long long I, J;
   double v;
   int ret;
   while ((ret = fscanf(f, "%lld %lld %le", &I, &J, &v)) != EOF) {
      if (ret != 3) { fclose(f); return set_error(HB200_ERROR_GENERIC, "Error in IJ matrix input file %s", path); }
      if (is_mm) { I--; J--; }
      rows.push_back(I); cols.push_back(J); vals.push_back(v);
      if (sym && I != J) { rows.push_back(J); cols.push_back(I); vals.push_back(v); }
   }
   fclose(f);
   return 0;
}

// hypre_IJMatrixReadBinary (src/IJ_mv/IJMatrix.c:252-470): `<name>.<5-digit rank>.bin` = 11 x uint64 header (version 1,
// bytes per index 4 | 8, bytes per value 4 | 8, global rows, global columns, global nonzeros, local nonzeros, ilower,
// iupper, jlower, jupper), then the row indices, the column indices, the values of the local entries
static int ij_parse_binary(const char *filename, int rank, int64_t *range4, std::vector<int64_t> &rows,
                           std::vector<int64_t> &cols, std::vector<double> &vals)
{
   char path[1024];
   snprintf(path, sizeof(path), "%s.%05d.bin", filename, rank);
   FILE *f = fopen(path, "rb");
   if (!f) return set_error(HB200_ERROR_ARG, "Could not open input file %s", path);
   uint64_t header[11];
   if (fread(header, sizeof(uint64_t), 11, f) != 11) { fclose(f); return set_error(HB200_ERROR_GENERIC, "%s: Could not read header entries", path); }
   if (header[0] != 1) { fclose(f); return set_error(HB200_ERROR_GENERIC, "%s: Unsupported header version: %llu", path, (unsigned long long) header[0]); }
   if (header[6] > 0x7fffffffULL) { fclose(f); return set_error(HB200_ERROR_GENERIC, "%s: Detected integer overflow at 7th header entry", path); }
   if ((header[1] != 4 && header[1] != 8) || (header[2] != 4 && header[2] != 8)) { fclose(f); return set_error(HB200_ERROR_GENERIC, "%s: Unsupported data type", path); }
   const size_t nnz = (size_t) header[6];
   range4[0] = (int64_t) header[7]; range4[1] = (int64_t) header[8]; range4[2] = (int64_t) header[9]; range4[3] = (int64_t) header[10];
   rows.resize(nnz); cols.resize(nnz); vals.resize(nnz);
   auto read_idx = [&](std::vector<int64_t> &dst) -> bool {
      if (header[1] == 8) {
         std::vector<uint64_t> b(nnz);
         if (fread(b.data(), 8, nnz, f) != nnz) return false;
         for (size_t k = 0; k < nnz; k++) dst[k] = (int64_t) b[k];
      } else {
         std::vector<uint32_t> b(nnz);
         if (fread(b.data(), 4, nnz, f) != nnz) return false;
         for (size_t k = 0; k < nnz; k++) dst[k] = (int64_t) b[k];
      }
      return true;
   };
   bool ok = read_idx(rows) && read_idx(cols);
   if (ok && header[2] == 8) ok = fread(vals.data(), 8, nnz, f) == nnz;
   else if (ok) {
      std::vector<float> b(nnz);
      ok = fread(b.data(), 4, nnz, f) == nnz;
      for (size_t k = 0; ok && k < nnz; k++) vals[k] = (double) b[k];
   }
   fclose(f);
   if (!ok) return set_error(HB200_ERROR_GENERIC, "%s: Could not read all entries", path);
   return 0;
}

// this rank's entries in the order the reference prints them: diag entries, then offd entries of a row
static int ij_download_coo(const hb200_parcsr *A, std::vector<int64_t> &rows, std::vector<int64_t> &cols, std::vector<double> &vals)
{
   Ctx &c = ctx();
   const int n = A->num_rows;
   std::vector<int> di((size_t) n + 1, 0), dj((size_t) A->diag.nnz), oi((size_t) n + 1, 0), oj((size_t) A->offd.nnz);
   std::vector<double> da((size_t) A->diag.nnz), oa((size_t) A->offd.nnz);
   HB_CUDA(cudaStreamSynchronize(c.s_comp));
   if (n) HB_CUDA(cudaMemcpy(di.data(), A->diag.i, sizeof(int) * di.size(), cudaMemcpyDeviceToHost));
   if (A->diag.nnz) {
      HB_CUDA(cudaMemcpy(dj.data(), A->diag.j, sizeof(int) * dj.size(), cudaMemcpyDeviceToHost));
      HB_CUDA(cudaMemcpy(da.data(), A->diag.a, sizeof(double) * da.size(), cudaMemcpyDeviceToHost));
   }
   if (A->num_cols_offd > 0 && A->offd.nnz) {
      HB_CUDA(cudaMemcpy(oi.data(), A->offd.i, sizeof(int) * oi.size(), cudaMemcpyDeviceToHost));
      HB_CUDA(cudaMemcpy(oj.data(), A->offd.j, sizeof(int) * oj.size(), cudaMemcpyDeviceToHost));
      HB_CUDA(cudaMemcpy(oa.data(), A->offd.a, sizeof(double) * oa.size(), cudaMemcpyDeviceToHost));
   }
   rows.clear(); cols.clear(); vals.clear();
   for (int i = 0; i < n; i++) {
      for (int p = di[(size_t) i]; p < di[(size_t) i + 1]; p++) { rows.push_back(A->first_row + i); cols.push_back(A->first_col + dj[(size_t) p]); vals.push_back(da[(size_t) p]); }
      for (int p = oi[(size_t) i]; p < oi[(size_t) i + 1]; p++) { rows.push_back(A->first_row + i); cols.push_back(A->col_map_offd[(size_t) oj[(size_t) p]]); vals.push_back(oa[(size_t) p]); }
   }
   return 0;
}

// a matrix this library wrote itself (hb200_parcsr_print_ij_binary: diag entries, then offd entries of every row) back
// into the same arrays: the order of the entries is kept as it is in the file (a level matrix of a saved hierarchy,
// amg.cu; interpolation matrices are rectangular and have no diagonal to move)
int parcsr_read_binary_exact(hb200_parcsr **A, const char *prefix)
{
   int64_t range[4];
   std::vector<int64_t> rows, cols;
   std::vector<double> vals;
   const int fr = ij_parse_binary(prefix, ctx().rank, range, rows, cols, vals);
   if (fr && ctx().nranks == 1) return fr;
   if (fr) { range[0] = 0; range[1] = -1; range[2] = 0; range[3] = -1; rows.clear(); cols.clear(); vals.clear(); }
   IjHost h;
   const int fa = ij_assemble_host(range[0], range[1], range[2], range[3], (int64_t) rows.size(), rows.data(), cols.data(),
                                   vals.data(), 0, h, false);
   if (fa && ctx().nranks == 1) return fa;
   if (fa) h = IjHost();
   const int fc = ij_create_parcsr(h, A);
   return fr ? fr : (fa ? fa : fc);
}

}  // namespace hb

using namespace hb;

extern "C" {

int hb200_parcsr_from_ij(hb200_parcsr **A, int64_t ilower, int64_t iupper, int64_t jlower, int64_t jupper,
                         int64_t num_entries, const int64_t *rows, const int64_t *cols, const double *values,
                         int add_duplicates)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A != nullptr && (num_entries == 0 || (rows && cols && values)), HB200_ERROR_ARG, "hb200_parcsr_from_ij: null argument");
   Ctx &c = ctx();
   // entries for rows of other ranks (HYPRE_IJMatrixAddToValues off-processor; hypre_IJMatrixRead adds every entry of a
   // part file that lies outside the part's row range): they travel to the owner, who ADDS them after its own entries
   // (hypre_IJMatrixAssembleOffProcValsParCSR, src/IJ_mv/IJMatrix_parcsr.c)
   std::vector<int64_t> own_r, own_c, far;       // far: (row, col, value bits) triples
   std::vector<double> own_v;
   bool any_far = false;
   for (int64_t k = 0; k < num_entries; k++) if (rows[k] < ilower || rows[k] > iupper) { any_far = true; break; }
   int64_t far_count = 0;
   if (any_far) {
      for (int64_t k = 0; k < num_entries; k++) {
         if (rows[k] < ilower || rows[k] > iupper) {
            int64_t bits;
            memcpy(&bits, &values[k], 8);
            far.push_back(rows[k]); far.push_back(cols[k]); far.push_back(bits);
         } else { own_r.push_back(rows[k]); own_c.push_back(cols[k]); own_v.push_back(values[k]); }
      }
      far_count = (int64_t) far.size() / 3;
      HB_REQUIRE(c.nranks > 1, HB200_ERROR_ARG, "hb200_parcsr_from_ij: entries outside [ilower, iupper] on one rank");
   }
   std::vector<int64_t> counts;
   HB_CHECK(allgather_i64(&far_count, 1, counts));
   int64_t max_far = 0;
   for (int64_t v : counts) max_far = std::max(max_far, v);
   if (max_far == 0) {
      IjHost h;
      HB_CHECK(ij_assemble_host(ilower, iupper, jlower, jupper, num_entries, rows, cols, values, add_duplicates, h));
      return ij_create_parcsr(h, A);
   }
   if (!any_far) { own_r.assign(rows, rows + num_entries); own_c.assign(cols, cols + num_entries); own_v.assign(values, values + num_entries); }
   const int64_t n_own = (int64_t) own_r.size();
   std::vector<int64_t> all;
   far.resize((size_t) max_far * 3, 0);
   HB_CHECK(allgather_i64(far.data(), (size_t) max_far * 3, all));
   for (int q = 0; q < c.nranks; q++) {
      if (q == c.rank) continue;
      const int64_t *lst = all.data() + (size_t) q * (size_t) max_far * 3;
      for (int64_t t = 0; t < counts[(size_t) q]; t++) {
         if (lst[3 * t] < ilower || lst[3 * t] > iupper) continue;
         double v;
         memcpy(&v, &lst[3 * t + 2], 8);
         own_r.push_back(lst[3 * t]); own_c.push_back(lst[3 * t + 1]); own_v.push_back(v);
      }
   }
   IjHost h;
   HB_CHECK(ij_assemble_host(ilower, iupper, jlower, jupper, (int64_t) own_r.size(), own_r.data(), own_c.data(), own_v.data(),
                             add_duplicates, h, true, n_own));
   return ij_create_parcsr(h, A);
}

// the host half alone (no GPU, one rank's rows): what the assembly produces, for tests against the reference's
// ParCSR arrays.  Two calls: sizes first (arrays NULL), then the arrays.
int hb200_host_ij_assemble(int64_t ilower, int64_t iupper, int64_t jlower, int64_t jupper, int64_t num_entries,
                           const int64_t *rows, const int64_t *cols, const double *values, int add_duplicates,
                           int *diag_nnz, int *offd_nnz, int *num_cols_offd, int *diag_i, int *diag_j,
                           double *diag_data, int *offd_i, int *offd_j, double *offd_data, int64_t *col_map_offd)
{
   HB_REQUIRE(num_entries == 0 || (rows && cols && values), HB200_ERROR_ARG, "hb200_host_ij_assemble: null argument");
   IjHost h;
   HB_CHECK(ij_assemble_host(ilower, iupper, jlower, jupper, num_entries, rows, cols, values, add_duplicates, h));
   if (diag_nnz) *diag_nnz = (int) h.diag_j.size();
   if (offd_nnz) *offd_nnz = (int) h.offd_j.size();
   if (num_cols_offd) *num_cols_offd = (int) h.col_map_offd.size();
   if (diag_i) std::copy(h.diag_i.begin(), h.diag_i.end(), diag_i);
   if (diag_j) std::copy(h.diag_j.begin(), h.diag_j.end(), diag_j);
   if (diag_data) std::copy(h.diag_a.begin(), h.diag_a.end(), diag_data);
   if (offd_i) std::copy(h.offd_i.begin(), h.offd_i.end(), offd_i);
   if (offd_j) std::copy(h.offd_j.begin(), h.offd_j.end(), offd_j);
   if (offd_data) std::copy(h.offd_a.begin(), h.offd_a.end(), offd_data);
   if (col_map_offd) std::copy(h.col_map_offd.begin(), h.col_map_offd.end(), col_map_offd);
   return 0;
}

// the host half of the CommPkg construction (no GPU): the all-gathered ownership and need lists in, rank `me`'s
// CommPkg out.  Output arrays sized by the caller: num_ranks, num_ranks + 1, own[5 me + ...] bounds.
int hb200_host_ij_commpkg(int num_ranks, int me, const int64_t *own5, const int64_t *need, int64_t max_offd,
                          int *num_sends, int *send_procs, int *send_map_starts, int *send_map_elmts, int send_capacity,
                          int *num_recvs, int *recv_procs, int *recv_vec_starts)
{
   HB_REQUIRE(num_ranks >= 1 && me >= 0 && me < num_ranks && own5 && (max_offd == 0 || need) && num_sends && num_recvs,
              HB200_ERROR_ARG, "hb200_host_ij_commpkg: bad argument");
   std::vector<int> sp, sms, sme, rp, rvs;
   HB_CHECK(ij_commpkg_host(num_ranks, me, own5, need, max_offd, sp, sms, sme, rp, rvs));
   if ((int) sme.size() > send_capacity) return set_error(HB200_ERROR_ARG, "hb200_host_ij_commpkg: send_map_elmts needs %d entries", (int) sme.size());
   *num_sends = (int) sp.size(); *num_recvs = (int) rp.size();
   if (send_procs) std::copy(sp.begin(), sp.end(), send_procs);
   if (send_map_starts) std::copy(sms.begin(), sms.end(), send_map_starts);
   if (send_map_elmts) std::copy(sme.begin(), sme.end(), send_map_elmts);
   if (recv_procs) std::copy(rp.begin(), rp.end(), recv_procs);
   if (recv_vec_starts) std::copy(rvs.begin(), rvs.end(), recv_vec_starts);
   return 0;
}

int hb200_parcsr_read_ij(hb200_parcsr **A, const char *filename, int is_matrix_market)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A != nullptr && filename != nullptr, HB200_ERROR_ARG, "hb200_parcsr_read_ij: null argument");
   HB_REQUIRE(is_matrix_market != 1 || ctx().nranks == 1, HB200_ERROR_ARG, "hb200_parcsr_read_ij: a Matrix Market file is read by one rank");
   int64_t range[4];
   std::vector<int64_t> rows, cols;
   std::vector<double> vals;
   // a rank that cannot read its part still takes part in the collectives below (with an empty range)
   const int fr = (is_matrix_market == 2) ? ij_parse_binary(filename, ctx().rank, range, rows, cols, vals)
                                          : ij_parse_file(filename, is_matrix_market, ctx().rank, range, rows, cols, vals);
   if (fr && ctx().nranks == 1) return fr;
   if (fr) { range[0] = 0; range[1] = -1; range[2] = 0; range[3] = -1; rows.clear(); cols.clear(); vals.clear(); }
   // hypre_IJMatrixRead: rows this rank owns are set (the last value wins); a symmetric Matrix Market file may
   // mirror an entry onto itself only through I == J, which is skipped
   const int f = hb200_parcsr_from_ij(A, range[0], range[1], range[2], range[3], (int64_t) rows.size(), rows.data(), cols.data(), vals.data(), 0);
   return fr ? fr : f;
}

// sizes of a matrix this library assembled itself (the caller of hb200_parcsr_create knows them already)
int hb200_parcsr_info(const hb200_parcsr *A, int64_t *info12)
{
   HB_REQUIRE(A != nullptr && info12 != nullptr, HB200_ERROR_ARG, "hb200_parcsr_info: null argument");
   info12[0] = A->num_rows; info12[1] = A->num_cols; info12[2] = A->num_cols_offd;
   info12[3] = A->diag.nnz; info12[4] = A->offd.nnz;
   info12[5] = A->pkg.num_sends; info12[6] = A->pkg.num_recvs;
   info12[7] = A->pkg.num_sends > 0 ? A->pkg.send_map_starts[(size_t) A->pkg.num_sends] : 0;
   info12[8] = A->first_row; info12[9] = A->first_col; info12[10] = A->global_rows; info12[11] = A->global_cols;
   return 0;
}

int hb200_parcsr_print_ij(const hb200_parcsr *A, const char *filename)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A != nullptr && filename != nullptr, HB200_ERROR_ARG, "hb200_parcsr_print_ij: null argument");
   std::vector<int64_t> rows, cols;
   std::vector<double> vals;
   HB_CHECK(ij_download_coo(A, rows, cols, vals));
   char path[1024];
   snprintf(path, sizeof(path), "%s.%05d", filename, ctx().rank);
   FILE *f = fopen(path, "w");
   if (!f) return set_error(HB200_ERROR_GENERIC, "Error: can't open output file %s", path);
   fprintf(f, "%lld %lld %lld %lld\n", (long long) A->first_row, (long long) (A->first_row + A->num_rows - 1),
           (long long) A->first_col, (long long) (A->first_col + A->num_cols - 1));
   for (size_t k = 0; k < rows.size(); k++) fprintf(f, "%lld %lld %.14e\n", (long long) rows[k], (long long) cols[k], vals[k]);
   fclose(f);
   return 0;
}

// HYPRE_IJMatrixPrintBinary -> hypre_ParCSRMatrixPrintBinaryIJ (src/parcsr_mv/par_csr_matrix.c:1120-1400): lossless
// (fp64 values, 64-bit indices as a HYPRE_BigInt build writes them); collective (the header holds the global count)
int hb200_parcsr_print_ij_binary(const hb200_parcsr *A, const char *filename)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A != nullptr && filename != nullptr, HB200_ERROR_ARG, "hb200_parcsr_print_ij_binary: null argument");
   std::vector<int64_t> rows, cols, all;
   std::vector<double> vals;
   HB_CHECK(ij_download_coo(A, rows, cols, vals));
   const int64_t mine = (int64_t) rows.size();
   HB_CHECK(allgather_i64(&mine, 1, all));
   uint64_t gnnz = 0;
   for (int64_t v : all) gnnz += (uint64_t) v;
   char path[1024];
   snprintf(path, sizeof(path), "%s.%05d.bin", filename, ctx().rank);
   FILE *f = fopen(path, "wb");
   if (!f) return set_error(HB200_ERROR_GENERIC, "Could not open output file %s", path);
   // index width: 4 bytes while the global sizes fit (what a default hypre build, 32-bit HYPRE_BigInt, writes), else 8
   const bool small = A->global_rows < 0x7fffffffLL && A->global_cols < 0x7fffffffLL;
   const uint64_t header[11] = {1, small ? 4u : 8u, 8, (uint64_t) A->global_rows, (uint64_t) A->global_cols, gnnz, (uint64_t) mine,
                                (uint64_t) A->first_row, (uint64_t) (A->first_row + A->num_rows - 1),
                                (uint64_t) A->first_col, (uint64_t) (A->first_col + A->num_cols - 1)};
   bool ok = fwrite(header, sizeof(uint64_t), 11, f) == 11;
   if (small) {
      std::vector<uint32_t> b(rows.size());
      for (size_t k = 0; k < rows.size(); k++) b[k] = (uint32_t) rows[k];
      ok = ok && fwrite(b.data(), 4, b.size(), f) == b.size();
      for (size_t k = 0; k < cols.size(); k++) b[k] = (uint32_t) cols[k];
      ok = ok && fwrite(b.data(), 4, b.size(), f) == b.size();
   } else {
      ok = ok && fwrite(rows.data(), 8, rows.size(), f) == rows.size();
      ok = ok && fwrite(cols.data(), 8, cols.size(), f) == cols.size();
   }
   ok = ok && fwrite(vals.data(), 8, vals.size(), f) == vals.size();
   fclose(f);
   if (!ok) return set_error(HB200_ERROR_GENERIC, "Could not write %s", path);
   return 0;
}

// vectors: hypre_ParVectorPrintIJ / hypre_IJVectorRead ("<name>.<5-digit rank>": `jlower jupper`, then `j value`)
int hb200_vector_print_ij(const double *x_dev, int64_t jlower, int num_values, const char *filename)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(filename != nullptr && (num_values == 0 || x_dev), HB200_ERROR_ARG, "hb200_vector_print_ij: null argument");
   Ctx &c = ctx();
   std::vector<double> h((size_t) num_values);
   HB_CUDA(cudaStreamSynchronize(c.s_comp));
   if (num_values) HB_CUDA(cudaMemcpy(h.data(), x_dev, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
   char path[1024];
   snprintf(path, sizeof(path), "%s.%05d", filename, c.rank);
   FILE *f = fopen(path, "w");
   if (!f) return set_error(HB200_ERROR_GENERIC, "Error: can't open output file %s", path);
   fprintf(f, "%lld %lld\n", (long long) jlower, (long long) (jlower + num_values - 1));
   for (int j = 0; j < num_values; j++) fprintf(f, "%lld %.14e\n", (long long) jlower + j, h[(size_t) j]);
   fclose(f);
   return 0;
}

int hb200_vector_read_ij(const char *filename, int64_t *jlower, int64_t *jupper, double *x_dev, int capacity)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(filename != nullptr, HB200_ERROR_ARG, "hb200_vector_read_ij: null argument");
   Ctx &c = ctx();
   char path[1024];
   snprintf(path, sizeof(path), "%s.%05d", filename, c.rank);
   FILE *f = fopen(path, "r");
   if (!f) return set_error(HB200_ERROR_ARG, "cannot open %s", path);
   long long lo, hi;
   if (fscanf(f, "%lld %lld", &lo, &hi) != 2) { fclose(f); return set_error(HB200_ERROR_GENERIC, "%s: no IJ vector header", path); }
   if (jlower) *jlower = lo;
   if (jupper) *jupper = hi;
   if (!x_dev) { fclose(f); return 0; }                  // size query
   const long long n = hi - lo + 1;
   if (n > capacity) { fclose(f); return set_error(HB200_ERROR_ARG, "hb200_vector_read_ij: %lld values, room for %d", n, capacity); }
   std::vector<double> h((size_t) (n > 0 ? n : 0), 0.0);
   long long j;
   double v;
   int ret;
   while ((ret = fscanf(f, "%lld %le", &j, &v)) != EOF) {
      if (ret != 2 || j < lo || j > hi) { fclose(f); return set_error(HB200_ERROR_GENERIC, "Error in IJ vector input file %s", path); }
      h[(size_t) (j - lo)] = v;
   }
   fclose(f);
   if (n > 0) HB_CUDA(cudaMemcpy(x_dev, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
   return 0;
}

}  // extern "C"
