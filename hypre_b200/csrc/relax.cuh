// relax.cuh — internal interface of the relaxation / coarse-solve translation units.
#pragma once
#include "hb_internal.cuh"

namespace hb {

bool relax_is_jacobi(int relax_type);
bool relax_is_gs(int relax_type);

int relax_jacobi_oop(hb200_parcsr *A, const double *f, const int *cf, int relax_type,
                     int relax_points, double w, const double *l1, const double *u_in,
                     double *u_out, bool zero_guess, bool *used_shortcut);

// in-place hybrid Gauss-Seidel family (gs.cu); vtemp: num_rows doubles of scratch
int relax_hybrid_gs(hb200_parcsr *A, const double *f, const int *cf, int relax_type,
                    int relax_points, double relax_weight, double omega, const double *l1,
                    double *u, double *vtemp);

int cheby_solve(hb200_parcsr *A, const double *f, const double *ds, const double *coefs_host,
                int order, int scale, double *u, double *v, double *r, double *orig_u, double *tmp);

struct GEData {
   int     n = 0, first_row = 0, num_local = 0;
   double *d_LfT = nullptr, *d_UT = nullptr, *d_Udiag = nullptr, *d_b = nullptr;
};
int ge_factor_host(const double *A_mat, int n, std::vector<double> &LfT, std::vector<double> &UT,
                   std::vector<double> &Udiag);
int ge_solve(const GEData &ge, const double *f_local, double *u_local);

}  // namespace hb
