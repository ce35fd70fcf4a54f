// Device-resident description of one model for the per-SNP kernels (built by model.cu).
#pragma once
#include <cstdint>

namespace gwsem {

// Plain struct of sizes and device pointers; passed to kernels by value.
struct FitProgram {
  int p = 0;        // manifest variables
  int q = 0;        // p(p+1)/2 covariance statistics
  int nv = 0;       // RAM variables
  int P = 0;        // free parameters
  int n_cells = 0;
  int nf = 0;       // features (monomials of SNP-independent columns); feature 0 is the constant 1
  int nfd = 0;      // features 1..nfd have class sums (needed with a genotype power >= 1)
  int nfm = 0;      // features 1..nfm have missing-class sums (= nf - 1)
  int n_incl = 0;   // people entering the model
  double min_variance = 0.0;
  const double* tot = nullptr;       // nf: feature totals over included people
  const int32_t* idx4 = nullptr;     // (p+1)^4: sorted multiset of manifest vars -> a * nf + f
  const int32_t* stat_i = nullptr;   // q: row of statistic e (vech, column-major lower triangle)
  const int32_t* stat_j = nullptr;   // q: column
  const int32_t* cell_mat = nullptr; // n_cells: 0 A, 1 S, 2 M
  const int32_t* cell_row = nullptr;
  const int32_t* cell_col = nullptr;
  const int32_t* cell_par = nullptr; // free parameter or -1
  const double* cell_val = nullptr;  // value when fixed
  const int32_t* par_cell_ptr = nullptr;  // P+1: CSR into par_cells
  const int32_t* par_cells = nullptr;
  const double* par_start = nullptr;
  const double* par_lbound = nullptr;
  const double* par_warm = nullptr;  // marginals: estimates of a reference SNP, the iteration's starting point (else par_start)
  // marginals only: per statistic, the free threshold parameter (or -1 and its fixed value);
  // RAM index of each exogenous covariate's latent variable
  const int32_t* thr_par = nullptr;
  const double* thr_val = nullptr;
  const int32_t* exo_latent = nullptr;
  // ML only: covariates (moment variables p .. p + n_exo - 1), which manifest variables are
  // SNP-independent, whether the genotype is an exogenous covariate (and which one)
  int n_exo = 0, snp_exo = 0, snp_exo_index = 0;
  const int32_t* man_static = nullptr;
};

size_t fit_scratch_doubles(int p, int q, int nv, int P, int nf);

}  // namespace gwsem
