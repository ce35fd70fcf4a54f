// buildItem() with a continuous phenotype and exogenous = TRUE under WLS: the item is regressed on
// definition variables that include the SNP and its gxe products (linreg.cu).
#pragma once
#include <cstdint>

#include "common.h"

namespace gwsem {

constexpr int kLinregMaxX = 15;  // regressors besides the intercept

struct LinregProgram {
  int K = 0;            // regressors besides the intercept
  int Ks = 0, Ksp = 1;  // SNP-independent data columns (covariates, moderators) and their row stride
  int P = 0;            // model parameters
  int32_t kind[kLinregMaxX] = {0};  // 0 data column, 1 snp, 2 snp * data column
  int32_t col[kLinregMaxX] = {0};   // data column (kind 0 and 2)
  const double* y = nullptr;         // [n_pad] file order
  const double* Xs = nullptr;        // [n_pad][Ksp]
  const int32_t* par_map = nullptr;  // model parameter -> index in (intercept, K slopes, residual variance)
  const double* par_start = nullptr;
  double min_variance = 0.0;
};

void launch_linreg_snp(gwsem_engine* e, const LinregProgram& lp, const uint8_t* d_rows, int64_t pitch, int n_snps,
                       int n_src, const uint8_t* d_inc8, int plink1, const double* d_geno, int n_pad, const double* d_maf,
                       const gwsem_result& d_res);

}  // namespace gwsem
