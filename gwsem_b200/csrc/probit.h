// buildItem() with an ordinal phenotype: the SNP is an exogenous predictor (probit.cu).
#pragma once
#include <cstdint>

#include "common.h"

namespace gwsem {

constexpr int kProbitMaxX = 15;

struct ProbitProgram {
  int n_cat = 0;        // categories of the item
  int K = 0;            // definition variables (snp, its gxe products, covariates), in the model's order
  int Kp = 1;           // row stride of X (its SNP-independent data columns)
  int P = 0;            // model parameters
  int32_t kind[kProbitMaxX] = {0};  // 0 data column, 1 snp, 2 snp * data column
  int32_t col[kProbitMaxX] = {0};   // data column (kind 0 and 2)
  const uint8_t* cat = nullptr;    // [n_pad] file order
  const double* X = nullptr;       // [n_pad][Kp]
  const double* theta0 = nullptr;  // start: C-1 thresholds and K slopes from the fit without the SNP (0 for snp terms)
  const int32_t* par_map = nullptr;  // model parameter -> index in (thresholds, K slopes)
  const double* par_start = nullptr;
  double min_variance = 0.0;
};

void launch_probit_snp(gwsem_engine* e, const ProbitProgram& pp, const uint8_t* d_rows, int64_t pitch, int n_snps,
                       int n_src, const uint8_t* d_inc8, int plink1, const double* d_geno, int n_pad, const double* d_maf,
                       const gwsem_result& d_res);

}  // namespace gwsem
