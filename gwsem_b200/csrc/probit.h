// buildItem() with an ordinal phenotype: the SNP is an exogenous predictor (probit.cu).
#pragma once
#include <cstdint>

#include "common.h"

namespace gwsem {

struct ProbitProgram {
  int n_cat = 0;        // categories of the item
  int K = 0, Kp = 1;    // covariates besides the SNP
  int P = 0;            // model parameters
  const uint8_t* cat = nullptr;    // [n_pad] file order
  const double* X = nullptr;       // [n_pad][Kp]
  const double* theta0 = nullptr;  // no-SNP fit: C-1 thresholds, K slopes
  const int32_t* par_map = nullptr;  // model parameter -> index in (thresholds, snp slope, covariate slopes)
  const double* par_start = nullptr;
  double min_variance = 0.0;
};

void launch_probit_snp(gwsem_engine* e, const ProbitProgram& pp, const uint8_t* d_rows, int64_t pitch, int n_snps,
                       int n_src, const uint8_t* d_inc8, int plink1, const double* d_geno, int n_pad, const double* d_maf,
                       const gwsem_result& d_res);

}  // namespace gwsem
