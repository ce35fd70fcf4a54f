// gwsem_model: host-side state of one model bound to an engine (see model.cu).
#pragma once
#include <vector>

#include "common.h"
#include "fit_program.h"
#include "kernels.h"
#include "marginals.h"
#include "probit.h"
#include "linreg.h"

struct gwsem_model {
  gwsem_engine* engine;
  // specification
  int nv = 0, p = 0, P = 0, q = 0, n_rows = 0;
  double min_variance = 0;
  std::vector<double> cells, par_start, par_lbound;
  std::vector<int32_t> kind, var_base;
  std::vector<std::vector<double>> base;  // SNP-independent columns (centred unless used as moderator)
  std::vector<uint8_t> row_ok;            // complete phenotype row
  // moment program
  int nf = 0, nfd = 0, nfm = 0, nfm_pad = 0;
  int pm = 0;             // moment variables: the manifest variables (+ the covariates under ML)
  int moment_order = 4;   // 4: cumulants (fourth-order moments); 2: ML
  std::vector<std::vector<int>> feature_monos;
  std::vector<int32_t> idx4, stat_i, stat_j;
  // people axis / device state
  bool prepared = false;
  int n_src = 0, n_pad = 0, n_incl = 0;
  gwsem::FitProgram fp;
  gwsem::PeopleLayout layout;
  const double* d_feat = nullptr;
  const double* d_featT = nullptr;
  const int8_t* d_featq = nullptr;        // dense features as base-256 digits (stats_imma.cu)
  const double* d_inv_scale = nullptr;    // 2^-s_f per dense feature
  const int32_t* d_dest_of = nullptr;
  std::vector<gwsem::DeviceBuffer<uint8_t>> pool;
  gwsem::DeviceBuffer<int32_t> d_counts, d_miss_list;
  bool use_tc5 = false;   // which class-sum kernel the packed digits were laid out for (cs_use_tc5 at prepare time)
  gwsem::DeviceBuffer<double> d_sums, d_miss, d_R, d_mafs;
  gwsem::DeviceBuffer<int32_t> d_nrows;
  gwsem::DeviceBuffer<uint8_t> d_stage2[2];
  cudaStream_t copy_stream = nullptr, out_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_fitted[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  gwsem::DeviceBuffer<double> r_est, r_vcov, r_fit, r_maf;
  gwsem::DeviceBuffer<int32_t> r_n, r_status, r_catch, r_cvar, r_iter;

  // WLS marginals (fit == 1), ML (fit == 2)
  int fit_kind = 0;
  void prepare_ml();
  void prepare_warm();
  int marg_mode = 0;   // see the constructor (model.cu)
  int snp_exo = -1;    // which exogenous covariate is the genotype column (mode 1)
  gwsem::ProbitProgram pp;
  gwsem::LinregProgram lp;   // marg_mode 2: one continuous item on definition variables that include the snp
  std::vector<int32_t> exo_kind;
  std::vector<int32_t> n_categories, exo_latent;
  std::vector<std::vector<double>> exo_cols;
  std::vector<uint8_t> exo_free;
  std::vector<double> thresholds;  // n x 4
  gwsem::MarginalsHost mh;
  gwsem::MargProgram mp;
  gwsem::ExactProgram ex;   // marg_exact_kernel: SNPs with a missing call (marg_mode 0), every SNP (marg_mode 3)
  gwsem::DeviceBuffer<double> d_summary, d_cs_sums, d_full, d_se_sums;
  gwsem::DeviceBuffer<uint8_t> d_se_done;
  const int8_t* d_seq = nullptr;          // marginals: digits of the series-path features (tcgen05 layout)
  const double* d_se_inv_scale = nullptr;
  cudaStream_t side_stream = nullptr;   // marg_exact_kernel next to marg_stats_kernel
  cudaEvent_t ev_side_go = nullptr, ev_side_done = nullptr;
  const int8_t* d_csq = nullptr;          // marginals: digits of the series products (tcgen05 layout)
  const double* d_cs_inv_scale = nullptr;
  void prepare_marginals(const std::vector<int>& dest_of);
  void exact_diag(int n);
  void run_marginals(const uint8_t* d_packed, int64_t pitch, const double* d_geno, const double* d_maf_in, int n,
                     int plink1, const gwsem_result& r);

  gwsem_model(gwsem_engine* e, const gwsem_model_spec* s);
  void build_program();
  void prepare(const int32_t* row_filter, int src_rows);
  void fit_2bit_device(const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1, const gwsem_result& d_res);
  void fit_rows_device(const double* d_geno, const double* d_maf, int n_snps, const gwsem_result& d_res);
  void fit_2bit_host(const uint8_t* packed, int64_t pitch, int n_snps, int plink1, const gwsem_result& h);
  void ensure_result_buffers(int n);
  gwsem_result device_result();
  void download_result(const gwsem_result& h, const gwsem_result& d, int n, int offset, cudaStream_t on = nullptr);
};
