// WLS "marginals" summary statistics (ordinal items / exogenous covariates), SURVEY.md section 8a
// row O4.  Host precompute (marginals_host.cpp) + per-SNP kernels (marginals.cu).
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

namespace gwsem {

constexpr int kMargMaxItems = 10;

struct MargItem {
  int n_cat = 0;                 // 0 = continuous, else ordinal with codes 0..n_cat-1
  const double* values = nullptr; // model data rows (codes for ordinal)
  int n_values = 0;
  // stage-1 estimates
  std::vector<double> th;        // ordinal thresholds
  std::vector<double> beta;      // continuous: intercept + slopes; ordinal: slopes
  double var = 1.0;              // continuous residual variance (ML)
  int n_par = 0;                 // ordinal: C-1+K ; continuous: K+2
  std::vector<double> hess;      // n_par x n_par: minus the exact second derivative of the log-likelihood at the estimate
  void fit(const std::vector<int>& rows, const std::vector<const double*>& X);
};

struct MarginalsHost {
  std::vector<MargItem> items;
  std::vector<const double*> X;              // exogenous covariates, model data rows
  // results of build()
  int n1 = 0, q0 = 0, n_rows_total = 0;
  std::vector<int> col_of;                   // first stage-1 column of each item
  std::vector<std::pair<int, int>> pairs;    // item pairs (a > b), column-major lower triangle
  std::vector<double> rho;                   // per pair
  std::vector<double> sc0;                   // [row][q0]: stage-1 scores of every item, then pair scores
  std::vector<double> z1, z0;                // [row][J]: ordinal: standardised thresholds; continuous: z1 = std residual
  std::vector<double> A21;                   // [pair][n1]
  std::vector<double> B00;                   // [q0][q0]
  void build(const std::vector<int>& rows);
};

// Device-resident program for the per-SNP kernels (plain struct, passed by value).
struct MargProgram {
  int J = 0, K = 0, Kp = 0, q0 = 0, q0p = 0, n1 = 0, npii = 0, n_incl = 0;
  int n_cat[kMargMaxItems] = {0};
  int col_of[kMargMaxItems + 1] = {0};
  double item_sd[kMargMaxItems] = {0};       // continuous items: residual sd
  double zmax[kMargMaxItems] = {0};          // ordinal items: largest |standardised threshold| any person carries (sentinels excluded)
  const double* sc0 = nullptr;               // [n_pad][q0p] file order, zero rows for excluded people
  const double* zz = nullptr;                // [n_pad][J][5]: z1, z0, phi(z1), phi(z0), Phi(z1)-Phi(z0) (continuous: v)
  const double* cc = nullptr;                // [n_pad][J][3]: tetrachoric-series coefficients c1..c3 (continuous: v)
  const uint8_t* cat = nullptr;              // [n_pad][J]
  const double* X = nullptr;                 // [n_pad][Kp]
  const double* B00 = nullptr;               // [q0][q0]
  const double* A21ii = nullptr;             // [npii][n1]
  const double* Wcc = nullptr;               // [q0][q0]: inverse of the item-side block of Muthen's A (marg_fit_fast_kernel)
  int diag = 0;                              // GWSEM_FIT_DIAG: per-phase cycle counters
  int fast_fit = 0;                          // set per launch: marg_fit_fast_kernel takes the SNPs without a full record
  const double* acov_cc = nullptr;           // [q0][q0]: Wcc B00 Wcc', the SNP-independent block of acov(theta)
  const double* theta1 = nullptr;            // n1 stage-1 estimates of the items, in column order
  const double* rho_ii = nullptr;            // npii
  const double* H11 = nullptr;               // exact stage-1 information of the items, blocks of n_par^2 back to back
  int h11_of[kMargMaxItems + 1] = {0};       // offset of each item's block in H11
  // statistic vector layout (built on the host): kind 0 mean, 1 threshold, 2 slope, 3 variance, 4 covariance
  int n_stat = 0;
  const int32_t* stat_kind = nullptr;
  const int32_t* stat_a = nullptr;           // manifest variable
  const int32_t* stat_b = nullptr;           // threshold index / covariate index / second variable
  const int32_t* stat_theta = nullptr;       // index into theta (snp 2 | items n1 | rho_j0 J | rho_ii)
  // parameters that enter exactly one statistic are eliminated from the iteration (marg_fit.cu):
  // pairs (elim_a_par[i], elim_a_row[i]); the iteration runs on parameters elim_b_par over rows elim_b_row
  int n_a = 0, n_b = 0, n_beta = 0;
  const int32_t* elim_a_par = nullptr;
  const int32_t* elim_a_row = nullptr;
  const int32_t* elim_b_par = nullptr;
  const int32_t* elim_b_row = nullptr;
  // iterated parameters that are the mean / variance of a continuous variable outright (unit derivative of that
  // statistic, e.g. the snp mean and variance): the structured fit moves them onto the observed value before iterating
  int n_init = 0;
  int init_par[4] = {0, 0, 0, 0}, init_row[4] = {0, 0, 0, 0};
  // series pass at rho = 0 from genotype-class sums (hardcalls without missing calls, all items ordinal): per item the
  // six products {c1, c2, c1^2, c3, c1^3, c1 c2} of the tetrachoric coefficients, class-summed exactly by the tcgen05
  // kernel; cs_sums[snp][2][6 J] (het, hom ALT) is set per launch, cs_tot[6 J] are the sums over all included people
  int cs_nf = 0;
  const double* cs_sums = nullptr;
  const double* cs_tot = nullptr;
  // Series path (marg_series.cu): for hardcalls without a missing call every SNP-dependent statistic is a power series
  // in rho_j whose coefficients are polynomials in the standardised genotype (Mehler expansion), i.e. a combination of
  // genotype-class sums of SNP-independent per-person features.  se_sums[snp][2][se_nf] (het, hom ALT; per launch),
  // se_tot[se_nf] the sums over all included people; feature offsets below (pair tables in marg_series.h).
  // se_done[snp] = 1: the record of this SNP is complete, marg_stats_kernel leaves at once.
  int se_nf = 0, se_n_wide = 0, se_n_narrow = 0;   // features in all; in four-digit (wide) and three-digit (narrow) chunks
  const double* se_sums = nullptr;
  const double* se_tot = nullptr;
  const double* se_err = nullptr;            // [J][10]: second moments of the first neglected coefficients (acceptance bound)
  uint8_t* se_done = nullptr;
  int se_o_psi0[kMargMaxItems] = {0};        // psi, weight 1, orders 0..3: 8 pairs   } the score polynomial: four digits
  int se_o_hi[kMargMaxItems] = {0};          // psi, weight 1, orders 4 and 5: 7 pairs } (everything below: three)
  int se_o_psi[kMargMaxItems] = {0};         // psi times SC0 column b: [b][8 pairs]
  int se_o_f3[kMargMaxItems] = {0};          // 5 pairs
  int se_o_thr[kMargMaxItems] = {0};         // [threshold][8 pairs]
  int se_o_slope[kMargMaxItems] = {0};       // [covariate][8 pairs]
  int se_o_diag[kMargMaxItems] = {0};        // psi_j^2: 10 terms
  int se_o_cross = 0;                        // [pair (a > b), column-major lower triangle][27 terms]
  int se_o_sc0 = 0;                          // plain SC0 columns: q0
  // summary record per SNP (doubles)
  int sum_stride = 0, o_rho = 0, o_a22 = 0, o_a21snp = 0, o_a21item = 0, o_bdd = 0, o_bd0 = 0;
  int o_msum = 0, o_mmat = 0;  // missing-call people: sum of their item-side score rows, and of the outer products
  // Full records (marg_exact.cu): SNPs whose every stage-1 / stage-2 estimate and score was recomputed over the people
  // retained for that SNP carry (theta, A21, B) themselves; marg_fit reads those instead of assembling them from the
  // SNP-independent item side.  full_all: every SNP of the launch has one (models without a static item side);
  // otherwise the SNPs with a missing call do (exact naAction = 'omit').  The theta layout is the oracle's: stage-1
  // blocks of the manifest variables in order (continuous: intercept, slopes, variance; ordinal: thresholds, slopes),
  // then one correlation per pair, column-major lower triangle.
  int full_all = 0, exact_missing = 0;
  int p = 0, nt = 0, n1t = 0;                // manifest variables, length of theta, stage-1 parameters in it
  int vcat[kMargMaxItems + 2] = {0};         // categories per manifest variable (0 = continuous)
  int voff[kMargMaxItems + 3] = {0};         // first theta index of each manifest variable's block
  int full_stride = 0, fo_theta = 0, fo_a21 = 0, fo_B = 0;
  const double* full = nullptr;              // [snp][full_stride]: n, catch code, catch variable, theta, A21, B
};

// The generic per-SNP marginals kernel (marg_exact.cu): every variable and covariate may depend on the SNP.
constexpr int kExMaxVars = kMargMaxItems + 2;
constexpr int kExMaxCov = 12;
struct ExactProgram {
  int p = 0, K = 0, Kp = 1, nt = 0, n1 = 0, npairs = 0, T = 256, only_missing = 0, n_incl = 0, snp_exo = 0;
  double min_variance = 0.0;
  int vkind[kExMaxVars] = {0};               // 0 static continuous, 1 static ordinal, 2 snp, 3 snp x static column
  int ncat[kExMaxVars] = {0};
  int voff[kExMaxVars + 1] = {0};
  int nx[kExMaxVars] = {0};                  // free covariates of each variable
  int xidx[kExMaxVars][kExMaxCov] = {{0}};
  int xkind[kExMaxCov] = {0};                // 0 static column, 1 snp, 2 snp x static column
  const double* V = nullptr;                 // [n_pad][p]: value / category code / moderator of each variable
  const double* X = nullptr;                 // [n_pad][Kp]
  const double* theta0 = nullptr;            // nt: starting point (the SNP-independent estimates where there are any)
  int gl_n[4] = {0, 0, 0, 0};                // Gauss-Legendre rules for the bivariate normal integral
  const double* gl = nullptr;                // nodes then weights of each rule, back to back
  int full_stride = 0, fo_theta = 0, fo_a21 = 0, fo_B = 0;
};

}  // namespace gwsem
