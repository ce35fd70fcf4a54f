// Internal launch interfaces between the engine's translation units.
#pragma once
#include <vector>

#include "common.h"

namespace gwsem {

// ---- decode.cu
void launch_decode_2bit(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int n_samples,
                        int plink1, const int32_t* d_dest_of, int dest_rows, double* d_out, double* d_maf);

// ---- stats.cu : genotype-class sums of SNP-independent per-person features
// Device-side description of the people axis, shared by all statistics kernels.
struct PeopleLayout {
  int n_src = 0;        // people in the genotype file (source rows)
  int n_pad = 0;        // n_src rounded up to a multiple of 32
  int n_incl = 0;       // people that enter the model (rowFilter == 0 and complete phenotype row)
  int n_kept = 0;       // people kept by rowFilter (the provider's destRows; MAF is over these)
  const uint8_t* d_inc8 = nullptr;                // n_pad: 3 = included, 0 = excluded
  const unsigned long long* d_incmask = nullptr;  // n_pad/32 words, 0b01 per included person
  const unsigned long long* d_rowmask = nullptr;  // n_pad/32 words, 0b01 per person kept by rowFilter
};

// classes 1 and 2 (het, hom-ALT) sums of `n_feat` feature rows (feature-major, n_pad doubles each):
// d_sums[snp][2][n_feat].  Launches one kernel per chunk of <= 8 features.
void launch_class_sums(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1,
                       const PeopleLayout& pl, const double* d_feat, int n_feat, double* d_sums);

// The same sums as an exact int8 tensor-core contraction (stats_imma.cu): features as four balanced
// base-256 digits of round(h * 2^s_f), packed by imma_pack_features; d_inv_scale[f] = 2^-s_f.
int imma_chunk_features(int n_feat);
size_t imma_feature_bytes(int n_feat, int n_pad);
// inc8 / kept (n_pad each, non-zero = member): when the last chunk has a free feature slot its first two
// columns hold these indicators and the kernel also writes the het / hom-ALT counts of both sets into
// d_counts[snp][0,1,3,4] (imma_has_counts) and, when d_miss_list is given, appends the SNPs that have a code-3
// (PLINK 1: code-1) call anywhere in their row to d_miss_list (count in *d_miss_n, zeroed by the caller);
// launch_count_missing(light, d_list, d_list_n) then visits only those rows
bool imma_has_counts(int n_feat);
void imma_pack_features(const std::vector<std::vector<int64_t>>& q, int n_pad, const std::vector<uint8_t>& inc8,
                        const std::vector<uint8_t>& kept, std::vector<int8_t>& out);
void launch_class_sums_imma(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1,
                            const PeopleLayout& pl, const int8_t* d_featq, const double* d_inv_scale, int n_feat,
                            double* d_sums, int32_t* d_counts, int32_t* d_miss_list = nullptr,
                            int32_t* d_miss_n = nullptr);

// The same contraction on the 5th-generation tensor cores (stats_tc5.cu: tcgen05.mma kind::i8, one-hot A operand
// written to tensor memory); its own digit layout, same buffer size.  cs_use_tc5(n_feat): which of the two the engine
// uses for a model (tcgen05 for chunks of 64 digit columns, mma.sync for 16; GWSEM_CS_KERNEL=imma|tc5 forces one).
bool cs_use_tc5(int n_feat);
void tc5_pack_features(const std::vector<std::vector<int64_t>>& q, int n_pad, const std::vector<uint8_t>& inc8,
                       const std::vector<uint8_t>& kept, std::vector<int8_t>& out);
void launch_class_sums_tc5(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1,
                           const PeopleLayout& pl, const int8_t* d_featq, const double* d_inv_scale, int n_feat,
                           double* d_sums, int32_t* d_counts, int32_t* d_miss_list = nullptr,
                           int32_t* d_miss_n = nullptr);

constexpr int kTcWideFeatures = 32;     // features per wide chunk (4 digits each: 128 digit columns)
constexpr int kTcNarrowFeatures = 48;   // features per narrow chunk (3 digits each: 144 digit columns)
size_t tc5_wide_chunk_bytes(int n_pad);
size_t tc5_narrow_chunk_bytes(int n_pad);
void tc5_pack_chunk(const int64_t* q, int fi0, int n_in_chunk, int n_pad, int8_t* chunk_base, bool narrow);
void launch_class_sums_tc5_wide(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1,
                                const PeopleLayout& pl, const int8_t* d_featq, const double* d_inv_scale, int n_wide,
                                int n_narrow, double* d_sums);

// counts[snp][8] = {n_class1, n_class2, n_missing over included people; the same three over the
// people kept by rowFilter (for MAF); 0; 0} and, for the missing class, sums of `n_featm`
// person-major features (d_featT[person][n_featm_pad]): d_miss[snp][n_featm].
void launch_count_missing(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1,
                          const PeopleLayout& pl, const double* d_featT, int n_featm, int n_featm_pad,
                          int32_t* d_counts, double* d_miss, bool light = false, const int32_t* d_list = nullptr,
                          const int32_t* d_list_n = nullptr);

// ---- fit.cu
struct FitProgram;  // device-resident model description, built in model.cu
// Raw moment table of every SNP: d_R[snp][a * nf + f] = sum over complete rows of g^a * h_f
// (a = 0..4), d_n[snp] = complete rows.
// (i) hardcalls: from class counts / class sums / missing-class sums; also fills d_maf
void launch_assemble_moments(gwsem_engine* e, const FitProgram& fp, const PeopleLayout& pl, int n_snps,
                             const int32_t* d_counts, const double* d_sums, const double* d_miss, double* d_R,
                             int32_t* d_n, double* d_maf);
// (ii) any genotype column (dosages, bgen): d_geno[snp][n_pad] doubles in file order, NaN = missing
void launch_power_sums(gwsem_engine* e, const FitProgram& fp, const PeopleLayout& pl, int n_snps,
                       const double* d_geno, const double* d_featT, int n_featm_pad, double* d_R, int32_t* d_n,
                       int moment_order = 4);
void launch_fit_cumulants(gwsem_engine* e, const FitProgram& fp, int n_snps, const double* d_R, const int32_t* d_n,
                          const double* d_maf, const gwsem_result& d_res);

// ML / FIML from the raw moment table (order-2 moment program over manifest variables + covariates)
void launch_fit_ml(gwsem_engine* e, const FitProgram& fp, int n_snps, const double* d_R, const int32_t* d_n,
                   const double* d_maf, const gwsem_result& d_res);

// ---- marginals.cu / marg_fit.cu : WLS "marginals" path (ordinal items, exogenous covariates)
struct MargProgram;
void launch_marg_stats(gwsem_engine* e, const MargProgram& mp, const uint8_t* d_rows, int64_t pitch, int n_snps,
                       int n_src, const uint8_t* d_inc8, const int32_t* d_counts, int plink1, const double* d_geno,
                       int n_pad, double* d_summary);
void launch_maf_from_counts(gwsem_engine* e, const int32_t* d_counts, int n_kept, int n_snps, double* d_maf);
// Jacobian of the model-implied statistics at two generic points (d_out[2][n_stat][P]): its zero pattern is structural
void launch_marg_pattern(gwsem_engine* e, const FitProgram& fp, const MargProgram& mp, double* d_out);
void launch_marg_fit(gwsem_engine* e, const FitProgram& fp, const MargProgram& mp, int n_snps, const double* d_summary,
                     const double* d_maf, const gwsem_result& d_res);

}  // namespace gwsem
