/* gwsem_b200 -- C ABI of the B200-native per-SNP association engine.
 *
 * This is the drop-in boundary for gwsem's hot path.  Everything
 * is plain C: opaque handles, caller-owned buffers, `int` status (0 = ok) with the
 * message available from gwsem_last_error(); no exceptions, no torch types.
 * What each group replaces in the reference (paths relative to the gwsem repo):
 *
 *   gwsem_source_*      the two OpenMx data providers registered by setup2()
 *                       (src/loader.cpp:421-433): LoadDataPGENProvider2 (147-406) and
 *                       LoadDataBGENProvider2 (58-145) behind LoadDataProviderBase2
 *                       (src/LoadDataAPI.h:13-123): open, getNumVariants, getLoadCounter,
 *                       loadRow, addCheckpointColumns.
 *   gwsem_decode_*      loadRowImpl's arithmetic on already-read bytes: 2-bit / plink1 ->
 *                       double with rowFilter compaction (loader.cpp:211-248), dosage overlay
 *                       (252-273), MAF (386-394), bgen probabilities -> dosage (38-50).
 *   gwsem_model_* /     what mxRun() does per SNP inside OpenMx for the plan assembled by
 *   gwsem_fit_* /       prepareComputePlan (R/model.R:122-195): data refresh + naAction,
 *   gwsem_gwas_run      WLS summary statistics, RAM expectation, optimisation from the
 *                       original starts, standard errors, one checkpoint row per SNP.
 *
 * Pointers named d_* are device (HBM) pointers on the engine's GPU; all others are host
 * pointers.  The library fails with an error (never falls back to the CPU) when no CUDA
 * device is usable.
 */
#ifndef GWSEM_B200_H_
#define GWSEM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GWSEM_ABI_VERSION 1

/* R's NA_real_ bit pattern; every NA the engine emits carries it. */
#define GWSEM_NA_BITS 0x7FF00000000007A2ULL

typedef struct gwsem_engine gwsem_engine;
typedef struct gwsem_source gwsem_source;
typedef struct gwsem_model gwsem_model;

int gwsem_abi_version(void);
const char* gwsem_last_error(void);
int gwsem_device_count(void);

/* ------------------------------------------------------------------ engine (one per GPU) */
int gwsem_engine_create(int device, gwsem_engine** out);
void gwsem_engine_destroy(gwsem_engine* e);
/* cudaStream_t to launch on (NULL = the engine's own stream). */
int gwsem_engine_set_stream(gwsem_engine* e, void* cuda_stream);
int gwsem_engine_synchronize(gwsem_engine* e);
/* kernels launched by this engine since creation (bench.py's gpu_launches) */
int64_t gwsem_engine_launch_count(const gwsem_engine* e);
/* int8 digit columns contracted by the series-path class-sum launches of this engine since creation, summed over the
 * launches (128 or 144 per launch): int8 operations = 4 x people x SNPs per launch x this (bench.py's tensor roofline) */
int64_t gwsem_engine_series_columns(const gwsem_engine* e);
/* Per-launch timing with CUDA events on the launch stream.  gwsem_engine_profile_read fills
 * launches[8] / total_ms[8] for the kernel families {class sums, count+missing, fit, decode,
 * marginals statistics (per-person path), marginals statistics (exact path), series assembly, series class sums}
 * seen since the previous read, and clears them (synchronises the stream). */
#define GWSEM_PROFILE_FAMILIES 8
int gwsem_engine_profile(gwsem_engine* e, int enable);
int gwsem_engine_profile_read(gwsem_engine* e, int64_t* launches, double* total_ms);
/* Measured FP64 FMA throughput of this GPU right now (TFLOP/s, DFMA micro-benchmark without memory
 * traffic): the roofline denominator of the FP64-bound statistics kernels (bench.py). */
int gwsem_engine_fp64_peak(gwsem_engine* e, double* tflops);

/* ----------------------------------------------------------- genotype sources (providers) */
/* method: "pgen" (also PLINK 1 .bed, as R/model.R:36) or "bgen".  src_rows: number of people
 * the caller's data has before rowFilter (required for .bed, checked otherwise; <= 0 = take
 * it from the file).  For bgen the SNP order is the .bgi index order (src/bgen.cpp:649-663). */
int gwsem_source_open(const char* path, const char* method, int src_rows, gwsem_source** out);
void gwsem_source_close(gwsem_source* s);
int gwsem_source_num_variants(const gwsem_source* s);
int gwsem_source_num_samples(const gwsem_source* s);
int gwsem_source_load_counter(const gwsem_source* s);
/* names of the checkpoint columns the provider contributes (loader.cpp:70-80, 188-192), tab separated */
const char* gwsem_source_checkpoint_columns(const gwsem_source* s);
/* tab separated values of those columns for SNP `index` (0-based), MAF excluded */
int gwsem_source_context(gwsem_source* s, int index, char* buf, size_t buflen);

/* loadRow for a batch: out[n_snps * dest_rows] doubles (dest_rows = people with row_filter == 0),
 * maf[n_snps].  row_filter: src_rows ints, nonzero = skip (LoadDataAPI.h:80), or NULL.
 * snp_index: 0-based SNP numbers, any order, repeats allowed. */
int gwsem_source_load_rows(gwsem_engine* e, gwsem_source* s, const int32_t* snp_index, int n_snps,
                           const int32_t* row_filter, double* out, double* maf);

/* ------------------------------------------------------------------------ decode kernels */
/* n_snps rows of 2-bit genotypes, `pitch` bytes apart (pitch >= ceil(n_samples/4)).
 * plink1 != 0: PLINK 1 .bed coding (src/pgenlib_read.cc:1979-1992), else PLINK 2 coding.
 * d_dest_of: n_samples ints, destination row of each person or -1 (device), or NULL.
 * d_out: n_snps * dest_rows doubles; d_maf: n_snps doubles. */
int gwsem_decode_2bit_device(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps,
                             int n_samples, int plink1, const int32_t* d_dest_of, int dest_rows,
                             double* d_out, double* d_maf);
/* same from host memory: copies in, decodes, copies out */
int gwsem_decode_2bit(gwsem_engine* e, const uint8_t* packed, int64_t pitch, int n_snps, int n_samples,
                      int plink1, const int32_t* row_filter, double* out, double* maf);

/* ------------------------------------------------------------------------------- models */
typedef struct {
  int32_t n_vars;          /* manifest + latent, manifest first (RAM order) */
  int32_t n_manifest;
  int32_t n_params;
  int32_t n_cells;
  const double* cells;     /* n_cells x 5: matrix (0 A, 1 S, 2 M), row, col, param (-1 = fixed), value */
  const double* par_start; /* n_params (mxComputeSetOriginalStarts restores these per SNP) */
  const double* par_lbound;/* n_params, NaN = unbounded */
  int32_t fit;             /* 0 = WLS cumulants, 1 = WLS marginals, 2 = ML */
  double min_variance;     /* mxData(minVariance=) */
  /* observed data, n_rows people (after rowFilter) */
  int32_t n_rows;
  const int32_t* manifest_kind;       /* n_manifest: 0 data column, 1 snp, 2 snp * moderator */
  const double* const* manifest_data; /* n_manifest: column (kind 0), moderator (kind 2), NULL (kind 1) */
  /* ---- WLS marginals (fit == 1) and ML (fit == 2): ordinal items and exogenous covariates */
  const int32_t* n_categories;        /* n_manifest: 0 continuous, C >= 2 ordinal; manifest_data then holds codes 0..C-1 */
  int32_t n_exo;                      /* exogenous covariates (definition variables, R/model.R:403-418) */
  const double* const* exo_data;      /* n_exo columns of n_rows; NULL = the genotype itself (buildItem, exogenous = TRUE) */
  const int32_t* exo_latent;          /* n_exo: RAM index of the covariate's latent variable */
  const uint8_t* exo_free;            /* n_manifest x n_exo, row major: slope is a statistic */
  int32_t n_thresholds;
  const double* thresholds;           /* n_thresholds x 4: manifest, k, param (-1 = fixed), value */
  const int32_t* exo_kind;            /* n_exo or NULL: 0 data column, 1 the genotype (exo_data NULL), 2 genotype x exo_data column
                                         (a gxe product `snp_<moderator>` as definition variable, R/model.R:427-434) */
} gwsem_model_spec;

int gwsem_model_create(gwsem_engine* e, const gwsem_model_spec* spec, gwsem_model** out);
void gwsem_model_destroy(gwsem_model* m);
/* rowFilter for the genotype side: src_rows ints (nonzero = person absent from the model data). */
int gwsem_model_set_row_filter(gwsem_model* m, const int32_t* row_filter, int src_rows);

/* status codes in gwsem_result.status (text as OpenMx prints statusCode) */
enum { GWSEM_STATUS_OK = 0, GWSEM_STATUS_ITERATION_LIMIT = 4, GWSEM_STATUS_NOT_CONVEX = 5, GWSEM_STATUS_NONZERO_GRADIENT = 6,
       GWSEM_STATUS_NA = -1 };
/* reasons the per-SNP TryCatch fired (column catch1) */
enum { GWSEM_CATCH_NONE = 0, GWSEM_CATCH_MIN_VARIANCE = 1, GWSEM_CATCH_ACOV_NOT_PD = 2 };

/* Per-SNP results, structure of arrays, each of length n_snps (times the inner size noted). */
typedef struct {
  double* est;        /* n_snps * n_params */
  double* vcov;       /* n_snps * n_params * n_params */
  double* fit;        /* n_snps */
  double* maf;        /* n_snps */
  int32_t* n_used;    /* n_snps: rows left after naAction='omit' */
  int32_t* status;    /* n_snps */
  int32_t* catch_code;/* n_snps */
  int32_t* catch_var; /* n_snps: manifest index the catch refers to */
  int32_t* iterations;/* n_snps */
} gwsem_result;

/* The hot path on HBM-resident 2-bit hardcalls: statistics + fit for n_snps SNPs.
 * d_packed rows are `pitch` bytes apart, pitch a multiple of 16.  `res` holds device pointers. */
int gwsem_fit_2bit_device(gwsem_model* m, const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1,
                          const gwsem_result* d_res);
/* Same through host buffers (pinned or pageable): H2D of the packed rows, fit, D2H of results. */
int gwsem_fit_2bit(gwsem_model* m, const uint8_t* packed, int64_t pitch, int n_snps, int plink1,
                   const gwsem_result* res);
/* Generic path on decoded doubles (dosages, bgen): d_geno[n_snps * n_rows], NA allowed. */
int gwsem_fit_rows_device(gwsem_model* m, const double* d_geno, int n_snps, const gwsem_result* d_res);

/* -------------------------------------------------------------- the whole loop (GWAS()) */
typedef struct {
  const int32_t* snp;         /* 1-based SNP numbers (the SNP= argument) */
  int32_t n_snp;              /* length of snp; 0 = no SNP at all (header-only log); -1 = every SNP from start_from */
  int32_t start_from;         /* 1-based, used when n_snp == -1 */
  const char* out_path;       /* checkpoint log */
  const char* const* par_names;     /* n_params labels for the header */
  const char* context_path;   /* .pvar / .bim (mxComputeLoadContext), or NULL for bgen */
  const int32_t* vcov_pairs;  /* 2 * n_vcov_pairs parameter indices: off-diagonals to write */
  int32_t n_vcov_pairs;
  int32_t batch_snps;         /* SNPs per device batch, 0 = auto */
  const char* model_name;     /* for catch1 messages ("<model>.data: ...") */
  const char* const* manifest_names; /* n_manifest names, for catch1 messages */
} gwsem_job;

int gwsem_gwas_run(gwsem_model* m, gwsem_source* s, const gwsem_job* job, int64_t* rows_written);

#ifdef __cplusplus
}
#endif
#endif
