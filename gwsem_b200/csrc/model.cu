// Host side of a model: turns the flat RAM specification + phenotype columns into the
// device-resident FitProgram (feature matrix, moment index tables, RAM cells) and runs the
// per-batch kernel sequence.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include "common.h"
#include "fit_program.h"
#include "kernels.h"
#include "model.h"

namespace gwsem {

template <typename T>
static const T* upload_vec(std::vector<DeviceBuffer<uint8_t>>& pool, const std::vector<T>& v, cudaStream_t st) {
  pool.emplace_back();
  auto& b = pool.back();
  b.ensure(std::max<size_t>(v.size() * sizeof(T), 16));
  if (!v.empty()) GW_CUDA(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  return reinterpret_cast<const T*>(b.p);
}

}  // namespace gwsem

using namespace gwsem;

gwsem_model::gwsem_model(gwsem_engine* e, const gwsem_model_spec* s) : engine(e) {
  if (s->fit < 0 || s->fit > 2) fail("unknown fit function %d", s->fit);
  fit_kind = s->fit;
  nv = s->n_vars;
  p = s->n_manifest;
  P = s->n_params;
  n_rows = s->n_rows;
  min_variance = s->min_variance;
  if (p < 1 || nv < p || P < 1 || n_rows < 2) fail("degenerate model specification");
  cells.assign(s->cells, s->cells + 5 * (size_t)s->n_cells);
  par_start.assign(s->par_start, s->par_start + P);
  par_lbound.assign(s->par_lbound, s->par_lbound + P);
  kind.assign(s->manifest_kind, s->manifest_kind + p);
  // distinct SNP-independent base columns
  std::map<const double*, int> base_of;
  var_base.assign(p, -1);
  std::vector<bool> used_as_moderator;
  for (int v = 0; v < p; ++v) {
    if (kind[v] == 1) continue;
    const double* col = s->manifest_data[v];
    if (!col) fail("manifest variable %d has no data column", v);
    auto it = base_of.find(col);
    if (it == base_of.end()) {
      it = base_of.emplace(col, (int)base.size()).first;
      base.emplace_back(col, col + n_rows);
      used_as_moderator.push_back(false);
    }
    var_base[v] = it->second;
    if (kind[v] == 2) used_as_moderator[it->second] = true;
  }
  pm = p;
  if (fit_kind == 2) {
    // ML: the definition-variable covariates join the manifest variables as moment variables
    // (second-order moments of [1, covariates, manifest] are the sufficient statistics).  A covariate
    // without a data column is the genotype itself (buildItem's default exogenous = TRUE under ML).
    moment_order = 2;
    for (int v = 0; v < p; ++v)
      if (s->n_categories && s->n_categories[v] > 0) fail("ML with ordinal items is not implemented (continuous manifest variables only)");
    exo_latent.assign(s->exo_latent, s->exo_latent + s->n_exo);
    for (int k = 0; k < s->n_exo; ++k) {
      const double* col = s->exo_data[k];
      const int ek = s->exo_kind ? s->exo_kind[k] : (col ? 0 : 1);
      if (ek == 1) {
        if (snp_exo >= 0) fail("ML: more than one exogenous covariate without data");
        snp_exo = k;
        kind.push_back(1);
        var_base.push_back(-1);
      } else {
        if (!col) fail("ML: exogenous covariate %d has no data column", k);
        auto it = base_of.find(col);
        if (it == base_of.end()) {
          it = base_of.emplace(col, (int)base.size()).first;
          base.emplace_back(col, col + n_rows);
          used_as_moderator.push_back(false);
        }
        kind.push_back(ek == 2 ? 2 : 0);   // 2: genotype x this column (gxe product as definition variable)
        var_base.push_back(it->second);
        if (ek == 2) used_as_moderator[it->second] = true;
      }
    }
    pm = p + s->n_exo;
    for (int v = 0; v < p; ++v)
      if (snp_exo >= 0 && kind[v] != 0) fail("ML: snp cannot be both a manifest variable and an exogenous covariate");
  }
  // naAction='omit' on the SNP-independent columns is static: drop those people once
  row_ok.assign(n_rows, 1);
  for (auto& col : base)
    for (int r = 0; r < n_rows; ++r)
      if (!std::isfinite(col[r])) row_ok[r] = 0;
  if (fit_kind == 1) {
    // marg_mode 0: snp is the first manifest variable and nothing else carries the genotype (buildOneFac /
    //              buildOneFacRes / buildTwoFac): SNP-independent item side + marg_stats_kernel, and
    //              marg_exact_kernel for the SNPs with a missing call
    //           1: one ordinal item regressed on snp (+ gxe products) and covariates (probit.cu)
    //           2: one continuous item regressed on the same (linreg.cu)
    //           3: anything else (several items with an exogenous snp, gxe products as manifest variables, snp
    //              elsewhere in the manifest list): marg_exact_kernel for every SNP
    n_categories.assign(s->n_categories, s->n_categories + p);
    exo_latent.assign(s->exo_latent, s->exo_latent + s->n_exo);
    exo_free.assign(s->exo_free, s->exo_free + (size_t)p * s->n_exo);
    exo_kind.assign(s->n_exo, 0);
    bool geno_in_exo = false, geno_elsewhere = false;
    for (int v = 1; v < p; ++v)
      if (kind[v] != 0) geno_elsewhere = true;
    for (int k = 0; k < s->n_exo; ++k) {
      const int ek = s->exo_kind ? s->exo_kind[k] : (s->exo_data[k] ? 0 : 1);
      exo_kind[k] = ek;
      if (ek != 0) geno_in_exo = true;
      if (ek == 1) {  // the genotype column as an exogenous predictor (buildItem with exogenous = TRUE)
        if (snp_exo >= 0) fail("marginals: more than one exogenous covariate without data");
        snp_exo = k;
        exo_cols.emplace_back();
        continue;
      }
      if (!s->exo_data[k]) fail("marginals: exogenous covariate %d has no data column", k);
      exo_cols.emplace_back(s->exo_data[k], s->exo_data[k] + n_rows);
      for (int r = 0; r < n_rows; ++r)
        if (!std::isfinite(exo_cols[k][r])) row_ok[r] = 0;
    }
    for (int v = 0; v < p; ++v) {
      if (n_categories[v] == 1) fail("marginals: manifest variable %d has a single category", v);
      if (n_categories[v] > 0 && kind[v] != 0) fail("marginals: the genotype cannot be an ordinal variable (src/loader.cpp:395-397)");
    }
    bool std_free = true;   // every item regressed on every covariate, snp on none
    for (int v = 0; v < p; ++v)
      for (int k = 0; k < s->n_exo; ++k)
        if (exo_free[(size_t)v * s->n_exo + k] != (kind[v] == 0)) std_free = false;
    if (kind[0] == 1 && !geno_elsewhere && !geno_in_exo && std_free && p - 1 <= gwsem::kMargMaxItems && p >= 2) marg_mode = 0;
    else if (p == 1 && kind[0] == 0 && snp_exo >= 0 && std_free && s->n_exo <= gwsem::kLinregMaxX)
      marg_mode = n_categories[0] > 0 ? 1 : 2;
    else marg_mode = 3;
    if (marg_mode == 3) {
      if (p > gwsem::kExMaxVars) fail("marginals: at most %d manifest variables", gwsem::kExMaxVars);
      if (s->n_exo > gwsem::kExMaxCov) fail("marginals: at most %d exogenous covariates", gwsem::kExMaxCov);
      bool any = geno_in_exo;
      for (int v = 0; v < p; ++v) any = any || kind[v] != 0;
      if (!any) fail("marginals: the model does not use the genotype");
    }
    thresholds.assign(s->thresholds, s->thresholds + 4 * (size_t)s->n_thresholds);
    q = p * (p + 1) / 2;
    nf = 1;
    stat_i.clear();
    stat_j.clear();
    for (int j = 0; j < p; ++j)
      for (int i = j; i < p; ++i) { stat_i.push_back(i); stat_j.push_back(j); }
    idx4.assign(1, -1);
    feature_monos.assign(1, {});
    return;
  }
  // centre the pure phenotype columns (covariances are shift invariant; keeps moments well scaled);
  // not under ML, whose mean structure is in the data's own units
  for (size_t b = 0; b < base.size() && fit_kind == 0; ++b) {
    if (used_as_moderator[b]) continue;
    long double sum = 0;
    long cnt = 0;
    for (int r = 0; r < n_rows; ++r)
      if (row_ok[r]) { sum += base[b][r]; ++cnt; }
    const double mean = cnt ? (double)(sum / cnt) : 0.0;
    for (int r = 0; r < n_rows; ++r) base[b][r] -= mean;
  }
  build_program();
}

// Enumerate every multiset of <= moment_order moment variables (manifest variables, plus the
// covariates under ML); each is g^a times a monomial of base columns.  idx4 is indexed by the
// sorted multiset padded with pm (= "none") in base pm + 1.
void gwsem_model::build_program() {
  const int p1 = pm + 1;
  q = p * (p + 1) / 2;
  std::map<std::vector<int>, int> mono_id;  // monomial (sorted base columns) -> provisional id
  std::vector<std::vector<int>> monos;
  std::vector<bool> dense;
  auto intern = [&](const std::vector<int>& m) {
    auto it = mono_id.find(m);
    if (it != mono_id.end()) return it->second;
    mono_id.emplace(m, (int)monos.size());
    monos.push_back(m);
    dense.push_back(false);
    return (int)monos.size() - 1;
  };
  intern({});  // the constant
  struct Slot { int a, mono; };
  size_t n_slots = 1;
  for (int k = 0; k < moment_order; ++k) n_slots *= p1;
  std::vector<Slot> slot(n_slots, Slot{-1, -1});
  std::vector<int> v(moment_order, 0);
  // odometer over non-decreasing tuples
  while (true) {
    if (v[0] != pm) {
      int a = 0;
      std::vector<int> m;
      size_t at = 0;
      for (int k = 0; k < moment_order; ++k) {
        at = at * p1 + v[k];
        if (v[k] == pm) continue;
        if (kind[v[k]] != 0) ++a;
        if (var_base[v[k]] >= 0) m.push_back(var_base[v[k]]);
      }
      std::sort(m.begin(), m.end());
      const int id = intern(m);
      if (a >= 1) dense[id] = true;
      slot[at] = Slot{a, id};
    }
    int k = moment_order - 1;
    while (k >= 0 && v[k] == pm) --k;
    if (k < 0) break;
    const int nv2 = v[k] + 1;
    for (int j = k; j < moment_order; ++j) v[j] = nv2;
  }
  // final order: constant, dense monomials, the rest
  std::vector<int> order(monos.size()), rank(monos.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
  std::stable_sort(order.begin() + 1, order.end(), [&](int x, int y) { return dense[x] > dense[y]; });
  for (size_t i = 0; i < order.size(); ++i) rank[order[i]] = (int)i;
  nf = (int)monos.size();
  nfd = 0;
  for (size_t i = 1; i < monos.size(); ++i) nfd += dense[i];
  nfm = nf - 1;
  feature_monos.resize(nf);
  for (int i = 0; i < nf; ++i) feature_monos[rank[i]] = monos[i];
  idx4.assign(slot.size(), -1);
  for (size_t i = 0; i < slot.size(); ++i)
    if (slot[i].a >= 0) idx4[i] = slot[i].a * nf + rank[slot[i].mono];
  stat_i.clear();
  stat_j.clear();
  for (int j = 0; j < p; ++j)
    for (int i = j; i < p; ++i) { stat_i.push_back(i); stat_j.push_back(j); }
}

// Lay the people axis out in genotype-file order and upload everything.
void gwsem_model::prepare(const int32_t* row_filter, int src_rows) {
  cudaStream_t st = engine->stream;
  GW_CUDA(cudaSetDevice(engine->device));
  pool.clear();
  int dest = 0;
  std::vector<int> dest_of(src_rows, -1);
  for (int s = 0; s < src_rows; ++s)
    if (!row_filter || !row_filter[s]) dest_of[s] = dest++;
  if (dest != n_rows)
    fail("rowFilter keeps %d of %d people, which does not match the number of rows of observed data (%d)", dest,
         src_rows, n_rows);
  n_src = src_rows;
  n_pad = (src_rows + 31) / 32 * 32;
  std::vector<uint8_t> inc8(n_pad, 0);
  std::vector<unsigned long long> incmask(n_pad / 32, 0), rowmask(n_pad / 32, 0);
  n_incl = 0;
  for (int s = 0; s < src_rows; ++s) {
    if (dest_of[s] < 0) continue;
    rowmask[s >> 5] |= 1ull << (2 * (s & 31));
    if (row_ok[dest_of[s]]) {
      inc8[s] = 3;
      incmask[s >> 5] |= 1ull << (2 * (s & 31));
      ++n_incl;
    }
  }
  // features
  const int nfm_pad = std::max(1, (nfm + 3) / 4 * 4);
  std::vector<double> feat((size_t)std::max(nfd, 1) * n_pad, 0.0), featT((size_t)n_pad * nfm_pad, 0.0), tot(nf, 0.0);
  std::vector<long double> acc(nf, 0.0L);
  // Features that need class sums are rounded to a fixed-point grid 2^-s_f (|h| 2^s_f <= 2^30) and the
  // rounded values are what every kernel sees: the class sums then are exact integer arithmetic on
  // the tensor cores (stats_imma.cu) and agree with the totals and the missing-class sums bit for bit.
  std::vector<double> scale(std::max(nfd, 1), 1.0), inv_scale(std::max(nfd, 1), 1.0);
  for (int f = 1; f <= nfd; ++f) {
    double mx = 0.0;
    for (int s = 0; s < src_rows; ++s) {
      if (!inc8[s]) continue;
      double h = 1.0;
      for (int b : feature_monos[f]) h *= base[b][dest_of[s]];
      mx = std::max(mx, std::fabs(h));
    }
    int ex = 0;
    if (mx > 0.0) { std::frexp(mx, &ex); }   // mx < 2^ex
    scale[f - 1] = std::ldexp(1.0, 30 - ex);
    inv_scale[f - 1] = std::ldexp(1.0, ex - 30);
  }
  std::vector<std::vector<int64_t>> fq(nfd, std::vector<int64_t>(n_pad, 0));
  for (int s = 0; s < src_rows; ++s) {
    if (!inc8[s]) continue;
    const int r = dest_of[s];
    for (int f = 1; f < nf; ++f) {
      double h = 1.0;
      for (int b : feature_monos[f]) h *= base[b][r];
      if (f - 1 < nfd) {
        const int64_t qv = std::llrint(h * scale[f - 1]);
        fq[f - 1][s] = qv;
        h = (double)qv * inv_scale[f - 1];
        feat[(size_t)(f - 1) * n_pad + s] = h;
      }
      featT[(size_t)s * nfm_pad + (f - 1)] = h;
      acc[f] += h;
    }
  }
  tot[0] = n_incl;
  for (int f = 1; f < nf; ++f) tot[f] = (double)acc[f];
  this->nfm_pad = nfm_pad;

  // RAM cells, grouped by parameter
  const int n_cells = (int)cells.size() / 5;
  std::vector<int32_t> cm, cr, cc, cp;
  std::vector<double> cv;
  std::vector<std::vector<int>> by_par(P);
  for (int c = 0; c < n_cells; ++c) {
    const double* x = &cells[5 * (size_t)c];
    if ((int)x[0] == 2 && fit_kind == 0) continue;  // means: no mean structure under cumulants
    const int id = (int)cm.size();
    cm.push_back((int)x[0]); cr.push_back((int)x[1]); cc.push_back((int)x[2]); cp.push_back((int)x[3]); cv.push_back(x[4]);
    if ((int)x[3] >= P) fail("cell refers to parameter %d of %d", (int)x[3], P);
    if ((int)x[3] >= 0) by_par[(int)x[3]].push_back(id);
  }
  std::vector<int32_t> ptr(P + 1, 0), flat;
  for (int k = 0; k < P; ++k) {
    ptr[k] = (int)flat.size();
    for (int id : by_par[k]) flat.push_back(id);
  }
  ptr[P] = (int)flat.size();

  fp = FitProgram();
  fp.p = p; fp.q = q; fp.nv = nv; fp.P = P; fp.n_cells = (int)cm.size();
  fp.nf = nf; fp.nfd = nfd; fp.nfm = nfm; fp.n_incl = n_incl; fp.min_variance = min_variance;
  fp.tot = upload_vec(pool, tot, st);
  fp.idx4 = upload_vec(pool, idx4, st);
  fp.stat_i = upload_vec(pool, stat_i, st);
  fp.stat_j = upload_vec(pool, stat_j, st);
  fp.cell_mat = upload_vec(pool, cm, st);
  fp.cell_row = upload_vec(pool, cr, st);
  fp.cell_col = upload_vec(pool, cc, st);
  fp.cell_par = upload_vec(pool, cp, st);
  fp.cell_val = upload_vec(pool, cv, st);
  fp.par_cell_ptr = upload_vec(pool, ptr, st);
  fp.par_cells = upload_vec(pool, flat, st);
  fp.par_start = upload_vec(pool, par_start, st);
  fp.par_lbound = upload_vec(pool, par_lbound, st);
  layout.n_src = n_src;
  layout.n_pad = n_pad;
  layout.n_incl = n_incl;
  layout.d_inc8 = upload_vec(pool, inc8, st);
  layout.d_incmask = upload_vec(pool, incmask, st);
  layout.d_rowmask = upload_vec(pool, rowmask, st);
  layout.n_kept = n_rows;
  d_feat = upload_vec(pool, feat, st);
  if (nfd > 0) {
    std::vector<int8_t> packed_q;
    std::vector<uint8_t> kept8(n_pad, 0);
    for (int s2 = 0; s2 < src_rows; ++s2) kept8[s2] = dest_of[s2] >= 0;
    use_tc5 = cs_use_tc5(nfd);
    if (use_tc5) tc5_pack_features(fq, n_pad, inc8, kept8, packed_q);
    else imma_pack_features(fq, n_pad, inc8, kept8, packed_q);
    d_featq = upload_vec(pool, packed_q, st);
    d_inv_scale = upload_vec(pool, inv_scale, st);
  }
  d_featT = upload_vec(pool, featT, st);
  d_dest_of = upload_vec(pool, dest_of, st);
  if (fit_kind == 1) prepare_marginals(dest_of);
  GW_CUDA(cudaStreamSynchronize(st));
  prepared = true;
  if (fit_kind == 2) prepare_ml();
}

// ML: covariate bookkeeping for the kernel, then the per-model warm start.
void gwsem_model::prepare_ml() {
  cudaStream_t st = engine->stream;
  std::vector<int32_t> man_static(p);
  for (int i = 0; i < p; ++i) man_static[i] = kind[i] == 0;
  fp.n_exo = (int)exo_latent.size();
  fp.exo_latent = upload_vec(pool, exo_latent, st);
  fp.man_static = upload_vec(pool, man_static, st);
  fp.snp_exo = snp_exo >= 0;
  fp.snp_exo_index = std::max(snp_exo, 0);
  prepare_warm();
}

// Per-model warm start (see prepare_marginals): the estimates for a reference genotype column
// (codes 0,1,2,0,1,2,... down the file), fitted once from the model's own start values.  Every SNP
// starts its iteration there, so a row still does not depend on its neighbours (R/model.R:188).
void gwsem_model::prepare_warm() {
  cudaStream_t st = engine->stream;
  fp.par_warm = nullptr;
  const int64_t rb = (n_pad / 4 + 15) / 16 * 16;
  std::vector<uint8_t> ref(rb, 0);
  for (int i = 0; i < n_src; ++i) ref[i >> 2] |= (uint8_t)((i % 3) << (2 * (i & 3)));
  const uint8_t* d_ref = upload_vec(pool, ref, st);
  DeviceBuffer<double> dd;
  DeviceBuffer<int32_t> di;
  dd.ensure((size_t)P + (size_t)P * P + 2);
  di.ensure(5);
  gwsem_result r;
  r.est = dd.p; r.vcov = dd.p + P; r.fit = dd.p + P + (size_t)P * P; r.maf = r.fit + 1;
  r.n_used = di.p; r.status = di.p + 1; r.catch_code = di.p + 2; r.catch_var = di.p + 3; r.iterations = di.p + 4;
  fit_2bit_device(d_ref, rb, 1, 0, r);
  int32_t flags[5];
  std::vector<double> warm(P);
  GW_CUDA(cudaMemcpyAsync(flags, di.p, sizeof(flags), cudaMemcpyDeviceToHost, st));
  GW_CUDA(cudaMemcpyAsync(warm.data(), dd.p, P * sizeof(double), cudaMemcpyDeviceToHost, st));
  GW_CUDA(cudaStreamSynchronize(st));
  if (flags[1] == GWSEM_STATUS_OK && flags[2] == GWSEM_CATCH_NONE) fp.par_warm = upload_vec(pool, warm, st);
  GW_CUDA(cudaStreamSynchronize(st));
}

void gwsem_model::fit_2bit_device(const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1,
                                  const gwsem_result& d_res) {
  if (!prepared) prepare(nullptr, n_rows);
  GW_CUDA(cudaSetDevice(engine->device));
  const int kMaxBatch = 1 << 16;
  for (int s0 = 0; s0 < n_snps; s0 += kMaxBatch) {
    const int n = std::min(kMaxBatch, n_snps - s0);
    d_counts.ensure(8 * (size_t)n);
    d_sums.ensure(2 * (size_t)std::max(nfd, 1) * n);
    d_miss.ensure((size_t)std::max(nfm, 1) * n);
    d_R.ensure(5 * (size_t)nf * n);
    d_nrows.ensure(n);
    d_mafs.ensure(n);
    const uint8_t* rows = d_packed + (int64_t)s0 * pitch;
    if (fit_kind == 1) {
      gwsem_result r = d_res;
      r.est += (size_t)s0 * P; r.vcov += (size_t)s0 * P * P; r.fit += s0; r.maf += s0; r.n_used += s0;
      r.status += s0; r.catch_code += s0; r.catch_var += s0; r.iterations += s0;
      run_marginals(rows, pitch, nullptr, nullptr, n, plink1, r);
      continue;
    }
    static const bool dfma_first = getenv("GWSEM_CS_DFMA") != nullptr;
    const bool light = !dfma_first && imma_has_counts(nfd);
    static const bool dfma = getenv("GWSEM_CS_DFMA") != nullptr;  // A/B aid: the FP64 class-sum kernel
    if (light) {
      // the class-sum kernel lists the SNPs that have a missing call; count_missing then reads only those rows
      // (none in imputed data) instead of streaming the whole batch a second time
      d_miss_list.ensure((size_t)n + 1);
      cudaStream_t st = engine->stream;
      GW_CUDA(cudaMemsetAsync(d_miss_list.p + n, 0, sizeof(int32_t), st));
      GW_CUDA(cudaMemsetAsync(d_counts.p, 0, 8 * (size_t)n * sizeof(int32_t), st));
      GW_CUDA(cudaMemsetAsync(d_miss.p, 0, (size_t)std::max(nfm, 1) * n * sizeof(double), st));
      if (use_tc5)
        launch_class_sums_tc5(engine, rows, pitch, n, plink1, layout, d_featq, d_inv_scale, nfd, d_sums.p, d_counts.p,
                              d_miss_list.p, d_miss_list.p + n);
      else
        launch_class_sums_imma(engine, rows, pitch, n, plink1, layout, d_featq, d_inv_scale, nfd, d_sums.p, d_counts.p,
                               d_miss_list.p, d_miss_list.p + n);
      launch_count_missing(engine, rows, pitch, n, plink1, layout, d_featT, nfm, nfm_pad, d_counts.p, d_miss.p, true,
                           d_miss_list.p, d_miss_list.p + n);
    } else {
      launch_count_missing(engine, rows, pitch, n, plink1, layout, d_featT, nfm, nfm_pad, d_counts.p, d_miss.p, false);
      if (dfma) launch_class_sums(engine, rows, pitch, n, plink1, layout, d_feat, nfd, d_sums.p);
      else if (use_tc5) launch_class_sums_tc5(engine, rows, pitch, n, plink1, layout, d_featq, d_inv_scale, nfd, d_sums.p, d_counts.p);
      else launch_class_sums_imma(engine, rows, pitch, n, plink1, layout, d_featq, d_inv_scale, nfd, d_sums.p, d_counts.p);
    }
    gwsem_result r = d_res;
    r.est += (size_t)s0 * P; r.vcov += (size_t)s0 * P * P; r.fit += s0; r.maf += s0; r.n_used += s0;
    r.status += s0; r.catch_code += s0; r.catch_var += s0; r.iterations += s0;
    launch_assemble_moments(engine, fp, layout, n, d_counts.p, d_sums.p, d_miss.p, d_R.p, d_nrows.p, d_mafs.p);
    if (fit_kind == 2) launch_fit_ml(engine, fp, n, d_R.p, d_nrows.p, d_mafs.p, r);
    else launch_fit_cumulants(engine, fp, n, d_R.p, d_nrows.p, d_mafs.p, r);
  }
}

// Generic path: genotype columns as doubles in file order (n_pad apart), NaN = missing.
void gwsem_model::fit_rows_device(const double* d_geno, const double* d_maf, int n_snps, const gwsem_result& d_res) {
  if (!prepared) prepare(nullptr, n_rows);
  GW_CUDA(cudaSetDevice(engine->device));
  const int kMaxBatch = 1 << 14;
  for (int s0 = 0; s0 < n_snps; s0 += kMaxBatch) {
    const int n = std::min(kMaxBatch, n_snps - s0);
    if (fit_kind == 1) {
      gwsem_result r = d_res;
      r.est += (size_t)s0 * P; r.vcov += (size_t)s0 * P * P; r.fit += s0; r.maf += s0; r.n_used += s0;
      r.status += s0; r.catch_code += s0; r.catch_var += s0; r.iterations += s0;
      run_marginals(nullptr, 0, d_geno + (size_t)s0 * n_pad, d_maf + s0, n, 0, r);
      continue;
    }
    d_R.ensure(5 * (size_t)nf * n);
    d_nrows.ensure(n);
    launch_power_sums(engine, fp, layout, n, d_geno + (size_t)s0 * n_pad, d_featT, nfm_pad, d_R.p, d_nrows.p, moment_order);
    gwsem_result r = d_res;
    r.est += (size_t)s0 * P; r.vcov += (size_t)s0 * P * P; r.fit += s0; r.maf += s0; r.n_used += s0;
    r.status += s0; r.catch_code += s0; r.catch_var += s0; r.iterations += s0;
    if (fit_kind == 2) launch_fit_ml(engine, fp, n, d_R.p, d_nrows.p, d_maf + s0, r);
    else launch_fit_cumulants(engine, fp, n, d_R.p, d_nrows.p, d_maf + s0, r);
  }
}

void gwsem_model::ensure_result_buffers(int n) {
  r_est.ensure((size_t)n * P); r_vcov.ensure((size_t)n * P * P); r_fit.ensure(n); r_maf.ensure(n);
  r_n.ensure(n); r_status.ensure(n); r_catch.ensure(n); r_cvar.ensure(n); r_iter.ensure(n);
}

gwsem_result gwsem_model::device_result() {
  gwsem_result r;
  r.est = r_est.p; r.vcov = r_vcov.p; r.fit = r_fit.p; r.maf = r_maf.p; r.n_used = r_n.p;
  r.status = r_status.p; r.catch_code = r_catch.p; r.catch_var = r_cvar.p; r.iterations = r_iter.p;
  return r;
}

void gwsem_model::download_result(const gwsem_result& h, const gwsem_result& d, int n, int offset, cudaStream_t on) {
  cudaStream_t st = on ? on : engine->stream;
  auto dl = [&](void* dst, const void* src, size_t bytes) {
    if (dst) GW_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
  };
  dl(h.est ? h.est + (size_t)offset * P : nullptr, d.est, sizeof(double) * n * P);
  dl(h.vcov ? h.vcov + (size_t)offset * P * P : nullptr, d.vcov, sizeof(double) * n * P * P);
  dl(h.fit ? h.fit + offset : nullptr, d.fit, sizeof(double) * n);
  dl(h.maf ? h.maf + offset : nullptr, d.maf, sizeof(double) * n);
  dl(h.n_used ? h.n_used + offset : nullptr, d.n_used, sizeof(int32_t) * n);
  dl(h.status ? h.status + offset : nullptr, d.status, sizeof(int32_t) * n);
  dl(h.catch_code ? h.catch_code + offset : nullptr, d.catch_code, sizeof(int32_t) * n);
  dl(h.catch_var ? h.catch_var + offset : nullptr, d.catch_var, sizeof(int32_t) * n);
  dl(h.iterations ? h.iterations + offset : nullptr, d.iterations, sizeof(int32_t) * n);
}

// Host buffers in, host results out.  The packed rows are cut into chunks that alternate between
// two HBM staging buffers: the H2D copy of chunk c+1 (copy stream) overlaps the kernels of chunk c
// (engine stream), and the results of chunk c come back on a third stream while chunk c+1 computes.  A chunk is a
// whole number of waves of the 128-SNP class-sum CTAs over the SMs (18,944 SNPs on 148 SMs) when the rows are short
// enough for two such staging buffers.
void gwsem_model::fit_2bit_host(const uint8_t* packed, int64_t pitch, int n_snps, int plink1, const gwsem_result& h) {
  if (!prepared) prepare(nullptr, n_rows);
  GW_CUDA(cudaSetDevice(engine->device));
  cudaStream_t st = engine->stream;
  const int64_t row_bytes = (n_src + 3) / 4;
  if (pitch < row_bytes) fail("pitch %lld is smaller than a %d-person record (%lld bytes)", (long long)pitch, n_src, (long long)row_bytes);
  const int64_t dpitch = (std::max<int64_t>(row_bytes, n_pad / 4) + 15) / 16 * 16;
  if (!copy_stream) {
    GW_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    GW_CUDA(cudaStreamCreateWithFlags(&out_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
      GW_CUDA(cudaEventCreateWithFlags(&ev_copied[b], cudaEventDisableTiming));
      GW_CUDA(cudaEventCreateWithFlags(&ev_done[b], cudaEventDisableTiming));
      GW_CUDA(cudaEventCreateWithFlags(&ev_fitted[b], cudaEventDisableTiming));
      GW_CUDA(cudaEventCreateWithFlags(&ev_out[b], cudaEventDisableTiming));
    }
  }
  // ~128 MiB of packed rows per chunk, rounded to whole waves of 128-SNP CTAs (up to 1 GiB per staging buffer)
  int chunk = (int)std::max<int64_t>(256, std::min<int64_t>(1 << 16, (128ll << 20) / dpitch));
  {
    const int64_t wave = 128ll * engine->sm_count;
    if (wave * dpitch <= (1ll << 30)) chunk = (int)(std::max<int64_t>(1, (chunk + wave / 2) / wave) * wave);
  }
  ensure_result_buffers(std::min(chunk, n_snps) * 2);
  for (int b = 0; b < 2; ++b) {
    d_stage2[b].ensure((size_t)std::min(chunk, n_snps) * dpitch + 16);
    if (dpitch != row_bytes) GW_CUDA(cudaMemsetAsync(d_stage2[b].p, 0, d_stage2[b].n, st));
  }
  GW_CUDA(cudaEventRecord(ev_done[0], st));
  GW_CUDA(cudaEventRecord(ev_done[1], st));
  GW_CUDA(cudaEventRecord(ev_out[0], st));
  GW_CUDA(cudaEventRecord(ev_out[1], st));
  int c = 0;
  for (int s0 = 0; s0 < n_snps; s0 += chunk, ++c) {
    const int b = c & 1;
    const int n = std::min(chunk, n_snps - s0);
    GW_CUDA(cudaStreamWaitEvent(copy_stream, ev_done[b], 0));  // staging buffer b is free again
    if (dpitch == pitch)
      GW_CUDA(cudaMemcpyAsync(d_stage2[b].p, packed + (int64_t)s0 * pitch, (size_t)n * pitch, cudaMemcpyHostToDevice, copy_stream));
    else
      GW_CUDA(cudaMemcpy2DAsync(d_stage2[b].p, dpitch, packed + (int64_t)s0 * pitch, pitch, row_bytes, n,
                                cudaMemcpyHostToDevice, copy_stream));
    GW_CUDA(cudaEventRecord(ev_copied[b], copy_stream));
    GW_CUDA(cudaStreamWaitEvent(st, ev_copied[b], 0));
    gwsem_result r = device_result();
    const size_t o = (size_t)b * std::min(chunk, n_snps);
    r.est += o * P; r.vcov += o * P * P; r.fit += o; r.maf += o; r.n_used += o;
    r.status += o; r.catch_code += o; r.catch_var += o; r.iterations += o;
    GW_CUDA(cudaStreamWaitEvent(st, ev_out[b], 0));   // the results that lived in this half two chunks ago are on the host
    fit_2bit_device(d_stage2[b].p, dpitch, n, plink1, r);
    GW_CUDA(cudaEventRecord(ev_done[b], st));
    GW_CUDA(cudaEventRecord(ev_fitted[b], st));
    GW_CUDA(cudaStreamWaitEvent(out_stream, ev_fitted[b], 0));
    download_result(h, r, n, s0, out_stream);
    GW_CUDA(cudaEventRecord(ev_out[b], out_stream));
  }
  GW_CUDA(cudaStreamSynchronize(st));
  GW_CUDA(cudaStreamSynchronize(out_stream));
}
