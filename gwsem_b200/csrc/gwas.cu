// Genotype sources on the C ABI, batch staging (file -> pinned host -> HBM -> canonical layout)
// and the whole per-SNP loop with its checkpoint log: the native equivalent of
//   mxComputeLoop(SetOriginalStarts, LoadData, LoadContext, TryCatch(GD, SE), Checkpoint)
// assembled by prepareComputePlan (reference R/model.R:187-194).
#include <sys/time.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <ctime>
#include <fstream>
#include <map>
#include <sstream>
#include <charconv>
#include <chrono>
#include <future>
#include <thread>

#include "common.h"
#include "kernels.h"
#include "model.h"
#include "records.h"
#include "source.h"

namespace gwsem {

template <typename T>
struct PinnedBuffer {
  T* p = nullptr;
  size_t n = 0;
  ~PinnedBuffer() { if (p) cudaFreeHost(p); }
  void ensure(size_t count) {
    if (count <= n) return;
    if (p) cudaFreeHost(p);
    p = nullptr;
    GW_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&p), count * sizeof(T), cudaHostAllocDefault));
    n = count;
  }
};

// Brings a batch of variants of one source into HBM in the engine's canonical layout.
struct Stager {
  gwsem_engine* e;
  gwsem_source* s;
  int n_pad;
  int64_t pitch;  // canonical row pitch (bytes, multiple of 16, >= n_pad / 4)
  PinnedBuffer<uint8_t> h_bytes;
  PinnedBuffer<uint8_t> h_bgen[2];   // inflated bgen blocks, read ahead like h_fixed
  std::vector<uint64_t> bgen_off[2];
  struct Borrowed { uint8_t* p = nullptr; } h_fixed[2];  // fixed-width records: read ahead by a host thread (see read_fixed); the engine owns the memory
  DeviceBuffer<uint8_t> d_bytes, d_vt, d_rows;
  DeviceBuffer<uint64_t> d_off;
  DeviceBuffer<int32_t> d_base, d_err;
  DeviceBuffer<uint16_t> d_dosage;
  std::vector<uint64_t> off;
  std::vector<uint8_t> vt;
  std::vector<int32_t> base_row, h_err;
  std::vector<uint8_t> scratch, inflated;
  // outputs of stage_pgen
  bool plink1 = false, has_dosage = false, rows_zeroed = false;
  int n_rows_staged = 0;

  Stager(gwsem_engine* e_, gwsem_source* s_) : e(e_), s(s_) {
    n_pad = ((int)s->n_samples + 31) / 32 * 32;
    const int64_t row_bytes = ((int64_t)s->n_samples + 3) / 4;
    pitch = (std::max<int64_t>(row_bytes, n_pad / 4) + 15) / 16 * 16;
  }

  void check_index(uint32_t v) const {
    if (v >= s->n_variants)  // loader.cpp:369-372
      fail("%s: out of data (record %u requested but only %u in file)", s->kind == gwsem_source::kPgen ? "pgen" : "bgen", v + 1,
           s->n_variants);
  }

  bool fixed_width() const { return s->kind == gwsem_source::kPgen && (s->mode == 0x01 || s->mode == 0x02); }

  // Fixed-width hardcall records (.bed, pgen mode 0x02): the record already is the canonical row.
  // Reads variants[0..n) into pinned slot `slot`; runs of consecutive variants are one request,
  // split over a few threads (page-cache copies scale with threads).  Host work only: runs on a
  // helper thread while the device works on the previous batch.
  void read_fixed(const uint32_t* variants, int n, int slot) {
    for (int k = 0; k < n; ++k) check_index(variants[k]);
    const int64_t row_bytes = ((int64_t)s->n_samples + 3) / 4;
    struct Run { uint64_t off; uint8_t* dst; size_t len; };
    std::vector<Run> runs;
    const size_t piece = 8u << 20;
    int k = 0;
    while (k < n) {
      int j = k + 1;
      while (j < n && variants[j] == variants[j - 1] + 1) ++j;
      uint64_t off0 = s->record_offset(variants[k]);
      uint8_t* dst = h_fixed[slot].p + (size_t)k * row_bytes;
      size_t len = (size_t)(j - k) * row_bytes;
      while (len) {
        const size_t take = std::min(len, piece);
        runs.push_back({off0, dst, take});
        off0 += take; dst += take; len -= take;
      }
      k = j;
    }
    const int threads = (int)std::min<size_t>(std::min(8u, std::max(1u, std::thread::hardware_concurrency())), runs.size());
    std::vector<std::exception_ptr> err(std::max(threads, 1));
    auto work = [&](int t) {
      try {
        for (size_t r = t; r < runs.size(); r += threads) s->read_at(runs[r].off, runs[r].dst, runs[r].len);
      } catch (...) { err[t] = std::current_exception(); }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    if (threads > 0) work(0);
    for (auto& th : pool) th.join();
    for (auto& e2 : err) if (e2) std::rethrow_exception(e2);
  }
  void ensure_fixed(int max_rows) {  // on the calling thread: pinned allocation is a CUDA call
    const int64_t row_bytes = ((int64_t)s->n_samples + 3) / 4;
    for (int b = 0; b < 2; ++b) {
      const size_t need = (size_t)max_rows * row_bytes + 16;
      if (e->pinned_stage_bytes[b] < need) {
        if (e->pinned_stage[b]) cudaFreeHost(e->pinned_stage[b]);
        e->pinned_stage[b] = nullptr;
        e->pinned_stage_bytes[b] = 0;
        GW_CUDA(cudaHostAlloc(&e->pinned_stage[b], need, cudaHostAllocDefault));
        e->pinned_stage_bytes[b] = need;
      }
      h_fixed[b].p = static_cast<uint8_t*>(e->pinned_stage[b]);
    }
  }
  void upload_fixed(int n, int slot) {
    cudaStream_t st = e->stream;
    const int64_t row_bytes = ((int64_t)s->n_samples + 3) / 4;
    plink1 = s->mode == 0x01;
    has_dosage = false;
    d_rows.ensure((size_t)n * pitch + 16);
    if (!rows_zeroed) {
      GW_CUDA(cudaMemsetAsync(d_rows.p, 0, d_rows.n, st));
      rows_zeroed = true;
    }
    GW_CUDA(cudaMemcpy2DAsync(d_rows.p, pitch, h_fixed[slot].p, row_bytes, row_bytes, n, cudaMemcpyHostToDevice, st));
    n_rows_staged = n;
  }

  // PGEN / BED: leaves variants[k] decoded in row k of d_rows (+ d_dosage when has_dosage)
  void stage_pgen(const uint32_t* variants, int n) {
    cudaStream_t st = e->stream;
    for (int k = 0; k < n; ++k) check_index(variants[k]);
    plink1 = s->mode == 0x01;
    has_dosage = false;
    const int64_t row_bytes = ((int64_t)s->n_samples + 3) / 4;
    if (s->mode == 0x01 || s->mode == 0x02) {
      // fixed-width hardcalls: the record already is the canonical row; runs of consecutive variants
      // are read with one pread and re-pitched by the copy engine
      d_rows.ensure((size_t)n * pitch + 16);
      if (!rows_zeroed) {
        GW_CUDA(cudaMemsetAsync(d_rows.p, 0, d_rows.n, st));
        rows_zeroed = true;
      }
      h_bytes.ensure((size_t)n * row_bytes + 16);
      int k = 0;
      while (k < n) {
        int j = k + 1;
        while (j < n && variants[j] == variants[j - 1] + 1) ++j;
        s->read_at(s->record_offset(variants[k]), h_bytes.p + (size_t)k * row_bytes, (size_t)(j - k) * row_bytes);
        k = j;
      }
      GW_CUDA(cudaMemcpy2DAsync(d_rows.p, pitch, h_bytes.p, row_bytes, row_bytes, n, cudaMemcpyHostToDevice, st));
      n_rows_staged = n;
      return;
    }
    // encoded records: requested variants first, then the LD bases they need
    std::vector<uint32_t> recs(variants, variants + n);
    std::map<uint32_t, int> row_of_base;
    base_row.assign(n, -1);
    bool any_ld = false;
    for (int k = 0; k < n; ++k) {
      const uint8_t t = s->record_vrtype(variants[k]);
      if ((t & 6) == 2) {
        any_ld = true;
        const uint32_t b = s->ld_base(variants[k]);
        auto it = row_of_base.find(b);
        if (it == row_of_base.end()) {
          it = row_of_base.emplace(b, (int)recs.size()).first;
          recs.push_back(b);
        }
        base_row[k] = it->second;
      }
    }
    const int total = (int)recs.size();
    base_row.resize(total, -1);
    off.assign(total + 1, 0);
    vt.assign(total, 0);
    uint64_t bytes = 0;
    for (int k = 0; k < total; ++k) {
      off[k] = bytes;
      bytes += s->record_length(recs[k]);
      vt[k] = s->record_vrtype(recs[k]);
      if (k < n && (vt[k] & 0x60)) has_dosage = true;
    }
    off[total] = bytes;
    h_bytes.ensure(bytes + 16);
    for (int k = 0; k < total;) {  // coalesce reads of adjacent records
      int j = k + 1;
      while (j < total && recs[j] == recs[j - 1] + 1) ++j;
      s->read_at(s->record_offset(recs[k]), h_bytes.p + off[k], (size_t)(off[j] - off[k]));
      k = j;
    }
    memset(h_bytes.p + bytes, 0, 16);
    d_bytes.upload(h_bytes.p, bytes + 16, st);
    d_off.upload(off.data(), off.size(), st);
    d_vt.upload(vt.data(), vt.size(), st);
    d_base.upload(base_row.data(), base_row.size(), st);
    d_rows.ensure((size_t)total * pitch + 16);
    rows_zeroed = false;
    d_err.ensure(total);
    GW_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int32_t) * total, st));
    if (has_dosage) d_dosage.ensure((size_t)total * n_pad);
    RecordBatch rb;
    rb.bytes = d_bytes.p; rb.off = d_off.p; rb.vrtype = d_vt.p; rb.base_row = d_base.p;
    rb.n = total; rb.n_samples = (int)s->n_samples; rb.n_pad = n_pad; rb.plink1 = 0;
    rb.rows = d_rows.p; rb.pitch = pitch; rb.dosage = has_dosage ? d_dosage.p : nullptr; rb.err = d_err.p;
    launch_pgen_records(e, rb, any_ld);
    h_err.resize(total);
    GW_CUDA(cudaMemcpyAsync(h_err.data(), d_err.p, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, st));
    GW_CUDA(cudaStreamSynchronize(st));  // pinned staging is reused by the next batch
    for (int k = 0; k < n; ++k)
      if (h_err[k])  // loader.cpp:380-381
        fail("pgen: read_dosages(varient %u) error code %d", variants[k], h_err[k]);
    n_rows_staged = n;
  }

  // BGEN: inflates the probability blocks of variants[] (index order) and uploads them
  // host half: inflated probability blocks of variants[] (index order) into pinned slot `slot`
  void read_bgen(const uint32_t* variants, int n, int slot) {
    std::vector<std::pair<uint64_t, uint32_t>> payload(n);
    std::vector<uint64_t> starts(n + 1), lens(n);
    uint64_t cur = 0;
    for (int k = 0; k < n; ++k) {
      if (variants[k] >= s->n_variants) fail("bgen: %s has no more varients", s->path.c_str());  // loader.cpp:122-124
      s->parse_bgen_variant(variants[k]);
      const auto& v = s->bgen[variants[k]];
      uint32_t hdr[2] = {0, 0};
      uint32_t ulen, plen;
      uint64_t pstart;
      if (s->bgen_layout == 1) {  // src/bgen.cpp:341-386
        ulen = 6 * s->n_samples;
        if (s->bgen_compression) {
          s->read_at(v.geno, hdr, 4);
          plen = hdr[0];
          pstart = v.geno + 4;
        } else {
          plen = ulen;
          pstart = v.geno;
        }
      } else {
        s->read_at(v.geno, hdr, s->bgen_compression ? 8 : 4);
        if (s->bgen_compression) { plen = hdr[0] - 4; ulen = hdr[1]; pstart = v.geno + 8; }
        else { plen = ulen = hdr[0]; pstart = v.geno + 4; }
      }
      if (pstart + plen > s->file_size) fail("%s: genotype block runs past the end of the file", s->path.c_str());
      payload[k] = {pstart, plen};
      starts[k] = cur;
      lens[k] = ulen;
      cur += ulen;
    }
    starts[n] = cur;
    if (cur + 16 > h_bgen[slot].n) fail("bgen: a batch of %d variants needs %llu bytes of staging, more than reserved", n, (unsigned long long)cur);
    uint8_t* dst = h_bgen[slot].p;
    {
      // blocks are independent: read + inflate them on a few host threads (zlib / zstd are the cost)
      if (s->bgen_compression == 2) {  // resolve the zstd entry point before the threads start
        std::vector<uint8_t> none;
        try { bgen_inflate(*s, nullptr, 0, none); } catch (const std::exception&) {}
      }
      const int threads = (int)std::min<unsigned>(std::min(64u, std::max(1u, std::thread::hardware_concurrency())), (unsigned)n);   // every host core: inflate bounds bgen end to end
      std::vector<std::exception_ptr> err(std::max(threads, 1));
      auto work = [&](int t) {
        try {
          std::vector<uint8_t> in, outv;
          for (int k = t; k < n; k += threads) {
            in.resize(payload[k].second + 8);
            s->read_at(payload[k].first, in.data(), payload[k].second);
            outv.resize(lens[k]);
            bgen_inflate(*s, in.data(), payload[k].second, outv);
            memcpy(dst + starts[k], outv.data(), lens[k]);
          }
        } catch (...) { err[t] = std::current_exception(); }
      };
      std::vector<std::thread> pool;
      for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
      if (threads > 0) work(0);
      for (auto& th : pool) th.join();
      for (auto& e2 : err) if (e2) std::rethrow_exception(e2);
    }
    bgen_off[slot] = starts;
  }
  // device half
  void upload_bgen(int n, int slot) {
    cudaStream_t st = e->stream;
    d_bytes.upload(h_bgen[slot].p, bgen_off[slot][n] + 16, st);
    d_off.upload(bgen_off[slot].data(), bgen_off[slot].size(), st);
    d_err.ensure(n);
    n_rows_staged = n;
  }
  // staging memory for batches of up to max_n variants: the uncompressed size of a block is bounded by its layout
  void ensure_bgen(int max_n) {
    const uint64_t per = s->bgen_layout == 1 ? 6ull * s->n_samples : 16ull + 9ull * s->n_samples;  // ploidy bytes + two probabilities of <= 32 bits
    for (auto& b2 : h_bgen) b2.ensure((size_t)max_n * per + 16);
  }
  void stage_bgen(const uint32_t* variants, int n) {
    ensure_bgen(n);
    read_bgen(variants, n, 0);
    upload_bgen(n, 0);
  }


  void check_bgen_errors(int n) {
    h_err.resize(n);
    GW_CUDA(cudaMemcpyAsync(h_err.data(), d_err.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, e->stream));
    GW_CUDA(cudaStreamSynchronize(e->stream));
    for (int k = 0; k < n; ++k) {
      if (h_err[k] == 21) fail("set_min_max_ploidy: only diploid unphased genotype probabilities are implemented");  // loader.cpp:22-25
      if (h_err[k] == 22) fail("not implemented");  // loader.cpp:52-55 (missing sample in a v1.2 block)
      if (h_err[k]) fail("%s: malformed genotype data block", s->path.c_str());
    }
  }
};

// ---------------------------------------------------------------------------- checkpoint log

const char* status_text(int code) {
  switch (code) {
    case 0: return "OK";
    case 1: return "OK/green";
    case 2: return "infeasible linear constraint";
    case 3: return "infeasible non-linear constraint";
    case 4: return "iteration limit/blue";
    case 5: return "not convex/red";
    case 6: return "nonzero gradient/red";
    default: return "NA";
  }
}

struct ContextFile {  // mxComputeLoadContext: line idx of .pvar / .bim (R/model.R:150-160)
  std::vector<std::string> lines;
  bool bim = false;
  void load(const std::string& path) {
    bim = path.size() > 4 && path.compare(path.size() - 4, 4, ".bim") == 0;
    std::ifstream f(path);
    if (!f) fail("cannot open %s", path.c_str());
    std::string line;
    bool header_seen = bim;  // .bim has no header
    while (std::getline(f, line)) {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      if (!header_seen) {
        if (line.rfind("##", 0) == 0) continue;
        header_seen = true;
        continue;
      }
      lines.push_back(line);
    }
  }
  // tab separated CHR BP SNP A2 A1 (pvar) or CHR SNP BP A1 A2 (bim), as the reference orders them
  std::string fields(uint32_t idx) const {
    if (idx >= lines.size()) fail("LoadContext: line %u requested but only %zu lines available", idx + 1, lines.size());
    std::vector<std::string> tok;
    std::stringstream ss(lines[idx]);
    std::string t;
    const char sep = lines[idx].find('\t') != std::string::npos ? '\t' : ' ';
    while (std::getline(ss, t, sep))
      if (!t.empty() || sep == '\t') tok.push_back(t);
    std::string out;
    const int pv[5] = {0, 1, 2, 3, 4}, bm[5] = {0, 1, 3, 4, 5};
    for (int k = 0; k < 5; ++k) {
      const int c = bim ? bm[k] : pv[k];
      if (c >= (int)tok.size()) fail("LoadContext: line %u has only %zu columns", idx + 1, tok.size());
      if (k) out += '\t';
      out += tok[c];
    }
    return out;
  }
};

void append_double(std::string& out, double v) {  // as printf("%.15g"), NA for missing
  if (std::isnan(v)) { out += "NA"; return; }
  char buf[40];
  const auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::general, 15);
  out.append(buf, r.ptr);
}

// The log holds the variances and a few covariances of the parameters, not the P x P matrix: gather them on the device
// (out[snp][i] = vcov[snp][pa[i]][pb[i]]) so that a batch's trip to the host is about 1 KB per SNP instead of 8 P^2 bytes.
__global__ void gather_vcov_kernel(const double* __restrict__ vcov, int P, const int32_t* __restrict__ pa,
                                   const int32_t* __restrict__ pb, int np, int n, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)n * np) return;
  const int64_t k = e / np;
  const int i = (int)(e % np);
  out[e] = vcov[k * P * P + (int64_t)pa[i] * P + pb[i]];
}

// One batch of per-SNP results on the host, formatted into log rows by a pool of threads while
// the next batch is staged and fitted.
struct RowBatch {
  int n = 0;
  std::vector<uint32_t> variant;
  // pinned: the device -> host copies of a batch run at link speed and asynchronously, behind the next batch's kernels
  PinnedBuffer<double> est, vcov, fit, maf;
  PinnedBuffer<int32_t> n_used, status, catch_code, catch_var, iterations;
  void resize(int rows, int P, int n_vc) {   // vcov: the n_vc gathered entries of every row
    n = rows;
    variant.resize(rows);
    est.ensure((size_t)rows * P); vcov.ensure((size_t)rows * n_vc); fit.ensure(rows); maf.ensure(rows);
    n_used.ensure(rows); status.ensure(rows); catch_code.ensure(rows); catch_var.ensure(rows); iterations.ensure(rows);
  }
  gwsem_result view() {
    gwsem_result h;
    h.est = est.p; h.vcov = nullptr; h.fit = fit.p; h.maf = maf.p; h.n_used = n_used.p;   // (vcov travels gathered)
    h.status = status.p; h.catch_code = catch_code.p; h.catch_var = catch_var.p; h.iterations = iterations.p;
    return h;
  }
};

}  // namespace gwsem

using namespace gwsem;

// ------------------------------------------------------------------------------- C ABI: sources

extern "C" {

int gwsem_source_open(const char* path, const char* method, int src_rows, gwsem_source** out) {
  return guarded([&] { *out = open_source(path, method, src_rows); });
}
void gwsem_source_close(gwsem_source* s) { delete s; }
int gwsem_source_num_variants(const gwsem_source* s) { return s->variants_known ? (int)s->n_variants : -1; }
int gwsem_source_num_samples(const gwsem_source* s) { return (int)s->n_samples; }
int gwsem_source_load_counter(const gwsem_source* s) { return s->load_counter; }
const char* gwsem_source_checkpoint_columns(const gwsem_source* s) { return s->checkpoint_columns.c_str(); }

int gwsem_source_context(gwsem_source* s, int index, char* buf, size_t buflen) {
  return guarded([&] {
    std::string out;
    if (s->kind == gwsem_source::kBgen) {
      if (index < 0 || (uint32_t)index >= s->n_variants) fail("bgen: %s has no more varients", s->path.c_str());
      s->parse_bgen_variant((uint32_t)index);
      const auto& v = s->bgen[index];  // loader.cpp:133-140: SNP, RSID, CHR, BP, A2, A1
      out = v.snpid + "\t" + v.rsid + "\t" + v.chrom + "\t" + std::to_string(v.pos) + "\t" + v.a2 + "\t" + v.a1;
    }
    if (out.size() + 1 > buflen) fail("context buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
  });
}

int gwsem_source_load_rows(gwsem_engine* e, gwsem_source* s, const int32_t* snp_index, int n_snps,
                           const int32_t* row_filter, double* out, double* maf) {
  return guarded([&] {
    GW_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    if (!s->variants_known) fail("%s: the number of people is needed to read a PLINK 1 .bed file", s->path.c_str());
    Stager sg(e, s);
    const int N = (int)s->n_samples;
    std::vector<int32_t> dest_of(N);
    int dest_rows = 0;
    for (int i = 0; i < N; ++i) dest_of[i] = (row_filter && row_filter[i]) ? -1 : dest_rows++;
    DeviceBuffer<int32_t> d_dest;
    d_dest.upload(dest_of.data(), N, st);
    GW_CUDA(cudaStreamSynchronize(st));
    const int batch = (int)std::max<int64_t>(1, std::min<int64_t>(4096, (int64_t)(1u << 27) / std::max(dest_rows, 1)));
    DeviceBuffer<double> d_out, d_maf;
    std::vector<uint32_t> v;
    for (int s0 = 0; s0 < n_snps; s0 += batch) {
      const int n = std::min(batch, n_snps - s0);
      v.resize(n);
      for (int k = 0; k < n; ++k) {
        if (snp_index[s0 + k] < 0) fail("negative SNP index");
        v[k] = (uint32_t)snp_index[s0 + k];
      }
      d_out.ensure((size_t)n * std::max(dest_rows, 1));
      d_maf.ensure(n);
      if (s->kind == gwsem_source::kPgen) {
        sg.stage_pgen(v.data(), n);
        if (sg.plink1)  // rows of a .bed are staged verbatim: decode with the PLINK 1 table
          launch_decode_2bit(e, sg.d_rows.p, sg.pitch, n, N, 1, d_dest.p, dest_rows, d_out.p, d_maf.p);
        else
          launch_rows_to_doubles(e, sg.d_rows.p, sg.pitch, sg.has_dosage ? sg.d_dosage.p : nullptr, sg.n_pad, n, N,
                                 d_dest.p, nullptr, dest_rows, d_out.p, d_maf.p);
      } else {
        sg.stage_bgen(v.data(), n);
        BgenBatch bb;
        bb.bytes = sg.d_bytes.p; bb.off = sg.d_off.p; bb.n = n; bb.n_samples = N; bb.layout = (int)s->bgen_layout;
        bb.dest_of = d_dest.p; bb.out_stride = dest_rows; bb.out = d_out.p; bb.maf = d_maf.p; bb.err = sg.d_err.p;
        launch_bgen_blocks(e, bb);
        sg.check_bgen_errors(n);
      }
      GW_CUDA(cudaMemcpyAsync(out + (size_t)s0 * dest_rows, d_out.p, sizeof(double) * (size_t)n * dest_rows,
                              cudaMemcpyDeviceToHost, st));
      if (maf) GW_CUDA(cudaMemcpyAsync(maf + s0, d_maf.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
      GW_CUDA(cudaStreamSynchronize(st));
    }
  });
}

int gwsem_fit_rows_device(gwsem_model* m, const double* d_geno, int n_snps, const gwsem_result* d_res) {
  return guarded([&] { m->fit_rows_device(d_geno, d_res->maf, n_snps, *d_res); });
}

// ------------------------------------------------------------------------------ the whole loop

int gwsem_gwas_run(gwsem_model* m, gwsem_source* s, const gwsem_job* job, int64_t* rows_written) {
  return guarded([&] {
    gwsem_engine* e = m->engine;
    GW_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    if (!s->variants_known) fail("%s: the number of people is needed to read a PLINK 1 .bed file", s->path.c_str());
    if (!m->prepared) m->prepare(nullptr, (int)s->n_samples);
    if (m->n_src != (int)s->n_samples)
      fail("%s has %d samples, which does not match the number of rows of observed data before filtering (%d)",
           s->path.c_str(), (int)s->n_samples, m->n_src);
    const int P = m->P;
    // SNP list (1-based in the API and in the log, R/model.R:59-64)
    std::vector<uint32_t> todo;
    if (job->n_snp >= 0) {   // an explicit list, possibly empty (a rank whose shard is empty writes the header only)
      if (job->n_snp > 0 && !job->snp) fail("n_snp = %d but no SNP list", job->n_snp);
      for (int k = 0; k < job->n_snp; ++k) {
        if (job->snp[k] < 1) fail("SNP indices are 1-based");
        todo.push_back((uint32_t)job->snp[k] - 1);
      }
    } else {
      const int from = std::max(1, job->start_from);
      for (uint32_t v = (uint32_t)from - 1; v < s->n_variants; ++v) todo.push_back(v);
    }
    ContextFile ctx;
    if (job->context_path && job->context_path[0]) ctx.load(job->context_path);

    // the log is the checkpoint (R/model.R:191): whatever happens later, rows already formatted reach the disk
    struct LogFile {
      FILE* f = nullptr;
      ~LogFile() { if (f) fclose(f); }
    } log_guard;
    FILE* fo = log_guard.f = fopen(job->out_path, "w");
    if (!fo) fail("cannot open %s for writing", job->out_path);
    std::string line;
    line.reserve(1 << 16);
    // header (column names are what R/report.R:91-113 selects by)
    line = "OpenMxEvals\titerations\ttimestamp\tMxComputeLoop1";
    for (int k = 0; k < P; ++k) { line += '\t'; line += job->par_names[k]; }
    line += "\tfit\tfitUnits\tsampleSize\tstatusCode";
    std::vector<std::pair<int, int>> vc;
    for (int k = 0; k < P; ++k) vc.emplace_back(k, k);
    for (int k = 0; k < job->n_vcov_pairs; ++k) {
      int a = job->vcov_pairs[2 * k], b = job->vcov_pairs[2 * k + 1];
      if (a < 0 || b < 0 || a >= P || b >= P || a == b) continue;
      vc.emplace_back(std::max(a, b), std::min(a, b));
    }
    std::sort(vc.begin(), vc.end(), [](auto& x, auto& y) { return x.second != y.second ? x.second < y.second : x.first < y.first; });
    vc.erase(std::unique(vc.begin(), vc.end()), vc.end());
    for (auto& pr : vc) { line += "\tV"; line += job->par_names[pr.first]; line += ':'; line += job->par_names[pr.second]; }
    if (s->kind == gwsem_source::kPgen) {
      if (!ctx.lines.empty()) line += ctx.bim ? "\tCHR\tSNP\tBP\tA1\tA2" : "\tCHR\tBP\tSNP\tA2\tA1";
      line += "\tMAF";
    } else {
      line += "\tSNP\tRSID\tCHR\tBP\tA2\tA1\tMAF";
    }
    line += "\tcatch1\n";
    fwrite(line.data(), 1, line.size(), fo);

    Stager sg(e, s);
    const bool hard_file = s->kind == gwsem_source::kPgen && !s->any_dosage;
    int batch = job->batch_snps > 0 ? job->batch_snps
                                    : (int)(hard_file ? std::min<int64_t>(1 << 16, std::max<int64_t>(64, (384ll << 20) / sg.pitch))
                                                      : std::min<int64_t>(4096, std::max<int64_t>(16, (512ll << 20) / (8ll * sg.n_pad))));
    if (job->batch_snps <= 0 && hard_file) {
      // whole waves of the 128-SNP class-sum CTAs over the SMs (18,944 SNPs on 148 SMs), else whole CTAs
      const int wave = 128 * e->sm_count;
      if (batch >= wave) batch = batch / wave * wave;
      else if (batch >= 128) batch = batch / 128 * 128;
    }
    RowBatch rb[2];
    std::future<void> pending;  // formatting + writing of the previous batch
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    // rows [k0, k1) of batch b as text
    auto format_rows = [&](const RowBatch& b, int k0, int k1, const char* tbuf, std::string& line) {
      for (int k = k0; k < k1; ++k) {
        const bool caught = b.catch_code.p[k] != GWSEM_CATCH_NONE;
        line += std::to_string(b.iterations.p[k] + 1);  // objective evaluations are not tracked separately
        line += '\t';
        line += std::to_string(b.iterations.p[k]);
        line += '\t';
        line += tbuf;
        line += '\t';
        line += std::to_string(b.variant[k] + 1);
        for (int a = 0; a < P; ++a) { line += '\t'; append_double(line, b.est.p[(size_t)k * P + a]); }
        line += '\t';
        append_double(line, b.fit.p[k]);
        line += m->fit_kind == 2 ? "\t-2lnL\t" : "\tr'Wr\t";
        line += std::to_string(b.n_used.p[k]);
        line += '\t';
        line += caught ? "NA" : status_text(b.status.p[k]);
        for (size_t i = 0; i < vc.size(); ++i) { line += '\t'; append_double(line, b.vcov.p[(size_t)k * vc.size() + i]); }
        if (s->kind == gwsem_source::kPgen) {
          if (!ctx.lines.empty()) { line += '\t'; line += ctx.fields(b.variant[k]); }
        } else {
          const auto& bv = s->bgen[b.variant[k]];
          line += '\t'; line += bv.snpid; line += '\t'; line += bv.rsid; line += '\t'; line += bv.chrom; line += '\t';
          line += std::to_string(bv.pos); line += '\t'; line += bv.a2; line += '\t'; line += bv.a1;
        }
        char mbuf[32];
        snprintf(mbuf, sizeof(mbuf), "\t%.8f\t", b.maf.p[k]);  // loader.cpp:404
        line += mbuf;
        if (b.catch_code.p[k] == GWSEM_CATCH_MIN_VARIANCE) {
          char cbuf[256];
          const char* var = b.catch_var.p[k] == -2 ? "snp" : (job->manifest_names && b.catch_var.p[k] >= 0) ? job->manifest_names[b.catch_var.p[k]] : "?";
          snprintf(cbuf, sizeof(cbuf), "%s.data: '%s' has observed variance less than %g", job->model_name ? job->model_name : "model",
                   var, m->min_variance);
          line += cbuf;
        } else if (b.catch_code.p[k] == GWSEM_CATCH_ACOV_NOT_PD) {
          line += job->model_name ? job->model_name : "model";
          line += ".data: asymptotic covariance matrix is not positive definite";
        }
        line += '\n';
      }
    };
    auto write_batch = [&](const RowBatch& b) {
      char tbuf[64];
      time_t now = time(nullptr);
      struct tm tmv;
      localtime_r(&now, &tmv);
      strftime(tbuf, sizeof(tbuf), "%b %d %Y %I:%M:%S %p", &tmv);
      const int parts = (int)std::min<int64_t>(std::min(hw, 32u), std::max(1, b.n / 256));
      std::vector<std::string> text(parts);
      std::vector<std::thread> pool;
      std::vector<std::exception_ptr> err(parts);   // format_rows can fail (a .pvar / .bim shorter than the file)
      auto work = [&](int t) {
        try {
          format_rows(b, (int)((int64_t)b.n * t / parts), (int)((int64_t)b.n * (t + 1) / parts), tbuf, text[t]);
        } catch (...) { err[t] = std::current_exception(); }
      };
      for (int t = 1; t < parts; ++t) pool.emplace_back(work, t);
      work(0);
      for (auto& th : pool) th.join();
      for (auto& e2 : err)
        if (e2) std::rethrow_exception(e2);
      for (auto& part : text)
        if (fwrite(part.data(), 1, part.size(), fo) != part.size()) fail("write error on %s", job->out_path);
    };
    DeviceBuffer<double> d_geno, d_maf;
    int64_t written = 0;
    // fixed-width files: a helper thread reads batch k+1 into pinned memory while batch k is on the device
    const bool ahead = sg.fixed_width();
    std::future<void> reading;
    auto start_read = [&](size_t at) {
      if (at >= todo.size()) return;
      const int cnt = (int)std::min<size_t>(batch, todo.size() - at);
      const int slot = (int)((at / batch) & 1);
      reading = std::async(std::launch::async, [&sg, &todo, at, cnt, slot] { sg.read_fixed(&todo[at], cnt, slot); });
    };
    const bool ahead_bgen = s->kind == gwsem_source::kBgen;
    auto start_read_bgen = [&](size_t at) {
      if (at >= todo.size()) return;
      const int cnt = (int)std::min<size_t>(batch, todo.size() - at);
      const int slot = (int)((at / batch) & 1);
      reading = std::async(std::launch::async, [&sg, &todo, at, cnt, slot] { sg.read_bgen(&todo[at], cnt, slot); });
    };
    if (ahead) {
      sg.ensure_fixed((int)std::min<size_t>(batch, todo.size()));
      start_read(0);
    }
    if (ahead_bgen) {
      sg.ensure_bgen((int)std::min<size_t>(batch, todo.size()));
      start_read_bgen(0);
    }
    static const bool timing = getenv("GWSEM_TIMING") != nullptr;   // development aid: phase times on stderr
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_stage = 0, t_fit = 0, t_wait = 0, t_dl = 0;
    const size_t cap = std::min<size_t>(batch, todo.size());
    m->ensure_result_buffers((int)(2 * cap));
    const int n_vc = (int)vc.size();
    DeviceBuffer<int32_t> d_vc;
    DeviceBuffer<double> d_vcg;   // gathered entries, two halves like the result buffers
    {
      std::vector<int32_t> hv(2 * (size_t)n_vc);
      for (int i = 0; i < n_vc; ++i) { hv[i] = vc[i].first; hv[n_vc + i] = vc[i].second; }
      d_vc.ensure(hv.size() + 1);
      GW_CUDA(cudaMemcpyAsync(d_vc.p, hv.data(), hv.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      GW_CUDA(cudaStreamSynchronize(st));
      d_vcg.ensure(2 * cap * (size_t)n_vc + 1);
    }
    cudaStream_t out_stream = nullptr;
    cudaEvent_t ev_fitted[2] = {nullptr, nullptr}, ev_down[2] = {nullptr, nullptr}, ev_up[2] = {nullptr, nullptr};
    GW_CUDA(cudaStreamCreateWithFlags(&out_stream, cudaStreamNonBlocking));
    for (int b2 = 0; b2 < 2; ++b2) {
      GW_CUDA(cudaEventCreateWithFlags(&ev_fitted[b2], cudaEventDisableTiming));
      GW_CUDA(cudaEventCreateWithFlags(&ev_down[b2], cudaEventDisableTiming));
      GW_CUDA(cudaEventCreateWithFlags(&ev_up[b2], cudaEventDisableTiming));
    }
    bool up_used[2] = {false, false};
    auto release_streams = [&] {
      if (out_stream) { cudaStreamSynchronize(out_stream); cudaStreamDestroy(out_stream); out_stream = nullptr; }
      for (int b2 = 0; b2 < 2; ++b2) {
        if (ev_fitted[b2]) cudaEventDestroy(ev_fitted[b2]);
        if (ev_down[b2]) cudaEventDestroy(ev_down[b2]);
        if (ev_up[b2]) cudaEventDestroy(ev_up[b2]);
        ev_fitted[b2] = ev_down[b2] = ev_up[b2] = nullptr;
      }
    };
    bool have_prev = false;
    int last_slot = 0;
    try {
    for (size_t t0 = 0; t0 < todo.size(); t0 += batch) {
      const int n = (int)std::min<size_t>(batch, todo.size() - t0);
      const uint32_t* v = &todo[t0];
      const double c0 = now();
      const int slot = (int)((t0 / batch) & 1);
      // results of consecutive batches alternate between the two halves of the device buffers: batch k computes while
      // batch k - 1 is copied out (the copy of batch k - 2, which used this half, was awaited one iteration ago)
      gwsem_result dres = m->device_result();
      {
        const size_t o = (size_t)slot * cap;
        dres.est += o * P; dres.vcov += o * P * P; dres.fit += o; dres.maf += o; dres.n_used += o;
        dres.status += o; dres.catch_code += o; dres.catch_var += o; dres.iterations += o;
      }
      if (ahead) {
        reading.get();
        sg.upload_fixed(n, slot);
        GW_CUDA(cudaEventRecord(ev_up[slot], st));
        up_used[slot] = true;
        // the reader refills the other pinned slot: the upload that read it (previous batch) must be over
        if (up_used[slot ^ 1]) GW_CUDA(cudaEventSynchronize(ev_up[slot ^ 1]));
        start_read(t0 + batch);
        m->fit_2bit_device(sg.d_rows.p, sg.pitch, n, sg.plink1, dres);
      } else if (s->kind == gwsem_source::kPgen) {
        sg.stage_pgen(v, n);
        if (!sg.has_dosage) {
          m->fit_2bit_device(sg.d_rows.p, sg.pitch, n, sg.plink1, dres);
        } else {
          d_geno.ensure((size_t)n * sg.n_pad);
          d_maf.ensure(n);
          launch_rows_to_doubles(e, sg.d_rows.p, sg.pitch, sg.d_dosage.p, sg.n_pad, n, (int)s->n_samples, nullptr,
                                 m->layout.d_rowmask, sg.n_pad, d_geno.p, d_maf.p);
          m->fit_rows_device(d_geno.p, d_maf.p, n, dres);
        }
      } else {
        reading.get();
        sg.upload_bgen(n, slot);
        GW_CUDA(cudaEventRecord(ev_up[slot], st));
        up_used[slot] = true;
        if (up_used[slot ^ 1]) GW_CUDA(cudaEventSynchronize(ev_up[slot ^ 1]));
        start_read_bgen(t0 + batch);
        if (timing) { GW_CUDA(cudaStreamSynchronize(st)); t_stage += now() - c0; }
        d_geno.ensure((size_t)n * sg.n_pad);
        d_maf.ensure(n);
        BgenBatch bb;
        bb.bytes = sg.d_bytes.p; bb.off = sg.d_off.p; bb.n = n; bb.n_samples = (int)s->n_samples;
        bb.layout = (int)s->bgen_layout; bb.dest_of = nullptr; bb.rowmask = m->layout.d_rowmask;
        bb.out_stride = sg.n_pad; bb.out = d_geno.p; bb.maf = d_maf.p; bb.err = sg.d_err.p;
        launch_bgen_blocks(e, bb);
        sg.check_bgen_errors(n);
        m->fit_rows_device(d_geno.p, d_maf.p, n, dres);
      }
      double c1 = 0;
      if (timing) { GW_CUDA(cudaStreamSynchronize(st)); c1 = now(); t_fit += c1 - c0; }
      double* vcg = d_vcg.p + (size_t)slot * cap * n_vc;
      if (n_vc > 0) {
        const int64_t work = (int64_t)n * n_vc;
        gather_vcov_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(dres.vcov, P, d_vc.p, d_vc.p + n_vc, n_vc, n, vcg);
        GW_CUDA(cudaGetLastError());
      }
      GW_CUDA(cudaEventRecord(ev_fitted[slot], st));
      // the previous batch: its results are on the host once its download is done; hand it to the writer
      // (one batch being written, one being downloaded, one on the device)
      if (have_prev) {
        GW_CUDA(cudaEventSynchronize(ev_down[slot ^ 1]));
        if (timing) { t_dl += now() - c1; c1 = now(); }
        if (pending.valid()) pending.get();
        if (timing) { t_wait += now() - c1; c1 = now(); }
        RowBatch& prev = rb[slot ^ 1];
        pending = std::async(std::launch::async, [&write_batch, &prev] { write_batch(prev); });
      } else if (pending.valid()) {
        pending.get();
      }
      RowBatch& cur = rb[slot];   // (its previous contents were written two batches ago: that writer was joined above)
      cur.resize(n, P, n_vc);
      std::copy(v, v + n, cur.variant.begin());
      GW_CUDA(cudaStreamWaitEvent(out_stream, ev_fitted[slot], 0));
      m->download_result(cur.view(), dres, n, 0, out_stream);
      if (n_vc > 0)
        GW_CUDA(cudaMemcpyAsync(cur.vcov.p, vcg, sizeof(double) * (size_t)n * n_vc, cudaMemcpyDeviceToHost, out_stream));
      GW_CUDA(cudaEventRecord(ev_down[slot], out_stream));
      have_prev = true;
      last_slot = slot;
      written += n;
    }
    if (have_prev) {
      GW_CUDA(cudaEventSynchronize(ev_down[last_slot]));
      if (pending.valid()) pending.get();
      RowBatch& prev = rb[last_slot];
      pending = std::async(std::launch::async, [&write_batch, &prev] { write_batch(prev); });
    }
    if (pending.valid()) pending.get();
    } catch (...) {
      // the helper threads use this frame's objects: let them finish before it unwinds
      if (pending.valid()) pending.wait();
      if (reading.valid()) reading.wait();
      cudaStreamSynchronize(st);
      release_streams();
      throw;
    }
    release_streams();
    if (timing)
      fprintf(stderr, "gwsem timing: batch %d, stage(bgen only) %.3f s, stage+fit %.3f s, wait-for-writer %.3f s, download %.3f s\n",
              batch, t_stage, t_fit, t_wait, t_dl);
    log_guard.f = nullptr;
    if (fflush(fo) != 0 || ferror(fo)) { fclose(fo); fail("write error on %s", job->out_path); }
    if (fclose(fo) != 0) fail("write error on %s (close)", job->out_path);
    if (rows_written) *rows_written = written;
  });
}

}  // extern "C"
