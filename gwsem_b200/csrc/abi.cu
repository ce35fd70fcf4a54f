// extern "C" surface of libgwsem_b200.so (include/gwsem_b200.h).
#include <algorithm>
#include <cstring>
#include <mutex>

#include "common.h"
#include "kernels.h"
#include "model.h"

namespace gwsem {

static thread_local std::string g_error;

void set_error(const char* fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
}

std::string format(const char* fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return buf;
}

void fail(const char* fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw Error(buf);
}

}  // namespace gwsem

using namespace gwsem;

// FP64 FMA issue-rate probe (the roofline of the statistics kernels of the WLS marginals path): NACC independent
// dependent chains of DFMAs per thread, no memory traffic.
template <int NACC>
__global__ void fp64_peak_kernel(double* out, int iters, double w) {
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = k + threadIdx.x;
  const double h = 1.0 + threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = fma(w, h, acc[k]);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < NACC; ++k) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

void gwsem_engine::prof_begin(int fam) {
  if (!profiling) return;
  std::pair<cudaEvent_t, cudaEvent_t> ev;
  if (!prof_free.empty()) {
    ev = prof_free.back();
    prof_free.pop_back();
  } else {
    GW_CUDA(cudaEventCreate(&ev.first));
    GW_CUDA(cudaEventCreate(&ev.second));
  }
  GW_CUDA(cudaEventRecord(ev.first, stream));
  prof_events[fam].push_back(ev);
}

void gwsem_engine::prof_end(int fam) {
  if (!profiling) return;
  GW_CUDA(cudaEventRecord(prof_events[fam].back().second, stream));
}

extern "C" {

int gwsem_engine_profile(gwsem_engine* e, int enable) {
  return guarded([&] { e->profiling = enable != 0; });
}

// Collects (and clears) the per-family launch statistics gathered since the last call:
// launches[f], total_ms[f] for the eight families of include/gwsem_b200.h.  Synchronises.
int gwsem_engine_profile_read(gwsem_engine* e, int64_t* launches, double* total_ms) {
  return guarded([&] {
    GW_CUDA(cudaSetDevice(e->device));
    GW_CUDA(cudaStreamSynchronize(e->stream));
    for (int f = 0; f < kFamCount; ++f) {
      launches[f] = (int64_t)e->prof_events[f].size();
      total_ms[f] = 0;
      for (auto& ev : e->prof_events[f]) {
        float ms = 0;
        GW_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
        total_ms[f] += ms;
        e->prof_free.push_back(ev);
      }
      e->prof_events[f].clear();
    }
  });
}

// Best of a few launch shapes of the DFMA probe, in TFLOP/s (2 flops per DFMA), timed with CUDA events on the
// engine's stream at whatever clocks the GPU runs at right now.
int gwsem_engine_fp64_peak(gwsem_engine* e, double* tflops) {
  return guarded([&] {
    GW_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    constexpr int kAcc = 16, kIters = 40000;
    DeviceBuffer<double> out;
    out.ensure((size_t)e->sm_count * 4 * 1024);
    cudaEvent_t a, b;
    GW_CUDA(cudaEventCreate(&a));
    GW_CUDA(cudaEventCreate(&b));
    double best = 0.0;
    for (int warps : {8, 16, 32}) {
      for (int rep = 0; rep < 3; ++rep) {
        const int grid = e->sm_count * (warps > 16 ? 2 : 1), block = 32 * (warps > 16 ? warps / 2 : warps);
        fp64_peak_kernel<kAcc><<<grid, block, 0, st>>>(out.p, rep == 0 ? 200 : kIters, 1.0);
        if (rep == 0) continue;
        GW_CUDA(cudaEventRecord(a, st));
        fp64_peak_kernel<kAcc><<<grid, block, 0, st>>>(out.p, kIters, 1.0);
        GW_CUDA(cudaEventRecord(b, st));
        GW_CUDA(cudaEventSynchronize(b));
        float ms = 0;
        GW_CUDA(cudaEventElapsedTime(&ms, a, b));
        const double flops = 2.0 * (double)grid * block * kIters * kAcc;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
      }
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = best;
  });
}

int gwsem_abi_version(void) { return GWSEM_ABI_VERSION; }
const char* gwsem_last_error(void) { return g_error.c_str(); }

int gwsem_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int gwsem_engine_create(int device, gwsem_engine** out) {
  return guarded([&] {
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess || n == 0)
      fail("no usable CUDA device (%s): gwsem_b200 has no CPU fallback", err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
    if (device < 0 || device >= n) fail("device %d out of range (0..%d)", device, n - 1);
    GW_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GW_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) fail("device %d is sm_%d%d; this build targets sm_100a (B200)", device, prop.major, prop.minor);
    auto* e = new gwsem_engine;
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    GW_CUDA(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
    e->stream = e->own_stream;
    *out = e;
  });
}

void gwsem_engine_destroy(gwsem_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  for (void* q : e->pinned_stage)
    if (q) cudaFreeHost(q);
  delete e;
}

int gwsem_engine_set_stream(gwsem_engine* e, void* cuda_stream) {
  return guarded([&] { e->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : e->own_stream; });
}

int gwsem_engine_synchronize(gwsem_engine* e) {
  return guarded([&] {
    GW_CUDA(cudaSetDevice(e->device));
    GW_CUDA(cudaStreamSynchronize(e->stream));
  });
}

int64_t gwsem_engine_launch_count(const gwsem_engine* e) { return e->launches; }
int64_t gwsem_engine_series_columns(const gwsem_engine* e) { return e->series_columns; }

int gwsem_decode_2bit_device(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int n_samples,
                             int plink1, const int32_t* d_dest_of, int dest_rows, double* d_out, double* d_maf) {
  return guarded([&] {
    GW_CUDA(cudaSetDevice(e->device));
    launch_decode_2bit(e, d_packed, pitch, n_snps, n_samples, plink1, d_dest_of, dest_rows, d_out, d_maf);
  });
}

int gwsem_decode_2bit(gwsem_engine* e, const uint8_t* packed, int64_t pitch, int n_snps, int n_samples, int plink1,
                      const int32_t* row_filter, double* out, double* maf) {
  return guarded([&] {
    GW_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    const int64_t row_bytes = (n_samples + 3) / 4;
    if (pitch < row_bytes) fail("pitch %lld smaller than record (%lld bytes)", (long long)pitch, (long long)row_bytes);
    int dest_rows = n_samples;
    const int32_t* d_dest = nullptr;
    if (row_filter) {
      std::vector<int32_t> dest_of(n_samples);
      dest_rows = 0;
      for (int s = 0; s < n_samples; ++s) dest_of[s] = row_filter[s] ? -1 : dest_rows++;
      e->d_dest_of.upload(dest_of.data(), n_samples, st);
      GW_CUDA(cudaStreamSynchronize(st));  // dest_of is a local
      d_dest = e->d_dest_of.p;
    }
    // bound the double matrix held on the device
    const int batch = (int)std::max<int64_t>(1, std::min<int64_t>(n_snps, (int64_t)(1u << 28) / std::max(dest_rows, 1)));
    for (int s0 = 0; s0 < n_snps; s0 += batch) {
      const int n = std::min(batch, n_snps - s0);
      e->d_packed.ensure((size_t)n * row_bytes);
      e->d_rows.ensure((size_t)n * dest_rows);
      e->d_maf.ensure(n);
      GW_CUDA(cudaMemcpy2DAsync(e->d_packed.p, row_bytes, packed + (int64_t)s0 * pitch, pitch, row_bytes, n,
                                cudaMemcpyHostToDevice, st));
      launch_decode_2bit(e, e->d_packed.p, row_bytes, n, n_samples, plink1, d_dest, dest_rows, e->d_rows.p, e->d_maf.p);
      GW_CUDA(cudaMemcpyAsync(out + (size_t)s0 * dest_rows, e->d_rows.p, sizeof(double) * (size_t)n * dest_rows,
                              cudaMemcpyDeviceToHost, st));
      if (maf) GW_CUDA(cudaMemcpyAsync(maf + s0, e->d_maf.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
      GW_CUDA(cudaStreamSynchronize(st));
    }
  });
}

int gwsem_model_create(gwsem_engine* e, const gwsem_model_spec* spec, gwsem_model** out) {
  return guarded([&] {
    GW_CUDA(cudaSetDevice(e->device));
    *out = new gwsem_model(e, spec);
  });
}

void gwsem_model_destroy(gwsem_model* m) {
  if (!m) return;
  cudaSetDevice(m->engine->device);
  cudaStreamSynchronize(m->engine->stream);
  if (m->side_stream) { cudaStreamSynchronize(m->side_stream); cudaStreamDestroy(m->side_stream); }
  if (m->copy_stream) { cudaStreamSynchronize(m->copy_stream); cudaStreamDestroy(m->copy_stream); }
  if (m->out_stream) { cudaStreamSynchronize(m->out_stream); cudaStreamDestroy(m->out_stream); }
  for (cudaEvent_t ev : {m->ev_side_go, m->ev_side_done, m->ev_copied[0], m->ev_copied[1], m->ev_done[0], m->ev_done[1],
                         m->ev_fitted[0], m->ev_fitted[1], m->ev_out[0], m->ev_out[1]})
    if (ev) cudaEventDestroy(ev);
  delete m;
}

int gwsem_model_set_row_filter(gwsem_model* m, const int32_t* row_filter, int src_rows) {
  return guarded([&] { m->prepare(row_filter, src_rows); });
}

int gwsem_fit_2bit_device(gwsem_model* m, const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1,
                          const gwsem_result* d_res) {
  return guarded([&] { m->fit_2bit_device(d_packed, pitch, n_snps, plink1, *d_res); });
}

int gwsem_fit_2bit(gwsem_model* m, const uint8_t* packed, int64_t pitch, int n_snps, int plink1,
                   const gwsem_result* res) {
  return guarded([&] { m->fit_2bit_host(packed, pitch, n_snps, plink1, *res); });
}

}  // extern "C"

