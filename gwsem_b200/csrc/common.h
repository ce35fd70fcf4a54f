// Shared declarations of the gwsem_b200 engine (internal; the public surface is
// include/gwsem_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gwsem_b200.h"

namespace gwsem {

void set_error(const char* fmt, ...);
std::string format(const char* fmt, ...);

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};
[[noreturn]] void fail(const char* fmt, ...);

#define GW_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t err__ = (call);                                                         \
    if (err__ != cudaSuccess)                                                           \
      ::gwsem::fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
  } while (0)

// Wraps an extern "C" entry point body: exceptions never cross the ABI.
template <typename F>
int guarded(F&& body) {
  try {
    body();
    return 0;
  } catch (const std::exception& e) {
    set_error("%s", e.what());
    return 1;
  } catch (...) {
    set_error("unknown error");
    return 1;
  }
}

template <typename T>
struct DeviceBuffer {
  T* p = nullptr;
  size_t n = 0;
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  DeviceBuffer(DeviceBuffer&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  ~DeviceBuffer() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void ensure(size_t count) {
    if (count <= n) return;
    release();
    GW_CUDA(cudaMalloc(&p, count * sizeof(T)));
    n = count;
  }
  void upload(const T* host, size_t count, cudaStream_t st) {
    ensure(count);
    GW_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, st));
  }
};

}  // namespace gwsem

namespace gwsem {
enum KernelFamily { kFamClassSums = 0, kFamCountMissing = 1, kFamFit = 2, kFamDecode = 3, kFamMargStats = 4, kFamMargExact = 5, kFamMargSeries = 6, kFamSeriesSums = 7, kFamCount = 8 };
}

struct gwsem_engine {
  // optional per-launch timing with CUDA events on the launch stream (bench.py's roofline leg)
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[gwsem::kFamCount];
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_free;
  void prof_begin(int fam);
  void prof_end(int fam);
  int device = 0;
  int sm_count = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  // pinned host staging of gwsem_gwas_run (two slots of packed rows), kept between calls: pinning a few hundred MB
  // costs more than fitting the batch it carries
  void* pinned_stage[2] = {nullptr, nullptr};
  size_t pinned_stage_bytes[2] = {0, 0};
  int64_t series_columns = 0;   // digit columns of the series-path class-sum launches, summed over launches
  // scratch reused by the host-pointer entry points
  gwsem::DeviceBuffer<uint8_t> d_packed;
  gwsem::DeviceBuffer<double> d_rows, d_maf;
  gwsem::DeviceBuffer<int32_t> d_dest_of;
  gwsem::DeviceBuffer<unsigned long long> d_maf_acc;
};
