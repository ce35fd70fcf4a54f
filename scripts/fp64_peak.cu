// Development aid: FP64 FMA issue-rate micro-benchmark (the DFMA roofline of class_sums_kernel).
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC, int INT_OPS>
__global__ void dfma_kernel(double* out, int iters, double w, unsigned seed) {
  double acc[NACC];
  for (int k = 0; k < NACC; ++k) acc[k] = k + threadIdx.x;
  double h = 1.0 + threadIdx.x * 1e-9;
  unsigned x = seed + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = fma(w, h, acc[k]);
#pragma unroll
    for (int k = 0; k < INT_OPS; ++k) x = (x >> 3) ^ (x * 5u + k);
  }
  double s = 0;
  for (int k = 0; k < NACC; ++k) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + x;
}
template <int NACC, int INT_OPS>
void run(int warps_per_sm, const char* name) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * warps_per_sm * 32);
  const int iters = 20000;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  dfma_kernel<NACC, INT_OPS><<<sms, warps_per_sm * 32>>>(out, 100, 1.0, 1);
  cudaEventRecord(a);
  dfma_kernel<NACC, INT_OPS><<<sms, warps_per_sm * 32>>>(out, iters, 1.0, 1);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double dfma = (double)sms * warps_per_sm * 32 * iters * NACC;
  printf("%-28s warps/SM=%2d  %.3f ms  %.2f TFLOP/s  %.1f DFMA/clk/SM @%d MHz nominal\n", name, warps_per_sm, ms,
         2 * dfma / ms / 1e9, dfma / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  cudaFree(out);
}
int main() {
  for (int w : {1, 2, 4, 8, 16}) run<16, 0>(w, "16 acc, pure DFMA");
  for (int w : {4, 8, 16}) run<48, 0>(w, "48 acc, pure DFMA");
  for (int w : {4, 8, 16}) run<12, 7>(w, "12 DFMA + 14 int (like cs)");
  for (int w : {4, 8, 16}) run<12, 3>(w, "12 DFMA + 6 int");
  return 0;
}
