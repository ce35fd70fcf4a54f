// Development aid: issue rate of the legacy tensor-core instructions the class-sum contraction can use:
// mma.sync m16n8k32 s8 (IMMA) against m16n8k256 b1 and.popc (BMMA).
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND, int NACC>
__global__ void mma_kernel(int* out, int iters, unsigned seed) {
  int acc[NACC][4];
  for (int k = 0; k < NACC; ++k)
    for (int j = 0; j < 4; ++j) acc[k][j] = 0;
  unsigned a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u, b0 = a0 ^ 0x1234u, b1 = a0 ^ 0x4321u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(acc[k][0]), "+r"(acc[k][1]), "+r"(acc[k][2]), "+r"(acc[k][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(acc[k][0]), "+r"(acc[k][1]), "+r"(acc[k][2]), "+r"(acc[k][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  int s = 0;
  for (int k = 0; k < NACC; ++k)
    for (int j = 0; j < 4; ++j) s += acc[k][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND, int NACC>
void run(int warps_per_sm, const char* name) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  int* out; cudaMalloc(&out, sizeof(int) * sms * warps_per_sm * 32);
  const int iters = 20000;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  mma_kernel<KIND, NACC><<<sms, warps_per_sm * 32>>>(out, 100, 1);
  cudaEventRecord(a);
  mma_kernel<KIND, NACC><<<sms, warps_per_sm * 32>>>(out, iters, 1);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double n = (double)sms * warps_per_sm * iters * NACC;
  printf("%-10s warps/SM=%2d  %.3f ms  %.3f warp-MMA/clk/SM @%d MHz nominal\n", name, warps_per_sm, ms,
         n / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8, 16}) run<0, 8>(w, "imma s8");
  for (int w : {4, 8, 16}) run<1, 8>(w, "bmma b1");
  return 0;
}
