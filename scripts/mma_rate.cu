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
      else if (KIND == 1)  // sm_100a has no BMMA: ptxas emits a CALL to an emulation routine per instruction
        asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(acc[k][0]), "+r"(acc[k][1]), "+r"(acc[k][2]), "+r"(acc[k][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+r"(acc[k][0]), "+r"(acc[k][1]), "+r"(acc[k][2]), "+r"(acc[k][3])
                     : "r"(a0), "r"(a1), "r"(b0));
    }
  }
  int s = 0;
  for (int k = 0; k < NACC; ++k)
    for (int j = 0; j < 4; ++j) s += acc[k][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// IMMA mixed with independent integer ALU work (PRMT / SHF as in the one-hot expansion): do the pipes overlap?
template <int NACC, int ALU>
__global__ void mix_kernel(int* out, int iters, unsigned seed) {
  int acc[NACC + 1][4];
  for (int k = 0; k < NACC; ++k)
    for (int j = 0; j < 4; ++j) acc[k][j] = 0;
  unsigned a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u, b0 = a0 ^ 0x1234u, b1 = a0 ^ 0x4321u;
  unsigned x[4] = {a0, a1, a2, a3};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < NACC; ++k)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(acc[k][0]), "+r"(acc[k][1]), "+r"(acc[k][2]), "+r"(acc[k][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
#pragma unroll
    for (int k = 0; k < ALU; ++k)
      asm volatile("prmt.b32 %0, %1, %2, %0;" : "+r"(x[k & 3]) : "r"(b0), "r"(b1));
  }
  int s = x[0] + x[1] + x[2] + x[3];
  for (int k = 0; k < NACC; ++k)
    for (int j = 0; j < 4; ++j) s += acc[k][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC, int ALU>
void run_mix(int warps_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  int* out; cudaMalloc(&out, sizeof(int) * sms * warps_per_sm * 32);
  const int iters = 20000;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  mix_kernel<NACC, ALU><<<sms, warps_per_sm * 32>>>(out, 100, 1);
  cudaEventRecord(a);
  mix_kernel<NACC, ALU><<<sms, warps_per_sm * 32>>>(out, iters, 1);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double cyc = ms * 1e-3 * clk * 1e3 / ((double)iters * warps_per_sm / 4);   // per SMSP per iteration of one warp
  printf("%d IMMA + %2d PRMT  warps/SM=%2d  %.3f ms  %.1f cycles per warp-iteration per SMSP\n", NACC, ALU, warps_per_sm, ms, cyc);
  cudaFree(out);
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
  for (int w : {4, 8, 16}) run<2, 8>(w, "imma k16");
  for (int w : {4, 16}) run<1, 8>(w, "bmma b1");
  for (int w : {8, 16}) { run_mix<8, 0>(w); run_mix<1, 24>(w); run_mix<8, 24>(w); run_mix<8, 48>(w); }
  return 0;
}
