// Genotype decode kernels: packed hardcalls / dosages -> the double column the reference's
// providers hand to OpenMx, bit for bit (reference src/loader.cpp:211-273, 386-394).
#include "common.h"
#include "kernels.h"

namespace gwsem {

__device__ __forceinline__ double na_real() { return __longlong_as_double((long long)GWSEM_NA_BITS); }

// One thread per (person, SNP row).  PLINK 2 codes: 0,1,2 = ALT count, 3 = missing;
// PLINK 1 .bed codes 00->2, 01->missing, 10->1, 11->0 (pgenlib_read.cc:1979-1992).
// dest_of[person] = output row or -1 (rowFilter compaction, loader.cpp:201-208).
// maf_acc[2*row] += sum of finite values (integer, exact), maf_acc[2*row+1] += count.
__global__ void decode_2bit_kernel(const uint8_t* __restrict__ packed, int64_t pitch, int n_samples,
                                   int plink1, const int32_t* __restrict__ dest_of, int dest_rows,
                                   double* __restrict__ out, unsigned long long* __restrict__ maf_acc) {
  const int row = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned value = 0, finite = 0;
  if (i < n_samples) {
    const int d = dest_of ? dest_of[i] : i;
    if (d >= 0) {
      unsigned code = (packed[row * pitch + (i >> 2)] >> (2 * (i & 3))) & 3u;
      if (plink1) code = (0x1Eu >> (2 * code)) & 3u;  // table {2,3,1,0} packed as 0b00011110
      finite = code != 3u;
      value = finite ? code : 0u;
      out[(int64_t)row * dest_rows + d] = finite ? (double)code : na_real();
    }
  }
  // warp-aggregate, then one atomic pair per warp
  unsigned packed_sum = value | (finite << 16);
#pragma unroll
  for (int o = 16; o; o >>= 1) packed_sum += __shfl_xor_sync(0xffffffffu, packed_sum, o);
  if ((threadIdx.x & 31) == 0 && packed_sum) {
    atomicAdd(&maf_acc[2 * row], (unsigned long long)(packed_sum & 0xffffu));
    atomicAdd(&maf_acc[2 * row + 1], (unsigned long long)(packed_sum >> 16));
  }
}

// maf = total / (2 * n), folded to <= 0.5 (loader.cpp:393-394).  `unit` scales the integer
// total back to allele counts (1 for hardcalls, 2^-14 for PGEN dosages).
__global__ void maf_finish_kernel(const unsigned long long* __restrict__ acc, int n_snps, double unit,
                                  double* __restrict__ maf) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_snps) return;
  double m = ((double)acc[2 * r] * unit) / (2.0 * (double)acc[2 * r + 1]);
  if (m > 0.5) m = 1 - m;
  maf[r] = m;
}

void launch_decode_2bit(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int n_samples,
                        int plink1, const int32_t* d_dest_of, int dest_rows, double* d_out, double* d_maf) {
  if (n_snps <= 0) return;
  e->d_maf_acc.ensure(2 * (size_t)n_snps);
  GW_CUDA(cudaMemsetAsync(e->d_maf_acc.p, 0, 2 * (size_t)n_snps * sizeof(unsigned long long), e->stream));
  const int threads = 256;
  for (int r0 = 0; r0 < n_snps; r0 += 65535) {  // gridDim.y limit
    const int rows = n_snps - r0 < 65535 ? n_snps - r0 : 65535;
    dim3 grid((n_samples + threads - 1) / threads, rows);
    decode_2bit_kernel<<<grid, threads, 0, e->stream>>>(d_packed + (int64_t)r0 * pitch, pitch, n_samples, plink1,
                                                        d_dest_of, dest_rows, d_out + (int64_t)r0 * dest_rows,
                                                        e->d_maf_acc.p + 2 * (size_t)r0);
    GW_CUDA(cudaGetLastError());
    e->launches++;
  }
  maf_finish_kernel<<<(n_snps + 255) / 256, 256, 0, e->stream>>>(e->d_maf_acc.p, n_snps, 1.0, d_maf);
  GW_CUDA(cudaGetLastError());
  e->launches++;
}

}  // namespace gwsem
