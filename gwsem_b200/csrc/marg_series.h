// Term tables of the series path of the WLS marginals statistics (marg_series.cu), shared by the host feature
// builder and the per-SNP assembly kernel.  A term (m, l) is the coefficient of rho^m u^l.
#pragma once

namespace gwsem {
namespace series {

constexpr int kOrderB = 3;      // order in rho of every weighted sum (the SNP-dependent rows of Muthen's A and B)
constexpr int kOrderRoot = 5;   // order of the unweighted score sum, whose root is the correlation estimate
constexpr int kL = 8;           // powers of u carried: u^0 .. u^7

// psi, psi dl_a, psi dl_b: the coefficient of rho^m has degree m + 1 in u and the parity of m + 1
constexpr int kPsiB = 8, kPsiAll = 15;
__host__ __device__ constexpr int psi_m(int p) { return p < 1 ? 0 : p < 3 ? 1 : p < 5 ? 2 : p < 8 ? 3 : p < 11 ? 4 : 5; }
__host__ __device__ constexpr int psi_l(int p) {
  return p == 0 ? 1 : p == 1 ? 2 : p == 2 ? 0 : p == 3 ? 3 : p == 4 ? 1 : p == 5 ? 4 : p == 6 ? 2 : p == 7 ? 0
       : p == 8 ? 5 : p == 9 ? 3 : p == 10 ? 1 : p == 11 ? 6 : p == 12 ? 4 : p == 13 ? 2 : 0;
}
// psi * (d log pi / du): orders 1..3, degree m, parity m
constexpr int kF3 = 5;
__host__ __device__ constexpr int f3_m(int p) { return p < 1 ? 1 : p < 3 ? 2 : 3; }
__host__ __device__ constexpr int f3_l(int p) { return p == 0 ? 1 : p == 1 ? 2 : p == 2 ? 0 : p == 3 ? 3 : 1; }
// psi^2: orders 0..3, degree m + 2, parity m
constexpr int kDiag = 10;
__host__ __device__ constexpr int diag_m(int p) { return p < 2 ? 0 : p < 4 ? 1 : p < 7 ? 2 : 3; }
__host__ __device__ constexpr int diag_l(int p) {
  return p == 0 ? 2 : p == 1 ? 0 : p == 2 ? 3 : p == 3 ? 1 : p == 4 ? 4 : p == 5 ? 2 : p == 6 ? 0 : p == 7 ? 5 : p == 8 ? 3 : 1;
}
// psi_a psi_b (a > b): rho_a^ma rho_b^mb u^l with ma + mb <= 3, degree ma + mb + 2, parity ma + mb
constexpr int kCross = 27;
struct CrossTerm { int ma, mb, l; };
__host__ __device__ inline CrossTerm cross_term(int p) {
  int tot = 0, base = 0;
  while (true) {
    const int nl = tot / 2 + 2, cnt = nl * (tot + 1);
    if (p < base + cnt) break;
    base += cnt;
    ++tot;
  }
  const int nl = tot / 2 + 2, q = p - base;
  CrossTerm t;
  t.ma = q / nl;
  t.mb = tot - t.ma;
  t.l = tot + 2 - 2 * (q % nl);
  return t;
}

}  // namespace series
}  // namespace gwsem
