// Device-side building blocks shared by the per-SNP fit kernels (fit.cu, marg_fit.cu): group
// synchronisation helpers, small dense linear algebra in shared memory, RAM model evaluation.
#pragma once
#include "common.h"
#include "fit_program.h"

namespace gwsem {

struct WarpScratch {
  // offsets in doubles into the per-warp shared-memory slab
  int R, mu, c2, sobs, L, theta, cand, A, S, Z, C, T, sig, r, rc, J, g, H, Hc, delta, total;
};

__host__ __device__ inline WarpScratch scratch_layout(int p, int q, int nv, int P, int nf) {
  WarpScratch w;
  int o = 0;
  auto take = [&](int n) { int at = o; o += n; return at; };
  w.R = take(5 * nf);
  w.mu = take(p);
  w.c2 = take(q);
  w.sobs = take(q);
  w.L = take(q * q);
  w.theta = take(P);
  w.cand = take(P);
  w.A = take(nv * nv);
  w.S = take(nv * nv);
  w.Z = take(nv * nv);
  w.C = take(nv * nv);
  w.T = take(nv * nv);
  w.sig = take(q);
  w.r = take(q);
  w.rc = take(q);
  w.J = take(q * P);
  w.g = take(P);
  w.H = take(P * P);
  w.Hc = take(P * P);
  w.delta = take(P);
  w.total = o;
  return w;
}

// A group of LANES threads works on one SNP: one thread, one warp, or (LANES > 32) the whole CTA.
template <int LANES>
__device__ __forceinline__ void gsync() {
  if (LANES > 32) __syncthreads();
  else if (LANES > 1) __syncwarp();
}
template <int LANES>
__device__ __forceinline__ double gbcast(double v) {   // the value of the group's first thread
  if (LANES > 32) {
    __shared__ double slot;
    __syncthreads();
    if (threadIdx.x == 0) slot = v;
    __syncthreads();
    return slot;
  }
  return LANES > 1 ? __shfl_sync(0xffffffffu, v, 0) : v;
}
template <int LANES>
__device__ __forceinline__ bool gany(bool p) {
  if (LANES > 32) return __syncthreads_or(p) != 0;
  return LANES > 1 ? __any_sync(0xffffffffu, p) != 0 : p;
}

__device__ __forceinline__ int vech_index(int i, int j, int p) {  // i >= j, column-major lower triangle
  return j * p - j * (j - 1) / 2 + (i - j);
}

// In-place Cholesky (lower) of the n x n row-major matrix a; returns false if not positive
// definite.  A pivot below 1e-10 of its natural scale (scale[j] if given, else the diagonal
// entry itself) counts as zero, so exactly singular matrices are reported, not inverted.
template <int LANES>
__device__ bool warp_cholesky(double* a, int n, int lane, const double* scale = nullptr) {
  for (int j = 0; j < n; ++j) {
    double d = 0.0, d0 = 0.0;
    if (lane == 0) {
      d0 = d = a[j * n + j];
      if (scale) d0 = scale[j];
      for (int k = 0; k < j; ++k) d -= a[j * n + k] * a[j * n + k];
    }
    d = gbcast<LANES>(d);
    d0 = gbcast<LANES>(d0);
    if (!(d > 1e-10 * d0) || !isfinite(d)) return false;
    const double dj = sqrt(d);
    if (lane == 0) a[j * n + j] = dj;
    for (int i = j + 1 + lane; i < n; i += LANES) {
      double v = a[i * n + j];
      for (int k = 0; k < j; ++k) v -= a[i * n + k] * a[j * n + k];
      a[i * n + j] = v / dj;
    }
    gsync<LANES>();
  }
  return true;
}

// Solve L y = b in place for `ncol` right-hand sides stored as columns of b (row-major n x ld).
template <int LANES>
__device__ void warp_forward_solve(const double* L, int n, double* b, int ld, int ncol, int lane) {
  for (int c = lane; c < ncol; c += LANES) {
    for (int i = 0; i < n; ++i) {
      double v = b[i * ld + c];
      for (int k = 0; k < i; ++k) v -= L[i * n + k] * b[k * ld + c];
      b[i * ld + c] = v / L[i * n + i];
    }
  }
  gsync<LANES>();
}

template <int LANES>
struct ModelEval {
  const FitProgram& fp;
  const WarpScratch& ws;
  double* sm;
  int lane;

  // A, S <- theta;  Z = (I-A)^-1;  C = Z S Z';  sig = vech(C[:p,:p])
  __device__ void implied(const double* theta, double* sig) {
    const int nv = fp.nv, p = fp.p;
    double *A = sm + ws.A, *S = sm + ws.S, *Z = sm + ws.Z, *C = sm + ws.C, *T = sm + ws.T;
    for (int e = lane; e < nv * nv; e += LANES) { A[e] = 0.0; S[e] = 0.0; Z[e] = (e / nv == e % nv) ? 1.0 : 0.0; }
    gsync<LANES>();
    for (int e = lane; e < fp.n_cells; e += LANES) {
      const int m = fp.cell_mat[e], r = fp.cell_row[e], c = fp.cell_col[e], k = fp.cell_par[e];
      const double v = k >= 0 ? theta[k] : fp.cell_val[e];
      if (m == 0) A[r * nv + c] = v;
      else if (m == 1) S[r * nv + c] = v;
    }
    gsync<LANES>();
    // Z = I + A + A^2 + ...   (A is nilpotent for recursive path models; nv terms suffice)
    for (int e = lane; e < nv * nv; e += LANES) T[e] = A[e];
    gsync<LANES>();
    for (int it = 0; it < nv; ++it) {
      bool nz = false;
      for (int e = lane; e < nv * nv; e += LANES) { Z[e] += T[e]; nz |= (T[e] != 0.0); }
      if (!gany<LANES>(nz)) break;
      gsync<LANES>();
      // C (temp) = T * A
      for (int e = lane; e < nv * nv; e += LANES) {
        const int i = e / nv, j = e % nv;
        double v = 0.0;
        for (int k = 0; k < nv; ++k) v += T[i * nv + k] * A[k * nv + j];
        C[e] = v;
      }
      gsync<LANES>();
      for (int e = lane; e < nv * nv; e += LANES) T[e] = C[e];
      gsync<LANES>();
    }
    gsync<LANES>();
    // T = Z S ; C = T Z'
    for (int e = lane; e < nv * nv; e += LANES) {
      const int i = e / nv, j = e % nv;
      double v = 0.0;
      for (int k = 0; k < nv; ++k) v += Z[i * nv + k] * S[k * nv + j];
      T[e] = v;
    }
    gsync<LANES>();
    for (int e = lane; e < nv * nv; e += LANES) {
      const int i = e / nv, j = e % nv;
      double v = 0.0;
      for (int k = 0; k < nv; ++k) v += T[i * nv + k] * Z[j * nv + k];
      C[e] = v;
    }
    gsync<LANES>();
    for (int e = lane; e < fp.q; e += LANES) sig[e] = C[fp.stat_i[e] * nv + fp.stat_j[e]];
    gsync<LANES>();
  }

  // J[e][k] = d sig_e / d theta_k (uses Z, C left by implied())
  __device__ void jacobian(double* J) {
    const int nv = fp.nv, P = fp.P;
    const double *Z = sm + ws.Z, *C = sm + ws.C;
    for (int idx = lane; idx < fp.q * P; idx += LANES) {
      const int e = idx / P, k = idx % P;
      const int i = fp.stat_i[e], j = fp.stat_j[e];
      double v = 0.0;
      for (int ce = fp.par_cell_ptr[k]; ce < fp.par_cell_ptr[k + 1]; ++ce) {
        const int c_id = fp.par_cells[ce];
        const int m = fp.cell_mat[c_id], r = fp.cell_row[c_id], c = fp.cell_col[c_id];
        if (m == 0) v += Z[i * nv + r] * C[c * nv + j] + Z[j * nv + r] * C[c * nv + i];
        else if (m == 1) v += Z[i * nv + r] * Z[j * nv + c];
      }
      J[idx] = v;
    }
    gsync<LANES>();
  }
};

}  // namespace gwsem
