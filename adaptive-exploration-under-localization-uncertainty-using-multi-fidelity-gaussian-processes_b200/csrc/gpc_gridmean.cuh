// gpc_gridmean.cuh -- posterior MEAN on a tensor grid as FP64 tensor-core GEMMs.
//
// Every test set of the reference is a tensor grid (np.meshgrid, exploreSimSettings.py:116-119, the planner's
// fieldGrid), and the squared-exponential cross-covariance separates per axis:
//   k_m((a_x[i], a_y[j], a_z[k]), x_n) = var_m Tx_m[i][n] Ty_m[j][n] Tz_m[k][n],  T_d,m[i][n] = exp(-((a_d[i] - x_nd) / l_md)^2 / 2)
// so   mean[i][j][k] = sum_m sum_n (c_m[n] Tx_m[i][n] Ty_m[j][n]) Tz_m[k][n],   c_m[n] = alpha_n coef[fi][m] var_m coef[f_n][m]
// -- (nx + ny + nz) N exps per term instead of nx ny nz N, and the contraction over the training index is a
// (nx ny) x nz x N GEMM: 2 M N flop on the DMMA pipe instead of M N kernel evaluations on the FP64 ALUs
// ("the grid path lifts the mean to DGEMM-bound").  The products of exps differ from the exp of the
// sum by O(ulp).  replaces: gp.predict(grid)[0], NIGP.predict(grid, return_var=False) on meshgrid inputs.
#pragma once
#include "gpc_gemm.cuh"

// T[i][n] for one axis d and one term m; rows i >= cnt and columns n >= N are zero.  grid (n_pad / 256, rows_pad).
__global__ void __launch_bounds__(256) k_gm_table(const __grid_constant__ GpcHyp h, const double* __restrict__ Xt,
                                                  long n_pad, long N, const double* __restrict__ axis, int cnt, int d,
                                                  int m, double* __restrict__ T) {
  const long n = (long)blockIdx.x * 256 + threadIdx.x;
  if (n >= n_pad) return;
  const int i = blockIdx.y;
  double v = 0.0;
  if (i < cnt && n < N) {
    const double s = (Xt[(long)d * n_pad + n] - axis[i]) * (h.inv_l[m][d] * 0.70710678118654752440);
    v = gpc_exp_neg(s * s);
  }
  T[(long)i * n_pad + n] = v;
}

// c_m[n] = alpha_n coef[fi][m] var_m coef[f_n][m]  (zero for m > min(fi, f_n) and for padding columns)
__global__ void __launch_bounds__(256) k_gm_coef(const __grid_constant__ GpcHyp h, const double* __restrict__ Xt,
                                                 const double* __restrict__ alpha, long n_pad, long N, int fi, int m,
                                                 double* __restrict__ c) {
  const long n = (long)blockIdx.x * 256 + threadIdx.x;
  if (n >= n_pad) return;
  double v = 0.0;
  if (n < N) {
    const int fj = gpc_fid(h, Xt[3 * n_pad + n]);
    if (m <= fi && m <= fj) v = alpha[n] * h.coef[fi][m] * h.var[m] * h.coef[fj][m];
  }
  c[n] = v;
}

// A[r][n] = c[n] Tx[ix][n] Ty[iy][n] for the (ix, iy) pairs pair0 .. pair0 + R - 1 (pair = ix ny + iy); rows past the
// last pair are zero.  grid (n_pad / 256, R).
__global__ void __launch_bounds__(256) k_gm_form(const double* __restrict__ c, const double* __restrict__ Tx,
                                                 const double* __restrict__ Ty, long n_pad, long pair0, long npairs,
                                                 int ny, double* __restrict__ A) {
  const long n = (long)blockIdx.x * 256 + threadIdx.x;
  if (n >= n_pad) return;
  const long pair = pair0 + blockIdx.y;
  double v = 0.0;
  if (pair < npairs) {
    const long ix = pair / ny, iy = pair - ix * ny;
    v = c[n] * Tx[ix * n_pad + n] * Ty[iy * n_pad + n];
  }
  A[(long)blockIdx.y * n_pad + n] = v;
}

// C[r][k] = beta C[r][k] + sum_n A[r][n] Tz[k][n]: 64 x 64 tiles, the DMMA loop of the factorisation kernels.
// grid (nz_pad / 64, R / 64).
__global__ void __launch_bounds__(gpc64::NT, 4) k_gm_gemm(const double* __restrict__ A, const double* __restrict__ Tz,
                                                          long n_pad, double* __restrict__ C, long ldc, double beta) {
  extern __shared__ double sm[];
  double acc[4][4][2];
  gpcg::zero_acc(acc);
  gpc64::mainloop<false>(A + (long)blockIdx.y * 64 * n_pad, n_pad, Tz + (long)blockIdx.x * 64 * n_pad, n_pad, 0, (int)n_pad,
                         acc, sm);
  gpc64::store_tile(C + (long)blockIdx.y * 64 * ldc + (long)blockIdx.x * 64, ldc, acc, 1.0, beta);
}

// mean[(pair0 + r) nz + k] = C[r][k]
__global__ void __launch_bounds__(256) k_gm_store(const double* __restrict__ C, long ldc, long pair0, long npairs, int nz,
                                                  long R, double* __restrict__ mean) {
  const long total = R * nz;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long r = idx / nz, k = idx - r * nz;
    if (pair0 + r < npairs) mean[(pair0 + r) * nz + k] = C[r * ldc + k];
  }
}
