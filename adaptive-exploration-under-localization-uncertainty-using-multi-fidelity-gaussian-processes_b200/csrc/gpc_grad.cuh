// gpc_grad.cuh -- analytic gradients of the negative log marginal likelihood on the device.
//
//   d NLML / d theta = 1/2 sum_ij W_ij dK_ij / d theta,   W = Ky^-1 - alpha alpha^T,
//   Ky^-1 = X^T X with X = L^-1  (one SYRK-shaped DMMA contraction, N^3/3 flops).
// The reference fits hyper-parameters with numerical gradients (SciPy L-BFGS-B in NIGP.py:231-239,
// ~2D+3 objective evaluations per iteration) or inside GPy (GPTrainers.py:68,84,94); here one
// gradient costs about half a factorisation more than the objective itself.
//
//   k_transpose      Xt = X^T                      (tile transpose through shared memory)
//   k_kinv           P(i, j) = sum_{k >= i} Xt(i, k) Xt(j, k)   for tiles it >= jt
//   k_nlml_grad      per-tile partial sums of W_ij dK_ij/d{var_m, l_md, rho_l}; diag(W) on the way
//   k_reduce_partial column sums of the per-tile partials
#pragma once
#include "gpc_gemm.cuh"

#define GPC_NGK (4 * GPC_MAXF + (GPC_MAXF - 1))  // kernel-parameter slots: (var, l x 3) per fidelity, then rho

struct GpcGradTab {
  // d coef[i][m] / d rho_l  (0 unless m <= l < i)
  double dcoef[GPC_MAXF - 1][GPC_MAXF][GPC_MAXF];
};

__global__ void __launch_bounds__(256) k_transpose(const double* __restrict__ A, double* __restrict__ At, long ld) {
  __shared__ double t[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = ty; r < 32; r += 8) t[r][tx] = A[(long)(by + r) * ld + bx + tx];
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) At[(long)(bx + r) * ld + by + tx] = t[tx][r];
}

__global__ void __launch_bounds__(gpcg::NTHREADS, 1) k_kinv(const double* __restrict__ Xt, long ld, int nb,
                                                            double* __restrict__ P) {
  extern __shared__ double sm[];
  const int jt = blockIdx.x, it = nb - 1 - blockIdx.y;  // small it = long k-range first
  if (jt > it) return;
  double acc[4][4][2];
  gpcg::zero_acc(acc);
  gpcg::mainloop<false>(Xt + (long)it * 128 * ld, ld, Xt + (long)jt * 128 * ld, ld, it * 128, nb * 128, acc, sm);
  gpcg::store_tile(P + (long)it * 128 * ld + (long)jt * 128, ld, acc, 1.0, 0.0);
}

// grid (nb, nb) over lower tiles, 256 threads.  partial[(it * nb + jt) * GPC_NGK + slot].
__global__ void __launch_bounds__(256) k_nlml_grad(const __grid_constant__ GpcHyp h,
                                                   const __grid_constant__ GpcGradTab tab,
                                                   const double* __restrict__ Xt, const double* __restrict__ alpha,
                                                   const double* __restrict__ P, long N, long n_pad, int nb,
                                                   double* __restrict__ partial, double* __restrict__ diagW) {
  const int jt = blockIdx.x, it = blockIdx.y;
  double* out = partial + ((long)it * nb + jt) * GPC_NGK;
  __shared__ double red[8][GPC_NGK];
  double g[GPC_NGK];
#pragma unroll
  for (int s = 0; s < GPC_NGK; ++s) g[s] = 0.0;
  if (jt <= it) {
    const double *xs = Xt, *ys = Xt + n_pad, *zs = Xt + 2 * n_pad, *fs = Xt + 3 * n_pad;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int F = h.F;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const long j = (long)jt * 128 + tx + 32 * c;
      const double bx = xs[j], by = ys[j], bz = zs[j], aj = alpha[j];
      const int fj = gpc_fid(h, fs[j]);
#pragma unroll 1
      for (int r = 0; r < 16; ++r) {
        const long i = (long)it * 128 + ty + 8 * r;
        if (i >= N || j >= N || j > i) continue;
        const double w0 = P[i * n_pad + j] - alpha[i] * aj;
        if (i == j) diagW[i] = w0;
        const double w = (i == j) ? w0 : 2.0 * w0;  // the strict upper triangle mirrors the lower one
        const int fi = gpc_fid(h, fs[i]);
        const double dx = xs[i] - bx, dy = ys[i] - by, dz = zs[i] - bz;
        const int mm = fi < fj ? fi : fj;
        for (int m = 0; m <= mm && m < F; ++m) {
          const double i0 = h.inv_l[m][0], i1 = h.inv_l[m][1], i2 = h.inv_l[m][2];
          const double sx = dx * i0, sy = dy * i1, sz = dz * i2;
          const double r2 = fma(sx, sx, fma(sy, sy, sz * sz));
          double e, dl;  // e = k_m / var_m;  dl = factor so that d k_m / d l_md = var_m dl s_d^2 / l_md
          if (h.base == 0) {
            e = exp(-0.5 * r2);
            dl = e;
          } else {
            const double rt = 1.7320508075688772 * sqrt(r2);
            const double ex = exp(-rt);
            e = (1.0 + rt) * ex;
            dl = 3.0 * ex;
          }
          const double cc = h.coef[fi][m] * h.coef[fj][m];
          const double wk = w * cc;
          g[4 * m] = fma(wk, e, g[4 * m]);
          const double wl = wk * h.var[m] * dl;
          g[4 * m + 1] = fma(wl, sx * sx * i0, g[4 * m + 1]);
          g[4 * m + 2] = fma(wl, sy * sy * i1, g[4 * m + 2]);
          g[4 * m + 3] = fma(wl, sz * sz * i2, g[4 * m + 3]);
          const double wke = w * h.var[m] * e;
          for (int l = m; l < F - 1; ++l)
            g[4 * GPC_MAXF + l] = fma(wke, tab.dcoef[l][fi][m] * h.coef[fj][m] + h.coef[fi][m] * tab.dcoef[l][fj][m],
                                      g[4 * GPC_MAXF + l]);
        }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < GPC_NGK; ++s) {
    const double v = warp_sum(g[s]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][s] = v;
  }
  __syncthreads();
  if (threadIdx.x < GPC_NGK) {
    double v = 0.0;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) v += red[wv][threadIdx.x];
    out[threadIdx.x] = v;
  }
}

// out[s] = 1/2 sum_b partial[b * GPC_NGK + s]
__global__ void __launch_bounds__(256) k_reduce_partial(const double* __restrict__ partial, long nblocks,
                                                        double* __restrict__ out) {
  __shared__ double s0[8];
  const int s = blockIdx.x;
  double a = 0.0;
  for (long b = threadIdx.x; b < nblocks; b += 256) a += partial[b * GPC_NGK + s];
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) s0[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s0[w];
    out[s] = 0.5 * t;
  }
}
