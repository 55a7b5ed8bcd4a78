// gpc_predict.cuh -- cross-covariance assembly fused with the posterior mean (and mean
// gradients), the L^-1 K* DMMA contraction with fused variance reduction, the full-covariance
// SYRK, and small element-wise epilogues.
#pragma once
#include "gpc_gemm.cuh"

// ------------------------------------------------------------------------------------------
// K* tile assembly + mean.  CTA = 256 threads = 8 warps; tile = 64 test rows x up to 1024 train
// columns.  The train chunk (x, y, z, fid, alpha as SoA) is staged into shared memory with five
// 1-D TMA bulk copies signalled on one mbarrier; each warp owns 8 test rows, lanes run along the
// train index so the K* row segments are written as coalesced 256-byte lines, and the mean
// (K* alpha) is accumulated per lane and warp-shuffle reduced once per row.
//   Kx        [m_pad][n_pad]   cross covariance (row = test point), zero for n >= M or j >= N
//   meanpart  [n_chunks][m_pad]   partial means per train chunk (summed in k_finalize_pred)
//   gradpart  [n_chunks][3][m_pad] optional: partial mean gradients (NIGP.py:55-64, 307-311)
// replaces: NIGP.py:292-293 (Kxs, Kxs @ alpha), GPy _raw_predict (Kx, Kx^T woodbury_vector).
// ------------------------------------------------------------------------------------------
constexpr int KS_ROWS = 64;
constexpr int KS_COLS = 1024;

template <bool WITH_GRAD, bool STORE_K>
__global__ void __launch_bounds__(256) k_kstar(const __grid_constant__ GpcHyp h, const double* __restrict__ Xt,
                                               const double* __restrict__ alpha, long N, long n_pad,
                                               const double* __restrict__ Xs4, long M, long m_pad,
                                               double* __restrict__ Kx, double* __restrict__ meanpart,
                                               double* __restrict__ gradpart) {
  __shared__ __align__(128) double tr[5][KS_COLS];
  __shared__ double ts[KS_ROWS][4];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long row0 = (long)blockIdx.x * KS_ROWS;
  const long j0 = (long)blockIdx.y * KS_COLS;
  const int ncol = (int)((n_pad - j0) < KS_COLS ? (n_pad - j0) : KS_COLS);  // multiple of 128

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)ncol * 8u;
    mbar_expect_tx(&bar, 5u * bytes);
#pragma unroll
    for (int c = 0; c < 4; ++c) tma_bulk_g2s(&tr[c][0], Xt + (long)c * n_pad + j0, bytes, &bar);
    tma_bulk_g2s(&tr[4][0], alpha + j0, bytes, &bar);
  }
  {
    const int r = tid >> 2, c = tid & 3;
    ts[r][c] = (row0 + r < M) ? Xs4[(row0 + r) * 4 + c] : 0.0;
  }
  __syncthreads();
  mbar_wait(&bar, 0);

  const double il0 = h.inv_l[0][0] * h.inv_l[0][0], il1 = h.inv_l[0][1] * h.inv_l[0][1],
               il2 = h.inv_l[0][2] * h.inv_l[0][2];
#pragma unroll 1
  for (int rr = 0; rr < 8; ++rr) {
    const int r = warp * 8 + rr;
    const long n = row0 + r;
    const double ax = ts[r][0], ay = ts[r][1], az = ts[r][2], af = ts[r][3];
    const bool live = n < M && af >= 0.0;  // fidelity < 0 marks a padding row (information-gain spans)
    double mu = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;
    double* krow = Kx + n * n_pad + j0;
#pragma unroll 2
    for (int j = lane; j < ncol; j += 32) {
      const double bx = tr[0][j], by = tr[1][j], bz = tr[2][j], bf = tr[3][j];
      double k = gpc_kval(h, ax, ay, az, af, bx, by, bz, bf);
      if (!live || j0 + j >= N) k = 0.0;
      if (STORE_K) krow[j] = k;
      const double ka = k * tr[4][j];
      mu += ka;
      if (WITH_GRAD) {
        g0 = fma(ka, (bx - ax) * il0, g0);
        g1 = fma(ka, (by - ay) * il1, g1);
        g2 = fma(ka, (bz - az) * il2, g2);
      }
    }
    mu = warp_sum(mu);
    if (WITH_GRAD) { g0 = warp_sum(g0); g1 = warp_sum(g1); g2 = warp_sum(g2); }
    if (lane == 0) {
      meanpart[(long)blockIdx.y * m_pad + n] = mu;
      if (WITH_GRAD) {
        double* gp = gradpart + (long)blockIdx.y * 3 * m_pad;
        gp[n] = g0; gp[m_pad + n] = g1; gp[2 * m_pad + n] = g2;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// THE HOT KERNEL:  Vt tile (test tile nt, train tile ib) = Kx[nt, 0:(ib+1)128] * X[ib, :]^T
// i.e. V = L^-1 K*^T evaluated as a triangular DMMA contraction with the explicit inverse
// factor.  Epilogue either reduces sum_i V(i, n)^2 over the tile (posterior variance; nothing but
// 128 doubles leaves the CTA) or stores the tile (full covariance / information gain).
// grid (m_pad / 128, nb); blockIdx.y = 0 is the heaviest train tile (longest k-range) so the
// hardware scheduler places long tiles first.
//   sumsq [nb][m_pad]
// replaces: NIGP.py:300-301 cho_solve(cho, Kxs.T), GPy dtrtrs + square().sum(0).
// ------------------------------------------------------------------------------------------
template <bool STORE_V, bool SUMSQ>
__global__ void __launch_bounds__(gpcg::NTHREADS, 1) k_vt(const double* __restrict__ Kx,
                                                          const double* __restrict__ X, long n_pad, int nb,
                                                          long m_pad, double* __restrict__ Vt,
                                                          double* __restrict__ sumsq) {
  extern __shared__ double sm[];
  const int nt = blockIdx.x, ib = nb - 1 - blockIdx.y;
  double acc[4][4][2];
  gpcg::zero_acc(acc);
  gpcg::mainloop<false>(Kx + (long)nt * 128 * n_pad, n_pad, X + (long)ib * 128 * n_pad, n_pad, 0, (ib + 1) * 128,
                        acc, sm);
  if (STORE_V) gpcg::store_tile(Vt + (long)nt * 128 * n_pad + (long)ib * 128, n_pad, acc, 1.0, 0.0);
  if (SUMSQ) {
    const double t = gpcg::rowsumsq_tile(acc, sm);
    if (threadIdx.x < 128) sumsq[(long)ib * m_pad + (long)nt * 128 + threadIdx.x] = t;
  }
}

// mean[n] = sum_c meanpart[c][n];  var[n] = kdiag(fid_n) - sum_ib sumsq[ib][n]  (+clip, +noise).
// With sx != NULL the NIGP test-input-noise term sum_d (d mean / d x_d)^2 sx_d^2 is added before
// the floor (NIGP.py:304-324); sx is (1 x 3) when sx_rows == 1, else (M x 3).
__global__ void __launch_bounds__(256) k_finalize_pred(const __grid_constant__ GpcHyp h,
                                                       const double* __restrict__ Xs4, long M, long m_pad,
                                                       const double* __restrict__ meanpart, int nchunks,
                                                       const double* __restrict__ sumsq, int nb,
                                                       const double* __restrict__ gradpart,
                                                       const double* __restrict__ sx, long sx_rows,
                                                       double* __restrict__ mean, double* __restrict__ var,
                                                       unsigned flags) {
  const long n = (long)blockIdx.x * 256 + threadIdx.x;
  if (n >= M) return;
  if (mean) {
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += meanpart[(long)c * m_pad + n];
    mean[n] = s;
  }
  if (var) {
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += sumsq[(long)b * m_pad + n];
    const int f = gpc_fid(h, Xs4[n * 4 + 3]);
    double v = h.kdiag[f] - s;
    if (flags & 2u) v = fmax(v, 1e-15);
    if (flags & 1u) v += h.noise[f];
    if (sx) {
      const double* sr = sx + (sx_rows == 1 ? 0 : n * 3);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        double g = 0.0;
        for (int c = 0; c < nchunks; ++c) g += gradpart[((long)c * 3 + d) * m_pad + n];
        v = fma(g * g, sr[d] * sr[d], v);
      }
    }
    if (flags & 8u) v = fmax(v + 1e-12, 1e-12);
    var[n] = v;
  }
}

// grads[n][d] = sum_c gradpart[c][d][n]   (M x 3 row-major output)
__global__ void __launch_bounds__(256) k_finalize_grad(const double* __restrict__ gradpart, int nchunks,
                                                       long M, long m_pad, double* __restrict__ grads) {
  const long n = (long)blockIdx.x * 256 + threadIdx.x;
  if (n >= M) return;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += gradpart[((long)c * 3 + d) * m_pad + n];
    grads[n * 3 + d] = s;
  }
}

// ------------------------------------------------------------------------------------------
// Full posterior covariance: cov(m, n) = k(xs_m, xs_n) - sum_i Vt(m, i) Vt(n, i)  (lower tiles,
// mirrored on store).  grid (mt, mt).  cov has leading dimension ldc; with pad_identity the rows /
// columns M .. m_pad-1 are written as an identity block (so the result can be factored in place).
// replaces NIGP.py:299-301, GPy `Kxx - tdot(tmp.T)`, emukit predict_covariance (+ element-wise clip).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(gpcg::NTHREADS, 1) k_cov(const __grid_constant__ GpcHyp h,
                                                           const double* __restrict__ Vt, long n_pad,
                                                           const double* __restrict__ Xs4, long M,
                                                           const double* __restrict__ extra_diag,
                                                           double* __restrict__ cov, long ldc, int pad_identity,
                                                           unsigned flags) {
  extern __shared__ double sm[];
  const int nt = blockIdx.x, mt = blockIdx.y;
  if (nt > mt) return;
  double acc[4][4][2];
  gpcg::zero_acc(acc);
  gpcg::mainloop<false>(Vt + (long)mt * 128 * n_pad, n_pad, Vt + (long)nt * 128 * n_pad, n_pad, 0, (int)n_pad, acc,
                        sm);
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const long m = (long)mt * 128 + gpcg::acc_row(f);
    double ax = 0, ay = 0, az = 0, af = 0;
    if (m < M) { ax = Xs4[m * 4]; ay = Xs4[m * 4 + 1]; az = Xs4[m * 4 + 2]; af = Xs4[m * 4 + 3]; }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const long n = (long)nt * 128 + gpcg::acc_col(g) + e;
        if (mt == nt && n > m) continue;  // diagonal tiles: lower half, mirrored below
        double v;
        if (m >= M || n >= M) {
          if (!pad_identity) continue;
          v = (m == n) ? 1.0 : 0.0;
        } else {
          v = gpc_kval(h, ax, ay, az, af, Xs4[n * 4], Xs4[n * 4 + 1], Xs4[n * 4 + 2], Xs4[n * 4 + 3]) -
              acc[f][g][e];
          if (m == n) {
            if (flags & 1u) v += h.noise[gpc_fid(h, af)];
            if (extra_diag) v += extra_diag[m];
            if (flags & 8u) v += 1e-12;
          }
          if (flags & 4u) v = fmax(v, 1e-10);
        }
        cov[m * ldc + n] = v;
        if (m != n) cov[n * ldc + m] = v;
      }
    }
  }
}

// Plain kernel matrix K(Xa, Xb) (no noise) -- NIGP.py:11-20, gpy_model.kern.K.
__global__ void __launch_bounds__(256) k_kernel_matrix(const __grid_constant__ GpcHyp h,
                                                       const double* __restrict__ Xa4, long na,
                                                       const double* __restrict__ Xb4, long nb_,
                                                       double* __restrict__ K) {
  const long j = (long)blockIdx.x * 32 + (threadIdx.x & 31);
  const long i = (long)blockIdx.y * 8 + (threadIdx.x >> 5);
  if (i >= na || j >= nb_) return;
  K[i * nb_ + j] = gpc_kval(h, Xa4[i * 4], Xa4[i * 4 + 1], Xa4[i * 4 + 2], Xa4[i * 4 + 3], Xb4[j * 4],
                            Xb4[j * 4 + 1], Xb4[j * 4 + 2], Xb4[j * 4 + 3]);
}
