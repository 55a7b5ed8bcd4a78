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
//
// AR1 sum without divergence: per row the warp builds the 4 x 4 table cw[m][fj] =
// coef[fi][m] var[m] coef[fj][m] (zero for m > min(fi, fj)); an element is then
// sum_{m <= mm} cw[m][fj] base_m(x - x'), with mm = min(fi, max fidelity of the 128-column
// block) -- warp-uniform, and exact (no wasted exp) because gpc_set_data sorts the training
// rows by fidelity.  Four columns per lane are evaluated together for instruction-level
// parallelism (the exp chains are the critical path: profiles/r01/README.md).
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
  __shared__ double cw[8][GPC_MAXF][GPC_MAXF];   // per warp
  __shared__ double hil[GPC_MAXF][4];            // inv lengthscales (pre-scaled), shared copy of the hypers
  __shared__ int bmax[KS_COLS / 128];
  __shared__ double T64[64];                     // 2^(j / 64): the table of gpc_exp_neg_tab_w
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 64) T64[tid] = exp2((double)tid * 0.015625);
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
  // exp(-r^2 / 2) = exp(-(sum_d (d_d c_d)^2)) with c_d = inv_l_d / sqrt(2); Matern keeps inv_l
  if (tid < GPC_MAXF * 3) {
    const int m = tid / 3, d = tid % 3;
    hil[m][d] = h.inv_l[m][d] * (h.base == 0 ? 0.70710678118654752440 : 1.0);
  }
  __syncthreads();
  mbar_wait(&bar, 0);
  // fidelity of every staged column as an int (in place: the double slot's low word), block maxima
  {
    int* fi32 = reinterpret_cast<int*>(&tr[3][0]);
    const int nblk = ncol >> 7;
    if (warp < nblk) {
      int mx = 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = warp * 128 + lane + 32 * u;
        const int f = gpc_fid(h, tr[3][j]);
        mx = f > mx ? f : mx;
        fi32[2 * j] = f;
      }
      mx = __reduce_max_sync(0xffffffffu, mx);
      if (lane == 0) bmax[warp] = mx;
    }
  }
  __syncthreads();
  const int* fi32 = reinterpret_cast<const int*>(&tr[3][0]);
  const int F = h.F, base = h.base;

#pragma unroll 1
  for (int rr = 0; rr < 8; ++rr) {
    const int r = warp * 8 + rr;
    const long n = row0 + r;
    const double ax = ts[r][0], ay = ts[r][1], az = ts[r][2], af = ts[r][3];
    const bool live = n < M && af >= 0.0;  // fidelity < 0 marks a padding row (information-gain spans)
    const int fi = gpc_fid(h, af);
    __syncwarp();
    if (lane < GPC_MAXF * GPC_MAXF) {
      const int m = lane >> 2, fj = lane & 3;
      cw[warp][m][fj] = (m <= fi && m <= fj && m < F && fj < F) ? h.coef[fi][m] * h.var[m] * h.coef[fj][m] : 0.0;
    }
    __syncwarp();
    double mu = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;
    double* krow = Kx + n * n_pad + j0;
#pragma unroll 1
    for (int cb = 0; cb < (ncol >> 7); ++cb) {
      const int mm = fi < bmax[cb] ? fi : bmax[cb];
      double dx[4], dy[4], dz[4], kv[4];
      int fj[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = cb * 128 + lane + 32 * u;
        dx[u] = tr[0][j] - ax;
        dy[u] = tr[1][j] - ay;
        dz[u] = tr[2][j] - az;
        fj[u] = fi32[2 * j];
        kv[u] = 0.0;
      }
#pragma unroll 1
      for (int m = 0; m <= mm; ++m) {
        const double c0 = hil[m][0], c1 = hil[m][1], c2 = hil[m][2];
        double q[4], e[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double sx = dx[u] * c0, sy = dy[u] * c1, sz = dz[u] * c2;
          q[u] = fma(sx, sx, fma(sy, sy, sz * sz));
        }
        if (base == 0) {
          gpc_exp_neg_tab_w<4>(q, e, T64);      // branch-free, the four columns of the lane advance together
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) q[u] = 1.7320508075688772 * sqrt(q[u]);
          gpc_exp_neg_tab_w<4>(q, e, T64);
#pragma unroll
          for (int u = 0; u < 4; ++u) e[u] *= 1.0 + q[u];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) kv[u] = fma(cw[warp][m][fj[u]], e[u], kv[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = cb * 128 + lane + 32 * u;
        double k = kv[u];
        if (!live || j0 + j >= N) k = 0.0;
        if (STORE_K) krow[j] = k;
        const double ka = k * tr[4][j];
        mu += ka;
        if (WITH_GRAD) {  // single-fidelity squared exponential: d k / d a_d = k (b_d - a_d) / l_d^2
          g0 = fma(ka, dx[u], g0);
          g1 = fma(ka, dy[u], g1);
          g2 = fma(ka, dz[u], g2);
        }
      }
    }
    mu = warp_sum(mu);
    if (WITH_GRAD) { g0 = warp_sum(g0); g1 = warp_sum(g1); g2 = warp_sum(g2); }
    if (lane == 0) {
      meanpart[(long)blockIdx.y * m_pad + n] = mu;
      if (WITH_GRAD) {
        double* gp = gradpart + (long)blockIdx.y * 3 * m_pad;
        gp[n] = g0 * h.inv_l[0][0] * h.inv_l[0][0];
        gp[m_pad + n] = g1 * h.inv_l[0][1] * h.inv_l[0][1];
        gp[2 * m_pad + n] = g2 * h.inv_l[0][2] * h.inv_l[0][2];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// THE HOT KERNEL:  V = L^-1 K*^T evaluated as a triangular DMMA contraction with the explicit
// inverse factor X:  Vt(test tile mt, train tile jb) = Kx[mt, 0:(jb+1)64] * X[jb, 0:(jb+1)64]^T.
// Epilogue either reduces sum_i V(i, n)^2 over the tile (posterior variance; nothing but 64
// doubles leaves the CTA) or stores the tile (full covariance / information gain).
//
// Shape (measured on B200, profiles/r01/gemm_variants_r01.txt): 64 x 64 CTA tiles, 4 warps of
// 32 x 32, BK = 16, 2 cp.async stages, 4 CTAs per SM -- 35.0 TFLOP/s on a square problem against
// 30.2 for the 128 x 128 / 16-warp / 1-CTA shape: many small independent CTAs hide each other's
// barrier and pipeline-fill bubbles, which one big CTA cannot.
// Schedule: CTA (p, mt) computes train tile nb2-1-p and then train tile p of the same test tile,
// so every CTA executes nb2+1 k-units (equal work, no tail), and the CTAs that share a K* row
// block are adjacent in launch order (they hit it in L2 instead of re-reading HBM).
// On the diagonal tile X is lower triangular: a warp skips the k-steps that only meet zeros.
//   sumsq [nb2][m_pad]
// replaces: NIGP.py:300-301 cho_solve(cho, Kxs.T), GPy dtrtrs + square().sum(0).
// ------------------------------------------------------------------------------------------
namespace gpvt {
constexpr int BM = 64, BN = 64, BK = 16, NT = 128, LDT = 20;
constexpr int STAGE = (BM + BN) * LDT;                 // doubles per stage
constexpr int SMEM_BYTES = (2 * STAGE + 2 * 64) * 8;  // two stages + row-sum scratch = 41984
}  // namespace gpvt

template <bool STORE_V, bool SUMSQ>
__global__ void __launch_bounds__(gpvt::NT, 4) k_vt(const double* __restrict__ Kx, const double* __restrict__ X,
                                                    long ld, int nb2, long m_pad, double* __restrict__ Vt,
                                                    double* __restrict__ sumsq) {
  using namespace gpvt;
  extern __shared__ double sm[];
  double* scratch = sm + 2 * STAGE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1, lr = lane >> 2, lc = lane & 3;
  const int mt = blockIdx.y;
  const int jb1 = nb2 - 1 - (int)blockIdx.x, jb2 = blockIdx.x;
  const int n1 = (jb1 + 1) * (BN / BK), n2 = (jb2 < jb1) ? (jb2 + 1) * (BN / BK) : 0;
  const int total = n1 + n2;
  const double* Ab = Kx + (long)mt * BM * ld;

  auto load = [&](int step) {
    const int jb = step < n1 ? jb1 : jb2;
    const int k0 = (step < n1 ? step : step - n1) * BK;
    double* As = sm + (step & 1) * STAGE;
    double* Bs = As + BM * LDT;
    const double* Bb = X + (long)jb * BN * ld;
#pragma unroll
    for (int c = tid; c < BM * (BK / 2); c += NT) {
      const int r = c >> 3, q = c & 7;
      cp_async16(As + r * LDT + 2 * q, Ab + (long)r * ld + k0 + 2 * q);
    }
#pragma unroll
    for (int c = tid; c < BN * (BK / 2); c += NT) {
      const int r = c >> 3, q = c & 7;
      cp_async16(Bs + r * LDT + 2 * q, Bb + (long)r * ld + k0 + 2 * q);
    }
  };

  double acc[4][4][2];
  gpcg::zero_acc(acc);
  load(0);
  cp_async_commit();
  for (int it = 0; it < total; ++it) {
    cp_async_wait<0>();
    __syncthreads();
    if (it + 1 < total) load(it + 1);
    cp_async_commit();
    // position inside the current segment; the last BN/BK steps of a segment are the diagonal tile
    const int seg_n = it < n1 ? n1 : n2;
    const int seg_it = it < n1 ? it : it - n1;
    const int jd = seg_it - (seg_n - BN / BK);  // >= 0 on the diagonal tile: k-step index inside it
    if (jd < 2 * (wn + 1)) {
      const double* As = sm + (it & 1) * STAGE;
      const double* Ap = As + (wm * 32 + lr) * LDT + lc;
      const double* Bp = As + BM * LDT + (wn * 32 + lr) * LDT + lc;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        double a[4], b[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) a[f] = Ap[f * 8 * LDT + kk * 4];
#pragma unroll
        for (int g = 0; g < 4; ++g) b[g] = Bp[g * 8 * LDT + kk * 4];
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
          for (int g = 0; g < 4; ++g) dmma884(acc[f][g][0], acc[f][g][1], a[f], b[g]);
      }
    }
    if (seg_it == seg_n - 1) {  // segment finished: flush the accumulators
      const int jb = it < n1 ? jb1 : jb2;
      if (STORE_V) {
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          double* row = Vt + ((long)mt * BM + wm * 32 + f * 8 + lr) * ld + (long)jb * BN + wn * 32 + 2 * lc;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<double2*>(row + g * 8) = make_double2(acc[f][g][0], acc[f][g][1]);
        }
      }
      if (SUMSQ) {
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          double s = 0.0;
#pragma unroll
          for (int g = 0; g < 4; ++g) s = fma(acc[f][g][0], acc[f][g][0], fma(acc[f][g][1], acc[f][g][1], s));
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          if (lc == 0) scratch[wn * 64 + wm * 32 + f * 8 + lr] = s;
        }
        __syncthreads();
        if (tid < 64) sumsq[(long)jb * m_pad + (long)mt * BM + tid] = scratch[tid] + scratch[64 + tid];
        // scratch is next written at the end of the following segment, many barriers later
      }
      gpcg::zero_acc(acc);
    }
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
