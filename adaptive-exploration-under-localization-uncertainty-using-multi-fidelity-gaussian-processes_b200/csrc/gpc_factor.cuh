// gpc_factor.cuh -- covariance assembly, blocked right-looking FP64 Cholesky, triangular
// inverse by recursive doubling, and the alpha / log-det solves.
//
// Storage: all N x N operands are row-major with leading dimension n_pad (N rounded up to 128);
// only tiles on or below the block diagonal are ever touched.  Rows/columns >= N are an identity
// block, so the padded factor is [[L, 0], [0, I]] and contributes nothing to log-det or solves.
#pragma once
#include "gpc_gemm.cuh"

// ------------------------------------------------------------------------------------------
// K + diag(noise) for the training set.  Xt is SoA: x[n_pad], y[n_pad], z[n_pad], f[n_pad].
// replaces: NIGP.py:41-42,150-151,285-287; GPy exact_gaussian_inference (K + (noise+1e-8) I).
// grid (nb, nb) over lower tiles, 256 threads.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_assemble_train(const __grid_constant__ GpcHyp h,
                                                        const double* __restrict__ Xt,
                                                        const double* __restrict__ extra, double* __restrict__ K,
                                                        long N, long n_pad) {
  const int jb = blockIdx.x, ib = blockIdx.y;
  if (jb > ib) return;
  const double *xs = Xt, *ys = Xt + n_pad, *zs = Xt + 2 * n_pad, *fs = Xt + 3 * n_pad;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    const long j = (long)jb * 128 + tx + 32 * c;
    const double bx = xs[j], by = ys[j], bz = zs[j], bf = fs[j];
#pragma unroll 4
    for (int r = 0; r < 16; ++r) {
      const long i = (long)ib * 128 + ty + 8 * r;
      double v;
      if (i >= N || j >= N) {
        v = (i == j) ? 1.0 : 0.0;
      } else {
        v = gpc_kval(h, xs[i], ys[i], zs[i], fs[i], bx, by, bz, bf);
        if (i == j) v += h.noise[gpc_fid(h, bf)] + h.jitter + (extra ? extra[i] : 0.0);
      }
      K[i * n_pad + j] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Diagonal block p: L_pp = chol(A_pp) and X_pp = L_pp^-1, one CTA of 256 threads working out of
// shared memory.  Blocked right-looking elimination with 16-column sub-blocks:
//   (1) warp 0 factors the 16 x 16 diagonal sub-block in registers (one row per lane, pivots and
//       multipliers exchanged by warp shuffles) and inverts it by forward substitution;
//   (2) the rows below are solved against that inverse (one thread per row);
//   (3) the trailing sub-matrix gets its rank-16 update in 4 x 4 register micro-tiles.
// The inverse X_pp is then assembled block row by block row from the 16 x 16 diagonal inverses
// (X_rc = -T_r sum_m L_rm X_mc).  X^T lives in the unused upper triangle of the same shared
// array (leading dimension 129 leaves room for the shifted diagonal).
// status receives p + 1 for the first non-positive pivot.
// replaces: scipy cho_factor (NIGP.py:43,154,288) / LAPACK dpotrf inside GPy pdinv.
// ------------------------------------------------------------------------------------------
#define GPC_PD_LD 129
constexpr int GPC_POTRF_SMEM = (128 * GPC_PD_LD + 7 * 256) * 8;

__global__ void __launch_bounds__(256, 1) k_potrf_diag(double* __restrict__ A, double* __restrict__ X, long ld,
                                                       int p, int* __restrict__ status) {
  extern __shared__ double sm[];
  double* S = sm;                       // L in the lower triangle (incl. diagonal)
  double* W = sm + 128 * GPC_PD_LD;     // 7 x (16 x 16) scratch for the inverse
#define XT(i, j) S[(j) * GPC_PD_LD + (i) + 1]  // X(i, j), i >= j, stored transposed above the diagonal
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* Ap = A + (long)p * 128 * ld + (long)p * 128;
  double* Xp = X + (long)p * 128 * ld + (long)p * 128;
  __shared__ int bad;
  if (tid == 0) bad = 0;
  for (int e = tid; e < 128 * 128; e += 256) {
    const int r = e >> 7, c = e & 127;
    if (c <= r) S[r * GPC_PD_LD + c] = Ap[(long)r * ld + c];
  }
  __syncthreads();

#pragma unroll 1
  for (int kb = 0; kb < 8; ++kb) {
    const int c0 = kb * 16;
    if (warp == 0) {
      const int l = lane & 15;  // lanes 16..31 mirror lanes 0..15 (keeps every shuffle full-warp)
      double a[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = (j <= l) ? S[(c0 + l) * GPC_PD_LD + c0 + j] : 0.0;
      bool isbad = false;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        double d = __shfl_sync(0xffffffffu, a[k], k);
        if (!(d > 0.0)) { isbad = true; d = 1.0; }
        const double dd = sqrt(d), inv = 1.0 / dd;
        if (l == k) a[k] = dd;
        else if (l > k) a[k] *= inv;
#pragma unroll
        for (int j = k + 1; j < 16; ++j) {
          const double ljk = __shfl_sync(0xffffffffu, a[k], j);
          if (l >= j) a[j] = fma(-a[k], ljk, a[j]);
        }
      }
      if (isbad && lane == 0) bad = 1;
      // inverse of the 16 x 16 factor: lane l computes column l of T
      double x[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        double sacc = (i == l) ? 1.0 : 0.0;
#pragma unroll
        for (int m = 0; m < i; ++m) sacc = fma(-__shfl_sync(0xffffffffu, a[m], i), x[m], sacc);
        x[i] = sacc / __shfl_sync(0xffffffffu, a[i], i);
      }
      if (lane < 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j <= l) S[(c0 + l) * GPC_PD_LD + c0 + j] = a[j];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (i >= l) XT(c0 + i, c0 + l) = x[i];
      }
    }
    __syncthreads();
    const int r0 = c0 + 16, nrow = 128 - r0;
    // (2) panel: L[r][c0 + c] = sum_{m <= c} A[r][c0 + m] T[c][m]
    if (tid < nrow) {
      const int r = r0 + tid;
      double av[16], out[16];
#pragma unroll
      for (int m = 0; m < 16; ++m) av[m] = S[r * GPC_PD_LD + c0 + m];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        double sacc = 0.0;
#pragma unroll
        for (int m = 0; m <= c; ++m) sacc = fma(av[m], XT(c0 + c, c0 + m), sacc);
        out[c] = sacc;
      }
#pragma unroll
      for (int c = 0; c < 16; ++c) S[r * GPC_PD_LD + c0 + c] = out[c];
    }
    __syncthreads();
    // (3) trailing update, 4 x 4 micro-tiles of the lower triangle
    const int nt = nrow >> 2;
    for (int idx = tid; idx < nt * nt; idx += 256) {
      const int ti = idx / nt, tj = idx - ti * nt;
      if (tj > ti) continue;
      const double* Pi = S + (r0 + 4 * ti) * GPC_PD_LD + c0;
      const double* Pj = S + (r0 + 4 * tj) * GPC_PD_LD + c0;
      double c[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) c[u][v] = 0.0;
#pragma unroll 4
      for (int m = 0; m < 16; ++m) {
        double ai[4], aj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { ai[u] = Pi[u * GPC_PD_LD + m]; aj[u] = Pj[u * GPC_PD_LD + m]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) c[u][v] = fma(ai[u], aj[v], c[u][v]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int i = r0 + 4 * ti + u, j = r0 + 4 * tj + v;
          if (j <= i) S[i * GPC_PD_LD + j] -= c[u][v];
        }
    }
    __syncthreads();
  }
  if (bad && tid == 0) atomicCAS(status, 0, p + 1);

  // X = L^-1 block row by block row:  X_rc = -T_r (sum_{m = c}^{r-1} L_rm X_mc),  T_r = X_rr
#pragma unroll 1
  for (int r = 1; r < 8; ++r) {
    for (int e = tid; e < r * 256; e += 256) {
      const int c = e >> 8, i = (e >> 4) & 15, j = e & 15;
      const double* Lrow = S + (16 * r + i) * GPC_PD_LD;
      double sacc = 0.0;
      for (int t = j; t < 16; ++t) sacc = fma(Lrow[16 * c + t], XT(16 * c + t, 16 * c + j), sacc);  // m == c: X_cc lower
      for (int k = 16 * (c + 1); k < 16 * r; ++k) sacc = fma(Lrow[k], XT(k, 16 * c + j), sacc);
      W[e] = sacc;
    }
    __syncthreads();
    for (int e = tid; e < r * 256; e += 256) {
      const int c = e >> 8, i = (e >> 4) & 15, j = e & 15;
      double sacc = 0.0;
      for (int t = 0; t <= i; ++t) sacc = fma(XT(16 * r + i, 16 * r + t), W[(c << 8) + (t << 4) + j], sacc);
      XT(16 * r + i, 16 * c + j) = -sacc;
    }
    __syncthreads();
  }
  for (int e = tid; e < 128 * 128; e += 256) {
    const int r = e >> 7, c = e & 127;
    Ap[(long)r * ld + c] = (c <= r) ? S[r * GPC_PD_LD + c] : 0.0;
    Xp[(long)r * ld + c] = (c <= r) ? XT(r, c) : 0.0;
  }
#undef XT
}

// ------------------------------------------------------------------------------------------
// Panel solve: L_ip = A_ip * X_pp^T  for i > p (TRSM expressed as a DMMA contraction with the
// inverted diagonal block).  grid = nb - p - 1.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(gpcg::NTHREADS, 1) k_trsm_panel(double* __restrict__ A,
                                                                  const double* __restrict__ X, long ld, int p) {
  extern __shared__ double sm[];
  const int i = p + 1 + blockIdx.x;
  double acc[4][4][2];
  gpcg::zero_acc(acc);
  double* Aip = A + (long)i * 128 * ld + (long)p * 128;
  const double* Xpp = X + (long)p * 128 * ld + (long)p * 128;
  gpcg::mainloop<false>(Aip, ld, Xpp, ld, 0, 128, acc, sm);
  gpcg::store_tile(Aip, ld, acc, 1.0, 0.0);
}

// Trailing update: A_ij -= L_ip L_jp^T for p < j <= i.  grid (m, m), m = nb - p - 1.
__global__ void __launch_bounds__(gpcg::NTHREADS, 1) k_syrk_panel(double* __restrict__ A, long ld, int p) {
  extern __shared__ double sm[];
  const int j = p + 1 + blockIdx.x, i = p + 1 + blockIdx.y;
  if (j > i) return;
  double acc[4][4][2];
  gpcg::zero_acc(acc);
  gpcg::mainloop<false>(A + (long)i * 128 * ld + (long)p * 128, ld, A + (long)j * 128 * ld + (long)p * 128, ld, 0,
                        128, acc, sm);
  gpcg::store_tile(A + (long)i * 128 * ld + (long)j * 128, ld, acc, -1.0, 1.0);
}

// ------------------------------------------------------------------------------------------
// Triangular inverse by recursive doubling.  At level `sb` (half-size in tiles) node q covers
// tiles [a0, a0 + 2 sb) with split mid = a0 + sb:
//   phase 0:  T[B, A] = L[B, A] * X[A, A]          (k over A-range, >= column tile)
//   phase 1:  X[B, A] = - X[B, B] * T[B, A]        (k over B-range, <= row tile)
// grid (sb, sb, nodes); tiles past the matrix end exit.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(gpcg::NTHREADS, 1) k_linv_level(const double* __restrict__ L,
                                                                  double* __restrict__ X,
                                                                  double* __restrict__ T, long ld, int nb, int sb,
                                                                  int phase) {
  extern __shared__ double sm[];
  const int a0 = blockIdx.z * 2 * sb, mid = a0 + sb;
  const int aj = a0 + blockIdx.x, bi = mid + blockIdx.y;
  if (bi >= nb) return;
  double acc[4][4][2];
  gpcg::zero_acc(acc);
  if (phase == 0) {
    gpcg::mainloop<true>(L + (long)bi * 128 * ld, ld, X + (long)aj * 128, ld, aj * 128, mid * 128, acc, sm);
    gpcg::store_tile(T + (long)bi * 128 * ld + (long)aj * 128, ld, acc, 1.0, 0.0);
  } else {
    gpcg::mainloop<true>(X + (long)bi * 128 * ld, ld, T + (long)aj * 128, ld, mid * 128, (bi + 1) * 128, acc, sm);
    gpcg::store_tile(X + (long)bi * 128 * ld + (long)aj * 128, ld, acc, -1.0, 0.0);
  }
}

// ------------------------------------------------------------------------------------------
// Lower-triangular mat-vec helpers (bandwidth bound, used once per factorisation):
//   mode 0:  out = bias_scale * bias + sign * M x        (M lower, row access; warp per row)
//   mode 1:  partial[rb][j] = sum_{i in row block rb, i >= j} M(i, j) x(i)   (M^T x, two-stage)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_trmv_n(const double* __restrict__ M, long ld, long n,
                                                const double* __restrict__ x, const double* __restrict__ bias,
                                                double bias_scale, double sign, double* __restrict__ out) {
  const long i = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  const double* row = M + i * ld;
  double s = 0.0;
  for (long j = lane; j <= i; j += 32) s = fma(row[j], x[j], s);
  s = warp_sum(s);
  if (lane == 0) out[i] = (bias ? bias_scale * bias[i] : 0.0) + sign * s;
}

__global__ void __launch_bounds__(128) k_trmv_t_partial(const double* __restrict__ M, long ld,
                                                        const double* __restrict__ x,
                                                        double* __restrict__ partial, long n_pad) {
  const int cb = blockIdx.x, rb = blockIdx.y;
  const long j = (long)cb * 128 + threadIdx.x;
  double s = 0.0;
  if (rb >= cb) {
    const long i0 = (long)rb * 128;
#pragma unroll 8
    for (int r = 0; r < 128; ++r) {
      const long i = i0 + r;
      const double m = M[i * ld + j];
      s = fma((i >= j) ? m : 0.0, x[i], s);
    }
  }
  partial[(long)rb * n_pad + j] = s;
}

__global__ void __launch_bounds__(256) k_colsum_partial(const double* __restrict__ partial, int nrb, long n_pad,
                                                        const double* __restrict__ bias, double bias_scale,
                                                        double sign, double* __restrict__ out) {
  const long j = (long)blockIdx.x * 256 + threadIdx.x;
  if (j >= n_pad) return;
  double s = 0.0;
  for (int rb = 0; rb < nrb; ++rb) s += partial[(long)rb * n_pad + j];
  out[j] = (bias ? bias_scale * bias[j] : 0.0) + sign * s;
}

__global__ void __launch_bounds__(256) k_axpy(double* __restrict__ y, const double* __restrict__ x, long n) {
  const long i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) y[i] += x[i];
}

// scal[0] = logdet = 2 sum log L_ii ; scal[1] = y' alpha.  One CTA.
// replaces NIGP.py:159-161 (warp-shuffle reductions of the diagonal and the data-fit dot).
__global__ void __launch_bounds__(1024) k_logdet_fit(const double* __restrict__ L, long ld, long N,
                                                     const double* __restrict__ y,
                                                     const double* __restrict__ alpha, double* __restrict__ scal) {
  __shared__ double s0[32], s1[32];
  double a = 0.0, b = 0.0;
  for (long i = threadIdx.x; i < N; i += 1024) {
    a += log(L[i * ld + i]);
    if (y) b = fma(y[i], alpha[i], b);
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = a; s1[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x < 32) {
    a = warp_sum(s0[threadIdx.x]);
    b = warp_sum(s1[threadIdx.x]);
    if (threadIdx.x == 0) { scal[0] = 2.0 * a; scal[1] = b; }
  }
}

// ------------------------------------------------------------------------------------------
// Evaluator helpers (GPTrainers.py:121-137: inv(SIG), Frobenius norm, e^T inv(SIG) e).
// With X = L^-1 of SIG:  inv(SIG) = X^T X,  ||inv(SIG)||_F = ||X X^T||_F,  e^T inv(SIG) e = |X e|^2.
// k_aat_fro: partial[at * mt + bt] = sum of squares of tile (at, bt) of X X^T (bt <= at), X lower.
// k_pad_identity: rows / columns M .. m_pad-1 of a padded square matrix become an identity block.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(gpcg::NTHREADS, 1) k_aat_fro(const double* __restrict__ X, long ld, int mt,
                                                               double* __restrict__ partial) {
  extern __shared__ double sm[];
  const int bt = blockIdx.x, at = blockIdx.y;
  if (bt > at) {
    if (threadIdx.x == 0) partial[at * mt + bt] = 0.0;
    return;
  }
  double acc[4][4][2];
  gpcg::zero_acc(acc);
  gpcg::mainloop<false>(X + (long)at * 128 * ld, ld, X + (long)bt * 128 * ld, ld, 0, (bt + 1) * 128, acc, sm);
  double s = 0.0;
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int g = 0; g < 4; ++g) s = fma(acc[f][g][0], acc[f][g][0], fma(acc[f][g][1], acc[f][g][1], s));
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < gpcg::NTHREADS / 32; ++w) t += sm[w];
    partial[at * mt + bt] = t;
  }
}

__global__ void __launch_bounds__(256) k_pad_identity(double* __restrict__ A, long M, long m_pad) {
  const long j = (long)blockIdx.x * 256 + threadIdx.x;
  const long i = blockIdx.y;
  if (j >= m_pad) return;
  if (i >= M || j >= M) A[i * m_pad + j] = (i == j) ? 1.0 : 0.0;
}

// out[0] = sum_i z_i^2
__global__ void __launch_bounds__(1024) k_sumsq_vec(const double* __restrict__ z, long n, double* __restrict__ out) {
  __shared__ double s0[32];
  double a = 0.0;
  for (long i = threadIdx.x; i < n; i += 1024) a = fma(z[i], z[i], a);
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) s0[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x < 32) {
    a = warp_sum(s0[threadIdx.x]);
    if (threadIdx.x == 0) out[0] = a;
  }
}
