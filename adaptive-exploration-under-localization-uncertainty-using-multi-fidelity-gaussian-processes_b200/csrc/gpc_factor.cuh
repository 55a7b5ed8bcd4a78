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
// Diagonal block p: L_pp = chol(A_pp) and X_pp = L_pp^-1, one CTA of 512 threads working out of
// shared memory.  It sits on the critical path of the factorisation (nb serial launches), so it
// is organised for latency: 8 steps over 16-column sub-blocks, two barriers each,
//   (1) warp 0 factors AND inverts the 16 x 16 diagonal sub-block in one register-resident pass (one row of L and
//       one column of L^-1 per lane, pivots and multipliers exchanged by warp shuffles, rsqrt instead of sqrt +
//       divide) -- for step kb + 1 this runs UNDER phase (3) of step kb;
//   (2) panel products on the FP64 tensor cores (DMMA 8x8x4, 8 x 8 output tiles, K = 16):
//       L[r0:, c0:c0+16] = A[r0:, c0:c0+16] T^T  (one warp per 8-row tile) and block row kb of the inverse,
//       X[c0:c0+16, :c0] = T Y[c0:c0+16, :c0]  (one warp per 8-column tile);
//   (3) rank-16 updates as DMMA tiles: the trailing sub-matrix A[r0:, r0:] -= P P^T (lower tiles) and the
//       running right-hand side of L X = I,  Y[r0:, :r0] -= P X[c0:c0+16, :r0].
// X^T (and Y before it) lives in the unused upper triangle of the same shared array, shifted by one column;
// the leading dimension 132 makes every DMMA fragment load (one double per lane) bank-conflict free.
// status receives p + 1 for the first non-positive pivot.
// replaces: scipy cho_factor (NIGP.py:43,154,288) / LAPACK dpotrf inside GPy pdinv.
// ------------------------------------------------------------------------------------------
#define GPC_PD_LD 132
#ifndef GPC_PD_SKIP
#define GPC_PD_SKIP 0   // profiling only (profiles/microbench/potrf_diag_phases.cu): 1 tiles off, 2 serial 16 x 16 off, 4 global I/O off
#endif
constexpr int GPC_POTRF_SMEM = (128 * GPC_PD_LD + 32) * 8;

constexpr int GPC_PD_NT = 512;   // threads of the diagonal-block CTA
__global__ void __launch_bounds__(GPC_PD_NT, 1) k_potrf_diag(double* __restrict__ A, double* __restrict__ X, long ld,
                                                       int p, int* __restrict__ status) {
  extern __shared__ double sm[];
  double* S = sm;  // L in the lower triangle (incl. diagonal), X^T above it
#define XT(i, j) S[(j) * GPC_PD_LD + (i) + 1]  // X(i, j), i >= j
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  double* Ap = A + (long)p * 128 * ld + (long)p * 128;
  double* Xp = X + (long)p * 128 * ld + (long)p * 128;
  __shared__ int bad;
  if (tid == 0) bad = 0;
#pragma unroll 8
  for (int e = tid; e < 128 * 128; e += GPC_PD_NT) {
    const int r = e >> 7, c = e & 127;
    const double v = (c <= r) ? ((GPC_PD_SKIP & 4) ? (r == c ? 1.0 : 0.0) : Ap[(long)r * ld + c]) : 0.0;
    if (c <= r) S[r * GPC_PD_LD + c] = v;
    else S[r * GPC_PD_LD + c + 1] = 0.0;  // Y = 0 (slot of X(c, r))
  }
  __syncthreads();

  auto factor16 = [&](int c0) {
    if (GPC_PD_SKIP & 2) return;
    // 16 x 16 diagonal sub-block L_kk and T = L_kk^-1 in ONE pass, entirely in registers: lane l owns row l of L and
    // column l of T (lanes 16..31 mirror lanes 0..15).  Per column k: the pivot travels by one shuffle, every lane
    // takes its reciprocal square root itself (no second broadcast), and the multipliers L[j][k], j > k -- one shuffle
    // each -- feed both the rank-1 update of the trailing rows and the right-looking substitution for T.  The chain
    // shuffle -> rsqrt -> scale -> shuffle -> fma per column is the critical path of the whole diagonal block.
    const int l = lane & 15;
    double a[16], x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = (j <= l) ? S[(c0 + l) * GPC_PD_LD + c0 + j] : 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = (i == l) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      double d = __shfl_sync(0xffffffffu, a[k], k);
      if (!(d > 0.0)) {
        if (lane == 0) bad = 1;
        d = 1.0;
      }
      const double inv = rsqrt(d);
      a[k] *= inv;                       // L[l][k]; on the diagonal d / sqrt(d) = sqrt(d)
      x[k] *= inv;                       // T[k][l]
#pragma unroll
      for (int j = k + 1; j < 16; ++j) {
        const double cj = __shfl_sync(0xffffffffu, a[k], j);   // L[j][k]
        a[j] = fma(-a[k], cj, a[j]);     // entries right of the diagonal (j > l) are never read back
        x[j] = fma(-cj, x[k], x[j]);
      }
    }
    if (lane < 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j <= l) S[(c0 + l) * GPC_PD_LD + c0 + j] = a[j];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i >= l) XT(c0 + i, c0 + l) = x[i];
    }
  };
  // (3a) trailing tile (I, J), J <= I, of A[r0:, r0:] -= P P^T with P = L[r0:, c0:c0+16]; only j <= i is stored (the
  // slots right of the diagonal belong to X^T).  (Groups of four tiles per warp with shared A fragments were tried:
  // no faster -- the phase runs at about half the DMMA rate of one SM and is balance-, not latency-bound;
  // profiles/r02/README.md.)
  auto chol_tile = [&](int r0, int c0, int I, int J) {
    const int ri = r0 + 8 * I + lr, rj = r0 + 8 * J + lr, cj = r0 + 8 * J + 2 * lc;
    double acc0 = S[ri * GPC_PD_LD + cj], acc1 = S[ri * GPC_PD_LD + cj + 1];
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4)
      dmma884(acc0, acc1, -S[ri * GPC_PD_LD + c0 + 4 * s4 + lc], S[rj * GPC_PD_LD + c0 + 4 * s4 + lc]);
    if (cj <= ri) S[ri * GPC_PD_LD + cj] = acc0;
    if (cj + 1 <= ri) S[ri * GPC_PD_LD + cj + 1] = acc1;
  };
  // (3b) tile (I, J) of Y[r0:, :r0] -= P X[c0:c0+16, :r0]  (X block row kb is final: zero right of its diagonal)
  auto inv_tile = [&](int r0, int c0, int I, int J) {
    const int r = r0 + 8 * I + lr, cC = 8 * J + 2 * lc, cB = 8 * J + lr;
    double acc0 = XT(r, cC), acc1 = XT(r, cC + 1);
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
      const int m = c0 + 4 * s4 + lc;
      const double b = (cB <= m) ? XT(m, cB) : 0.0;
      dmma884(acc0, acc1, -S[r * GPC_PD_LD + m], b);
    }
    XT(r, cC) = acc0;
    XT(r, cC + 1) = acc1;
  };

  if (warp == 0) factor16(0);
  __syncthreads();
#pragma unroll 1
  for (int kb = 0; kb < 8; ++kb) {
    const int c0 = kb * 16;
    const int r0 = c0 + 16, nrow = 128 - r0;
    const int mt = nrow >> 3;          // 8-row tiles below the sub-block
    // ---- (2) panel products with T = L_kk^-1 (X^T slots of the sub-block) ----------------------------------
    const int n2a = mt, n2b = c0 >> 3;
    for (int t = warp; t < ((GPC_PD_SKIP & 1) ? 0 : n2a + n2b); t += GPC_PD_NT / 32) {
      if (t < n2a) {
        // L[r][c0 + c] = sum_{m <= c} A[r][c0 + m] T[c][m]: the warp owns rows r0 + 8 t .. + 7 and both 8-column tiles
        const int r = r0 + 8 * t + lr;
        double a[4], o[2][2];
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) a[s4] = S[r * GPC_PD_LD + c0 + 4 * s4 + lc];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          o[nt][0] = o[nt][1] = 0.0;
          const int cc = 8 * nt + lr;
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            const int m = 4 * s4 + lc;
            const double b = (m <= cc) ? XT(c0 + cc, c0 + m) : 0.0;
            dmma884(o[nt][0], o[nt][1], a[s4], b);
          }
        }
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          S[r * GPC_PD_LD + c0 + 8 * nt + 2 * lc] = o[nt][0];
          S[r * GPC_PD_LD + c0 + 8 * nt + 2 * lc + 1] = o[nt][1];
        }
      } else {
        // X[c0 + i][j] = sum_{t <= i} T[i][t] Y[c0 + t][j], j < c0: the warp owns columns 8 J .. + 7 and both 8-row tiles
        const int J = t - n2a;
        double b[4], o[2][2];
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) b[s4] = XT(c0 + 4 * s4 + lc, 8 * J + lr);
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          o[it][0] = o[it][1] = 0.0;
          const int i = 8 * it + lr;
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            const int tt = 4 * s4 + lc;
            const double a = (tt <= i) ? XT(c0 + i, c0 + tt) : 0.0;
            dmma884(o[it][0], o[it][1], a, b[s4]);
          }
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          XT(c0 + 8 * it + lr, 8 * J + 2 * lc) = o[it][0];
          XT(c0 + 8 * it + lr, 8 * J + 2 * lc + 1) = o[it][1];
        }
      }
    }
    __syncthreads();
    // ---- (3) rank-16 updates.  Warp 0 takes the three tiles of the next diagonal sub-block and goes on to factor it
    //      (the serial part of step kb + 1); the other warps share everything else. ---------------------------
    if (mt > 0) {
      const int n3a = mt * (mt + 1) / 2, n3b = mt * (r0 >> 3), nj = r0 >> 3;
      if (warp == 0) {
        chol_tile(r0, c0, 0, 0);
        chol_tile(r0, c0, 1, 0);
        chol_tile(r0, c0, 1, 1);
        __syncwarp();
        factor16(r0);
      } else {
        for (int idx = 3 + (warp - 1); idx < ((GPC_PD_SKIP & 1) ? 0 : n3a + n3b); idx += GPC_PD_NT / 32 - 1) {
          if (idx < n3a) {
            int I = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
            while (I * (I + 1) / 2 > idx) --I;
            while ((I + 1) * (I + 2) / 2 <= idx) ++I;
            chol_tile(r0, c0, I, idx - I * (I + 1) / 2);
          } else {
            const int e = idx - n3a;
            inv_tile(r0, c0, e / nj, e % nj);
          }
        }
      }
    }
    __syncthreads();
  }
  if (bad && tid == 0) atomicCAS(status, 0, p + 1);
  if (GPC_PD_SKIP & 4) {
    if (tid == 0) Ap[0] = S[0] + XT(127, 0);
    return;
  }
#pragma unroll 8
  for (int e = tid; e < 128 * 128; e += GPC_PD_NT) {
    const int r = e >> 7, c = e & 127;
    Ap[(long)r * ld + c] = (c <= r) ? S[r * GPC_PD_LD + c] : 0.0;
    Xp[(long)r * ld + c] = (c <= r) ? XT(r, c) : 0.0;
  }
#undef XT
}

// ------------------------------------------------------------------------------------------
// Panel solve: L_ip = A_ip * X_pp^T  for block rows i > p (TRSM expressed as a DMMA contraction
// with the inverted diagonal block), in place.  One CTA owns 64 rows of the panel and computes
// both 64-column halves before it stores either (its rows are read by no other CTA).
// grid = 2 (nb - p - 1) - row0/64 ...: rows [r_lo, r_hi) in units of 64.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(gpc64::NT, 2) k_trsm_panel(double* __restrict__ A, const double* __restrict__ X,
                                                             long ld, int p, int r64_lo) {
  extern __shared__ double sm[];
  const long r0 = ((long)r64_lo + blockIdx.x) * 64;
  double* Arow = A + r0 * ld + (long)p * 128;
  const double* Xpp = X + (long)p * 128 * ld + (long)p * 128;
  double acc0[4][4][2], acc1[4][4][2];
  gpcg::zero_acc(acc0);
  gpcg::zero_acc(acc1);
  // four-stage ring: the panel solve sits on the serial chain of the factorisation with far fewer CTAs than SMs
  gpc64::mainloop<false, 4>(Arow, ld, Xpp, ld, 0, 64, acc0, sm);          // X_pp lower: columns 0..63 need k < 64
  gpc64::mainloop<false, 4>(Arow, ld, Xpp + 64 * ld, ld, 0, 128, acc1, sm);
  gpc64::store_tile(Arow, ld, acc0, 1.0, 0.0);
  gpc64::store_tile(Arow + 64, ld, acc1, 1.0, 0.0);
}

// Trailing update: A_ij -= L_i,[p..] L_j,[p..]^T (kdim = 128: panel p, 256: panels p and p+1) over 64 x 64 tiles with row tile ti in [t_lo, t_hi) and column
// tile tj in [c_lo, c_hi), tj <= ti (all in units of 64 rows).  grid (c_hi - c_lo, t_hi - t_lo).
// DEEP = the four-stage ring (80 KB) for the narrow look-ahead launches on the serial chain.
template <bool DEEP>
__global__ void __launch_bounds__(gpc64::NT, DEEP ? 2 : 4) k_syrk_panel(double* __restrict__ A, long ld, int p, int t_lo,
                                                                        int c_lo, int kdim) {
  extern __shared__ double sm[];
  const int tj = c_lo + blockIdx.x, ti = t_lo + blockIdx.y;
  if (tj > ti) return;
  double acc[4][4][2];
  gpcg::zero_acc(acc);
  gpc64::mainloop<false, DEEP ? 4 : 2>(A + (long)ti * 64 * ld + (long)p * 128, ld, A + (long)tj * 64 * ld + (long)p * 128, ld,
                                       0, kdim, acc, sm);
  gpc64::store_tile(A + (long)ti * 64 * ld + (long)tj * 64, ld, acc, -1.0, 1.0);
}

// ------------------------------------------------------------------------------------------
// Triangular inverse by recursive doubling.  At level `sb` (half-size in 128-blocks) node q covers
// blocks [a0, a0 + 2 sb) with split mid = a0 + sb:
//   phase 0:  T[B, A] = L[B, A] * X[A, A]          (k over the A-range, >= the column tile)
//   phase 1:  X[B, A] = - X[B, B] * T[B, A]        (k over the B-range, <= the row tile)
// grid (2 sb, 2 sb, nodes) over 64 x 64 tiles, nodes node0 .. node0 + gridDim.z - 1; tiles past the matrix end exit.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(gpc64::NT, 4) k_linv_level(const double* __restrict__ L, double* __restrict__ X,
                                                             double* __restrict__ T, long ld, int nb, int sb,
                                                             int phase, int node0) {
  extern __shared__ double sm[];
  const int a0 = (node0 + blockIdx.z) * 2 * sb, mid = a0 + sb;  // in 128-blocks
  const int aj = 2 * a0 + blockIdx.x;                          // column tile (64s) inside the A-range
  const int bi = 2 * mid + (2 * sb - 1 - (int)blockIdx.y);     // row tile (64s) inside the B-range, long k first
  if (bi >= 2 * nb) return;
  double acc[4][4][2];
  gpcg::zero_acc(acc);
  if (phase == 0) {
    gpc64::mainloop<true>(L + (long)bi * 64 * ld, ld, X + (long)aj * 64, ld, aj * 64, mid * 128, acc, sm);
    gpc64::store_tile(T + (long)bi * 64 * ld + (long)aj * 64, ld, acc, 1.0, 0.0);
  } else {
    gpc64::mainloop<true>(X + (long)bi * 64 * ld, ld, T + (long)aj * 64, ld, mid * 128, (bi + 1) * 64, acc, sm);
    gpc64::store_tile(X + (long)bi * 64 * ld + (long)aj * 64, ld, acc, -1.0, 0.0);
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
  // four independent accumulators, eight loads in flight per lane: the loop is latency-, not bandwidth-bound
  double s = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  long j = lane;
  for (; j + 96 <= i; j += 128) {
    s = fma(row[j], x[j], s);
    s1 = fma(row[j + 32], x[j + 32], s1);
    s2 = fma(row[j + 64], x[j + 64], s2);
    s3 = fma(row[j + 96], x[j + 96], s3);
  }
  for (; j <= i; j += 32) s = fma(row[j], x[j], s);
  s = warp_sum((s + s1) + (s2 + s3));
  if (lane == 0) out[i] = (bias ? bias_scale * bias[i] : 0.0) + sign * s;
}

__global__ void __launch_bounds__(512) k_trmv_t_partial(const double* __restrict__ M, long ld,
                                                        const double* __restrict__ x,
                                                        double* __restrict__ partial, long n_pad) {
  // 512 threads: column j = threadIdx.x % 128 of the tile, rows split over four thread groups of 32 rows each
  // (a quarter of the serial chain per thread), combined through shared memory
  __shared__ double red[3][128];
  const int cb = blockIdx.x, rb = blockIdx.y;
  const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;
  const long j = (long)cb * 128 + tx;
  double s = 0.0;
  if (rb >= cb) {
    const long i0 = (long)rb * 128 + 32 * ty;
#pragma unroll 16
    for (int r = 0; r < 32; ++r) {
      const long i = i0 + r;
      const double m = M[i * ld + j];
      s = fma((i >= j) ? m : 0.0, x[i], s);
    }
  }
  if (ty > 0) red[ty - 1][tx] = s;
  __syncthreads();
  if (ty == 0) partial[(long)rb * n_pad + j] = ((s + red[0][tx]) + (red[1][tx] + red[2][tx]));
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
