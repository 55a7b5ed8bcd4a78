// gpc_ig.cuh -- information-gain scoring of candidate planner paths.
//
// Every candidate owns a "span" of rows inside one 128-row tile of the candidate-row matrices
// (a span never straddles a tile, so all of its k x k Gram blocks come out of ONE diagonal-tile
// DMMA contraction):
//      rows [rs, rs + k)        the candidate's own points (fidelity = its own label)
//      rows [rs + k, rs + 2k)   only when the query fidelity differs: the same points at pred_fid
// Heavy steps reuse the shared DMMA mainloop (gpc_gemm.cuh):
//      Vt   = Kx  X^T          k_vt            (L^-1 k for every candidate row)
//      Gram = Vt_tile Vt_tile^T   k_gram_diag  (Schur complements of all candidates of a tile)
//      Bt   = K(cand, grid) - Vt Vg^T          k_cross_cov   (log-det variant)
//      Zt   = Bt Xg^T          k_vt again, with the grid factor
// and the per-candidate k x k work (Cholesky, forward substitution, log sums) runs one CTA per
// candidate out of shared memory (k_ig_seq_cand / k_ig_logdet_cand).
//
// replaces: GraceRIGV3.py:443-562 and PhysicalExperimentCode/GraceRIGV3.py:446-678, where each
// candidate costs one (log-det) or k (sequential) full O((N+k)^3) GPy refits.
#pragma once
#include "gpc_gemm.cuh"

// Gram[tile] (128 x 128, row-major, compact) = A_tile A_tile^T, A = rows [128 tile, +128) of a
// [m_pad][ld] matrix, contraction over kdim (multiple of 16).  grid = number of tiles.
__global__ void __launch_bounds__(gpcg::NTHREADS, 1) k_gram_diag(const double* __restrict__ A, long ld, int kdim,
                                                                 double* __restrict__ Gram) {
  extern __shared__ double sm[];
  const double* At = A + (long)blockIdx.x * 128 * ld;
  double acc[4][4][2];
  gpcg::zero_acc(acc);
  gpcg::mainloop<false>(At, ld, At, ld, 0, kdim, acc, sm);
  gpcg::store_tile(Gram + (long)blockIdx.x * 128 * 128, 128, acc, 1.0, 0.0);
}

// Bt[r][j] = k(row r, grid j) - sum_n Vt[r][n] Vg[j][n]  (latent cross-covariance given the data),
// zero for invalid candidate rows (fid < 0) and for j >= G.  grid (m_pad/128, g_pad/128).
__global__ void __launch_bounds__(gpcg::NTHREADS, 1) k_cross_cov(const __grid_constant__ GpcHyp h,
                                                                 const double* __restrict__ Vt,
                                                                 const double* __restrict__ Vg, long n_pad,
                                                                 const double* __restrict__ Xr4,
                                                                 const double* __restrict__ Xg4, long G,
                                                                 long g_pad, double* __restrict__ Bt) {
  extern __shared__ double sm[];
  const int rt = blockIdx.x, gt = blockIdx.y;
  double acc[4][4][2];
  gpcg::zero_acc(acc);
  gpcg::mainloop<false>(Vt + (long)rt * 128 * n_pad, n_pad, Vg + (long)gt * 128 * n_pad, n_pad, 0, (int)n_pad, acc,
                        sm);
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const long r = (long)rt * 128 + gpcg::acc_row(f);
    const double ax = Xr4[r * 4], ay = Xr4[r * 4 + 1], az = Xr4[r * 4 + 2], af = Xr4[r * 4 + 3];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      double2 v;
      double* ve = &v.x;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const long j = (long)gt * 128 + gpcg::acc_col(g) + e;
        double o = 0.0;
        if (af >= 0.0 && j < G)
          o = gpc_kval(h, ax, ay, az, af, Xg4[j * 4], Xg4[j * 4 + 1], Xg4[j * 4 + 2], Xg4[j * 4 + 3]) -
              acc[f][g][e];
        ve[e] = o;
      }
      *reinterpret_cast<double2*>(Bt + r * g_pad + (long)gt * 128 + gpcg::acc_col(g)) = v;
    }
  }
}

__global__ void k_fill(double* __restrict__ x, int n, double v) {
  if ((int)threadIdx.x < n) x[threadIdx.x] = v;
}

// INT8 path: the contraction P = V_c V_g^T comes from the tcgen05 kernel (digit images of V); this finishes
// Bt[r][j] = k(row r, grid j) - P[r][j] in place (zero for invalid candidate rows and j >= G).
__global__ void __launch_bounds__(256) k_cross_fin(const __grid_constant__ GpcHyp h, const double* __restrict__ Xr4,
                                                   const double* __restrict__ Xg4, long G, long g_pad, long m_pad,
                                                   double* __restrict__ Bt) {
  const long total = m_pad * g_pad;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long r = idx / g_pad, j = idx - r * g_pad;
    const double af = Xr4[r * 4 + 3];
    double o = 0.0;
    if (af >= 0.0 && j < G)
      o = gpc_kval(h, Xr4[r * 4], Xr4[r * 4 + 1], Xr4[r * 4 + 2], af, Xg4[j * 4], Xg4[j * 4 + 1], Xg4[j * 4 + 2],
                   Xg4[j * 4 + 3]) - Bt[idx];
    Bt[idx] = o;
  }
}

// ---- per-candidate small-matrix helpers (shared memory, leading dimension 65) ----------------
#define GPC_IG_LD 65
// dynamic shared memory of the per-candidate kernels: two k x k matrices, the span's points, scratch
constexpr int GPC_IG_SMEM = (2 * GPC_MAXK * GPC_IG_LD + 2 * GPC_MAXK * 4 + 6 + GPC_MAXK / 8) * 8;

// In-place lower Cholesky of the k x k matrix S (only the lower triangle is read / written).
// Returns false (to every thread) when a pivot is not positive.  blockDim = 128.
__device__ __forceinline__ bool ig_chol(double* S, int k, int* flag) {
  const int tid = threadIdx.x;
  if (tid == 0) *flag = 0;
  __syncthreads();
  for (int j = 0; j < k; ++j) {
    if (tid == 0) {
      const double d = S[j * GPC_IG_LD + j];
      if (!(d > 0.0)) *flag = 1;
      S[j * GPC_IG_LD + j] = sqrt(d);
    }
    __syncthreads();
    const double inv = 1.0 / S[j * GPC_IG_LD + j];
    __syncthreads();
    for (int i = j + 1 + tid; i < k; i += blockDim.x) S[i * GPC_IG_LD + j] *= inv;
    __syncthreads();
    // trailing update of the lower triangle, one thread per (i, l) with j < l <= i < k
    const int m = k - j - 1;
    for (int e = tid; e < m * m; e += blockDim.x) {
      const int i = j + 1 + e / m, l = j + 1 + e % m;
      if (l <= i) S[i * GPC_IG_LD + l] = fma(-S[i * GPC_IG_LD + j], S[l * GPC_IG_LD + j], S[i * GPC_IG_LD + l]);
    }
    __syncthreads();
  }
  return *flag == 0;
}

__device__ __forceinline__ double ig_block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}

// ------------------------------------------------------------------------------------------
// Sequential information gain of one candidate (one CTA, 128 threads):
//   I = sum_i log(1 + (max(lat_i, 1e-15) + noise_q) / sig_n)
//   lat_i = q_i - sum_{j in cond, j < i or (j == i and pre_i)} W[j][i]^2,  W = chol(S)^-1 cross
// with S the noise-inclusive Schur complement of the conditioned candidate rows (targets are
// zero, so only covariances matter), cross[j][i] = cov(f(c_j), f(p_i) | data), q_i = var(f(p_i) | data).
// rowmask bit0 = the row is conditioned on by later rows, bit1 = the row is appended before it
// is predicted itself (GraceRIGV3.py:454-455; windowed variants :486-501, :547-556).
//   span[c] = {row start, k, has separate query rows}
// replaces: calcPathInfoSF2/SF3 (k refits), calculatePathInfoEmu (k refits, query at fidelity 0).
// ------------------------------------------------------------------------------------------
struct GpcSpan {
  int rs;    // first row of the span in the padded candidate-row matrices
  int k;     // number of points
  int pred;  // 1: rows [rs + k, rs + 2k) hold the query copies
  int pad;
};

__global__ void __launch_bounds__(128) k_ig_seq_cand(const __grid_constant__ GpcHyp h,
                                                     const GpcSpan* __restrict__ spans,
                                                     const double* __restrict__ Xr4,
                                                     const double* __restrict__ Gram,
                                                     const unsigned char* __restrict__ rowmask, double sig_n,
                                                     double* __restrict__ I_out) {
  extern __shared__ double dsm[];
  double* S = dsm;
  double* Cr = S + GPC_MAXK * GPC_IG_LD;
  double(*pt)[4] = reinterpret_cast<double(*)[4]>(Cr + GPC_MAXK * GPC_IG_LD);
  double* red = Cr + GPC_MAXK * GPC_IG_LD + 2 * GPC_MAXK * 4;
  int* flagp = reinterpret_cast<int*>(red + 4);
  unsigned char* msk = reinterpret_cast<unsigned char*>(red + 6);
  const GpcSpan sp = spans[blockIdx.x];
  const int k = sp.k, tid = threadIdx.x;
  if (k == 0) {
    if (tid == 0) I_out[blockIdx.x] = 0.0;
    return;
  }
  const int tile = sp.rs >> 7, r0 = sp.rs & 127;
  const double* Gt = Gram + (long)tile * 128 * 128;
  const int rows = sp.pred ? 2 * k : k;
  for (int e = tid; e < rows * 4; e += 128) pt[e >> 2][e & 3] = Xr4[(long)(sp.rs + (e >> 2)) * 4 + (e & 3)];
  for (int e = tid; e < k; e += 128) msk[e] = rowmask ? rowmask[sp.rs + e] : (unsigned char)1;
  __syncthreads();
  const int qo = sp.pred ? k : 0;  // offset of the query rows
  for (int e = tid; e < k * k; e += 128) {
    const int j = e / k, i = e % k;  // j: conditioned (own) row, i: query column
    const bool cj = msk[j] & 1;
    // cross[j][i]
    double c = 0.0;
    if (cj)
      c = gpc_kval(h, pt[j][0], pt[j][1], pt[j][2], pt[j][3], pt[qo + i][0], pt[qo + i][1], pt[qo + i][2],
                   pt[qo + i][3]) -
          Gt[(r0 + j) * 128 + r0 + qo + i];
    Cr[j * GPC_IG_LD + i] = c;
    if (i <= j) {
      double s;
      if (cj && (msk[i] & 1)) {
        s = gpc_kval(h, pt[j][0], pt[j][1], pt[j][2], pt[j][3], pt[i][0], pt[i][1], pt[i][2], pt[i][3]) -
            Gt[(r0 + j) * 128 + r0 + i];
        if (i == j) s += h.noise[gpc_fid(h, pt[j][3])] + h.jitter;
      } else {
        s = (i == j) ? 1.0 : 0.0;
      }
      S[j * GPC_IG_LD + i] = s;
    }
  }
  __syncthreads();
  const bool ok = ig_chol(S, k, flagp);
  double term = 0.0;
  if (tid < k) {
    const int i = tid;
    const int fq = gpc_fid(h, pt[qo + i][3]);
    double lat = h.kdiag[fq] - Gt[(r0 + qo + i) * 128 + r0 + qo + i];
    const int upto = (msk[i] & 2) ? i + 1 : i;
    // forward substitution of column i (in place in Cr; other threads own other columns)
    for (int j = 0; j < upto; ++j) {
      double w = Cr[j * GPC_IG_LD + i];
      for (int l = 0; l < j; ++l) w = fma(-S[j * GPC_IG_LD + l], Cr[l * GPC_IG_LD + i], w);
      w /= S[j * GPC_IG_LD + j];
      Cr[j * GPC_IG_LD + i] = w;
      lat = fma(-w, w, lat);
    }
    lat = fmax(lat, 1e-15);  // GPy clips the latent marginal variance
    term = log(1.0 + (lat + h.noise[fq]) / sig_n);
  }
  const double I = ig_block_sum(term, red);
  if (tid == 0) I_out[blockIdx.x] = ok ? I : __longlong_as_double(0x7ff8000000000000LL);
}

// ------------------------------------------------------------------------------------------
// Log-det information gain of one candidate on the fixed grid:
//   I = 0.5 (logdet S - logdet T),  S = Schur complement of the candidate (noise + jitter on the
//   diagonal), T = S - Z^T Z with Z = Lg^-1 B  (matrix-determinant lemma: equals
//   0.5 (logdet Sigma_prior(grid) - logdet Sigma_post(grid | data u candidate))).
// replaces: calcPathInfoSFBatch / calculatePathInfoEmuBatch (one refit + G x G determinant each).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_ig_logdet_cand(const __grid_constant__ GpcHyp h,
                                                        const GpcSpan* __restrict__ spans,
                                                        const double* __restrict__ Xr4,
                                                        const double* __restrict__ Gram,
                                                        const double* __restrict__ GramZ,
                                                        double* __restrict__ I_out) {
  extern __shared__ double dsm[];
  double* S = dsm;
  double* T = S + GPC_MAXK * GPC_IG_LD;
  double(*pt)[4] = reinterpret_cast<double(*)[4]>(T + GPC_MAXK * GPC_IG_LD);
  double* red = T + GPC_MAXK * GPC_IG_LD + 2 * GPC_MAXK * 4;
  int* flagp = reinterpret_cast<int*>(red + 4);
  const GpcSpan sp = spans[blockIdx.x];
  const int k = sp.k, tid = threadIdx.x;
  if (k == 0) {
    if (tid == 0) I_out[blockIdx.x] = 0.0;
    return;
  }
  const int tile = sp.rs >> 7, r0 = sp.rs & 127;
  const double* Gt = Gram + (long)tile * 128 * 128;
  const double* Gz = GramZ + (long)tile * 128 * 128;
  for (int e = tid; e < k * 4; e += 128) pt[e >> 2][e & 3] = Xr4[(long)(sp.rs + (e >> 2)) * 4 + (e & 3)];
  __syncthreads();
  for (int e = tid; e < k * k; e += 128) {
    const int j = e / k, i = e % k;
    if (i > j) continue;
    double s = gpc_kval(h, pt[j][0], pt[j][1], pt[j][2], pt[j][3], pt[i][0], pt[i][1], pt[i][2], pt[i][3]) -
               Gt[(r0 + j) * 128 + r0 + i];
    if (i == j) s += h.noise[gpc_fid(h, pt[j][3])] + h.jitter;
    S[j * GPC_IG_LD + i] = s;
    T[j * GPC_IG_LD + i] = s - Gz[(r0 + j) * 128 + r0 + i];
  }
  __syncthreads();
  const bool ok1 = ig_chol(S, k, flagp);
  __syncthreads();
  const bool ok2 = ig_chol(T, k, flagp);
  double term = 0.0;
  if (tid < k) term = log(S[tid * GPC_IG_LD + tid]) - log(T[tid * GPC_IG_LD + tid]);
  const double I = ig_block_sum(term, red);
  if (tid == 0) I_out[blockIdx.x] = (ok1 && ok2) ? I : __longlong_as_double(0x7ff8000000000000LL);
}

// ------------------------------------------------------------------------------------------
// "Self-grid" log-det information gain of one candidate (calculatePathInfoEmu2,
// GraceRIGV3.py:505-523): the evaluation grid is the candidate itself queried at pred_fid,
//   I = 0.5 (logdet K(Xp) - logdet Sigma_post(Xp | data u candidate)),
// K(Xp) the PRIOR kernel matrix (no noise, no data), Sigma_post the noise-inclusive predictive
// covariance, optionally clipped element-wise at 1e-10 like emukit's predict_covariance.
//   Sigma_post = Cpp - W^T W + noise_p I,  W = chol(S)^-1 Ccp,  S = Ccc + noise_c + jitter,
// C.. = latent covariances given the data (kernel value minus the Gram block of L^-1 k).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_ig_selfgrid_cand(const __grid_constant__ GpcHyp h,
                                                          const GpcSpan* __restrict__ spans,
                                                          const double* __restrict__ Xr4,
                                                          const double* __restrict__ Gram, int clip,
                                                          double* __restrict__ I_out) {
  extern __shared__ double dsm[];
  double* S = dsm;
  double* Cr = S + GPC_MAXK * GPC_IG_LD;
  double(*pt)[4] = reinterpret_cast<double(*)[4]>(Cr + GPC_MAXK * GPC_IG_LD);
  double* red = Cr + GPC_MAXK * GPC_IG_LD + 2 * GPC_MAXK * 4;
  int* flagp = reinterpret_cast<int*>(red + 4);
  const GpcSpan sp = spans[blockIdx.x];
  const int k = sp.k, tid = threadIdx.x;
  if (k == 0) {
    if (tid == 0) I_out[blockIdx.x] = 0.0;
    return;
  }
  const int tile = sp.rs >> 7, r0 = sp.rs & 127;
  const double* Gt = Gram + (long)tile * 128 * 128;
  const int rows = sp.pred ? 2 * k : k, qo = sp.pred ? k : 0;
  for (int e = tid; e < rows * 4; e += 128) pt[e >> 2][e & 3] = Xr4[(long)(sp.rs + (e >> 2)) * 4 + (e & 3)];
  __syncthreads();
  for (int e = tid; e < k * k; e += 128) {
    const int j = e / k, i = e % k;
    Cr[j * GPC_IG_LD + i] = gpc_kval(h, pt[j][0], pt[j][1], pt[j][2], pt[j][3], pt[qo + i][0], pt[qo + i][1],
                                     pt[qo + i][2], pt[qo + i][3]) -
                            Gt[(r0 + j) * 128 + r0 + qo + i];
    if (i <= j) {
      double sv = gpc_kval(h, pt[j][0], pt[j][1], pt[j][2], pt[j][3], pt[i][0], pt[i][1], pt[i][2], pt[i][3]) -
                  Gt[(r0 + j) * 128 + r0 + i];
      if (i == j) sv += h.noise[gpc_fid(h, pt[j][3])] + h.jitter;
      S[j * GPC_IG_LD + i] = sv;
    }
  }
  __syncthreads();
  const bool ok0 = ig_chol(S, k, flagp);
  if (tid < k) {  // W = Ls^-1 Ccp, column tid
    const int i = tid;
    for (int j = 0; j < k; ++j) {
      double w = Cr[j * GPC_IG_LD + i];
      for (int l = 0; l < j; ++l) w = fma(-S[j * GPC_IG_LD + l], Cr[l * GPC_IG_LD + i], w);
      Cr[j * GPC_IG_LD + i] = w / S[j * GPC_IG_LD + j];
    }
  }
  __syncthreads();
  for (int e = tid; e < k * k; e += 128) {  // Sigma_post (lower) into S
    const int a = e / k, b = e % k;
    if (b > a) continue;
    double v = gpc_kval(h, pt[qo + a][0], pt[qo + a][1], pt[qo + a][2], pt[qo + a][3], pt[qo + b][0], pt[qo + b][1],
                        pt[qo + b][2], pt[qo + b][3]) -
               Gt[(r0 + qo + a) * 128 + r0 + qo + b];
    for (int j = 0; j < k; ++j) v = fma(-Cr[j * GPC_IG_LD + a], Cr[j * GPC_IG_LD + b], v);
    if (a == b) v += h.noise[gpc_fid(h, pt[qo + a][3])];
    if (clip) v = fmax(v, 1e-10);
    S[a * GPC_IG_LD + b] = v;
  }
  __syncthreads();
  const bool ok1 = ig_chol(S, k, flagp);
  double term = (tid < k) ? -log(S[tid * GPC_IG_LD + tid]) : 0.0;
  __syncthreads();
  for (int e = tid; e < k * k; e += 128) {  // prior kernel matrix of the query rows
    const int a = e / k, b = e % k;
    if (b <= a)
      S[a * GPC_IG_LD + b] = gpc_kval(h, pt[qo + a][0], pt[qo + a][1], pt[qo + a][2], pt[qo + a][3], pt[qo + b][0],
                                      pt[qo + b][1], pt[qo + b][2], pt[qo + b][3]);
  }
  __syncthreads();
  const bool ok2 = ig_chol(S, k, flagp);
  if (tid < k) term += log(S[tid * GPC_IG_LD + tid]);
  const double I = ig_block_sum(term, red);
  if (tid == 0) I_out[blockIdx.x] = (ok0 && ok1 && ok2) ? I : __longlong_as_double(0x7ff8000000000000LL);
}

// ------------------------------------------------------------------------------------------
// Log-det information gain with emukit's ELEMENT-WISE clip reproduced (calculatePathInfoEmuBatch,
// PhysicalExperimentCode/GraceRIGV3.py:599-618: both determinants are taken of
// GPyMultiOutputWrapper.predict_covariance(grid), which returns np.clip(cov, 1e-10, inf) -- every
// negative posterior covariance between two grid points becomes 1e-10).  The clip is not a low-rank
// change, so the determinant lemma of k_ig_logdet_cand does not apply: per candidate the G x G matrix
//   A = clip(S0 - W W^T),   S0 = noise-inclusive grid covariance given the data,
//   W = B L_c^-T,  B = cov(grid, candidate | data) (G x k),  L_c L_c^T = candidate covariance + noise
// is formed and factored.  One CTA (256 threads) per candidate, persistent over the chunk:
//   * W^T (k x GP) lives in shared memory;
//   * a left-looking blocked Cholesky with 32-column panels: thread = matrix row, 32 accumulators in
//     registers, the panel entries are generated on the fly (S0 row segment - W W^T, clipped), updated
//     with the previous panels (L^T kept column-major in an L2-resident scratch so that the row-per-
//     thread reads coalesce; the 32 x 32 block of multipliers is staged in shared memory), the
//     diagonal block is factored by one warp, the rows below are solved in registers.
// logdet = 2 sum log L_ii; NaN when the clipped matrix is not positive definite (the reference's
// log(det) is nan or meaningless there).  Rows / columns G..GP-1 are an identity pad.
//   out[c] = ldp ? 0.5 (*ldp - logdet) : logdet      (spans == NULL: one evaluation with k = 0 = the prior)
// ------------------------------------------------------------------------------------------
constexpr int GPC_CLIP_NB = 32;
// shared memory (doubles): [S (kcap x 65) -- reused for D and Lp (2 x 32 x 33) once W^T is formed][points kcap x 4]
// [scratch 18 + 32 reciprocal pivots][W^T kcap x GP].  Two CTAs fit per SM for k <= 32, G <= 320: while one CTA is in
// the one-warp diagonal-block factorisation the other keeps the FP64 pipe busy.
inline int ig_clip_head(int kcap) {
  const int a = kcap * GPC_IG_LD, b = 2 * GPC_CLIP_NB * (GPC_CLIP_NB + 1);
  return (a > b ? a : b) + kcap * 4 + 18 + GPC_CLIP_NB;
}
inline size_t ig_clip_smem(int kcap, int GP) { return ((size_t)ig_clip_head(kcap) + (size_t)kcap * GP) * 8; }

__global__ void __launch_bounds__(256, 2) k_ig_logdet_clip(const __grid_constant__ GpcHyp h,
                                                           const GpcSpan* __restrict__ spans, int ncand,
                                                           const double* __restrict__ Xr4,
                                                           const double* __restrict__ Gram,
                                                           const double* __restrict__ Bt, long ldb,
                                                           const double* __restrict__ S0, long lds, int G, int GP,
                                                           double* scratch, double clip_lo,
                                                           const double* __restrict__ ldp, double* __restrict__ out,
                                                           int kcap) {
  extern __shared__ double dsm[];
  constexpr int NB = GPC_CLIP_NB, LDD = GPC_CLIP_NB + 1;
  const int sdim = (kcap * GPC_IG_LD > 2 * NB * LDD) ? kcap * GPC_IG_LD : 2 * NB * LDD;
  double* S = dsm;                 // candidate covariance, dead once W^T exists ...
  double* D = dsm;                 // ... then the diagonal block and the staged multipliers live there
  double* Lp = D + NB * LDD;       // [32][32], no padding: its reads are warp-wide broadcasts, 16-byte vectorisable
  double(*pt)[4] = reinterpret_cast<double(*)[4]>(dsm + sdim);
  double* red = dsm + sdim + kcap * 4;
  int* flagp = reinterpret_cast<int*>(red + 16);
  double* rdiag = red + 18;        // reciprocal pivots of the current diagonal block
  double* Wt = rdiag + NB;
  double* Lt = scratch + (size_t)blockIdx.x * GP * GP;   // Lt[j * GP + i] = L[i][j]
  const int tid = threadIdx.x;
  for (int c = blockIdx.x; c < ncand; c += gridDim.x) {
    __syncthreads();
    int k = 0;
    bool ok = true;
    if (spans) {
      const GpcSpan sp = spans[c];
      k = sp.k;
      if (k > 0) {
        const int tile = sp.rs >> 7, r0 = sp.rs & 127;
        const double* Gt = Gram + (long)tile * 128 * 128;
        for (int e = tid; e < k * 4; e += blockDim.x) pt[e >> 2][e & 3] = Xr4[(long)(sp.rs + (e >> 2)) * 4 + (e & 3)];
        __syncthreads();
        for (int e = tid; e < k * k; e += blockDim.x) {
          const int j = e / k, i = e % k;
          if (i > j) continue;
          double s = gpc_kval(h, pt[j][0], pt[j][1], pt[j][2], pt[j][3], pt[i][0], pt[i][1], pt[i][2], pt[i][3]) -
                     Gt[(r0 + j) * 128 + r0 + i];
          if (i == j) s += h.noise[gpc_fid(h, pt[j][3])] + h.jitter;
          S[j * GPC_IG_LD + i] = s;
        }
        __syncthreads();
        ok = ig_chol(S, k, flagp);
        // W^T = L_c^-1 B^T, one grid column per thread
        for (int i = tid; i < GP; i += blockDim.x) {
          for (int t = 0; t < k; ++t) {
            double v = (i < G) ? Bt[(long)(sp.rs + t) * ldb + i] : 0.0;
            for (int s2 = 0; s2 < t; ++s2) v = fma(-S[t * GPC_IG_LD + s2], Wt[s2 * GP + i], v);
            Wt[t * GP + i] = v / S[t * GPC_IG_LD + t];
          }
        }
        __syncthreads();
      }
    }
    double ldsum = 0.0;
    int bad = 0;
    for (int p0 = 0; p0 < GP; p0 += NB) {
      for (int base = p0; base < GP; base += blockDim.x) {
        const int i = base + tid;
        const bool act = i < GP;
        double acc[NB];
        if (act) {
          if (i < G) {
            const double* srow = S0 + (long)i * lds + p0;
#pragma unroll
            for (int cc = 0; cc < NB; ++cc) acc[cc] = (p0 + cc < G) ? srow[cc] : 0.0;
          } else {
#pragma unroll
            for (int cc = 0; cc < NB; ++cc) acc[cc] = (p0 + cc == i) ? 1.0 : 0.0;
          }
          for (int t = 0; t < k; ++t) {
            const double w = Wt[t * GP + i];
            const double* wp = Wt + t * GP + p0;
#pragma unroll
            for (int cc = 0; cc < NB; ++cc) acc[cc] = fma(-w, wp[cc], acc[cc]);
          }
          if (i < G) {
#pragma unroll
            for (int cc = 0; cc < NB; ++cc)
              if (p0 + cc < G) acc[cc] = fmax(acc[cc], clip_lo);
          }
        }
        for (int j0 = 0; j0 < p0; j0 += NB) {
          __syncthreads();
          for (int e = tid; e < NB * NB; e += blockDim.x) Lp[e] = Lt[(size_t)(j0 + (e >> 5)) * GP + p0 + (e & 31)];
          __syncthreads();
          if (act) {
            // the multipliers of this row come out of the L2-resident scratch: fetch eight at a time so that
            // their latency overlaps (one load followed by its 32 FMAs would expose it every iteration)
#pragma unroll 1
            for (int j8 = 0; j8 < NB; j8 += 8) {
              double av[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) av[u] = Lt[(size_t)(j0 + j8 + u) * GP + i];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const double2* lp2 = reinterpret_cast<const double2*>(Lp + (j8 + u) * NB);
#pragma unroll
                for (int cc = 0; cc < NB; cc += 2) {
                  const double2 l2 = lp2[cc >> 1];
                  acc[cc] = fma(-av[u], l2.x, acc[cc]);
                  acc[cc + 1] = fma(-av[u], l2.y, acc[cc + 1]);
                }
              }
            }
          }
        }
        if (base == p0) {   // the 32 rows of the diagonal block are threads 0..31 (warp 0) of this chunk
          // acc[] of lane l IS row l of the diagonal block: factor it in place, in registers.  Per column the pivot
          // travels by one shuffle, every lane takes its reciprocal square root itself, and the multipliers
          // L[l2][j] come by one shuffle each (the scheme of k_potrf_diag's 16 x 16 block; the shared-memory version
          // with sqrt, divide and a per-lane update loop cost ~25 us per block, a third of the kernel).
          __syncthreads();                                    // every reader of the previous D / rdiag is done
          if (tid < NB) {
            const int lane = tid;
            double mydiag = 1.0;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
              double d = __shfl_sync(0xffffffffu, acc[j], j);
              if (!(d > 0.0)) {
                bad = 1;
                d = 1.0;
              }
              const double inv = rsqrt(d);
              acc[j] *= inv;                                  // L[lane][j] (unused for lane < j)
              if (lane == j) mydiag = d * inv;                // sqrt(d)
#pragma unroll
              for (int l2 = j + 1; l2 < NB; ++l2) {
                const double cj = __shfl_sync(0xffffffffu, acc[j], l2);   // L[l2][j]
                acc[l2] = fma(-acc[j], cj, acc[l2]);
              }
            }
            ldsum += log(mydiag);
            rdiag[lane] = 1.0 / mydiag;
#pragma unroll
            for (int cc = 0; cc < NB; ++cc) D[lane * LDD + cc] = acc[cc];   // entries right of the diagonal are never read
          }
          __syncthreads();
        }
        if (act) {
          if (i >= p0 + NB) {
#pragma unroll
            for (int cc = 0; cc < NB; ++cc) {
              double v = acc[cc];
#pragma unroll
              for (int c2 = 0; c2 < cc; ++c2) v = fma(-acc[c2], D[cc * LDD + c2], v);
              acc[cc] = v * rdiag[cc];
            }
#pragma unroll
            for (int cc = 0; cc < NB; ++cc) Lt[(size_t)(p0 + cc) * GP + i] = acc[cc];
          } else {
#pragma unroll
            for (int cc = 0; cc < NB; ++cc) Lt[(size_t)(p0 + cc) * GP + i] = (cc <= i - p0) ? D[(i - p0) * LDD + cc] : 0.0;
          }
        }
      }
    }
    // threads 0..31 hold the log-diagonal partial sums and the not-PD flags
    if (tid < 32) {
      ldsum = warp_sum(ldsum);
      bad = __any_sync(0xffffffffu, bad);
      if (tid == 0) {
        const double ld = (ok && !bad) ? 2.0 * ldsum : __longlong_as_double(0x7ff8000000000000LL);
        out[c] = ldp ? 0.5 * (*ldp - ld) : ld;
      }
    }
  }
}

// best[0] = argmax_c I[c] (first index on ties, NaNs ignored; -1 when every value is NaN or C == 0).
__global__ void __launch_bounds__(1024) k_argmax(const double* __restrict__ I, long C, long* __restrict__ best) {
  __shared__ double sv[32];
  __shared__ long si[32];
  double bv = 0.0;
  long bi = -1;
  for (long c = threadIdx.x; c < C; c += 1024) {
    const double v = I[c];
    if (v == v && (bi < 0 || v > bv)) { bv = v; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    bv = sv[threadIdx.x];
    bi = si[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    if (threadIdx.x == 0) best[0] = bi;
  }
}
