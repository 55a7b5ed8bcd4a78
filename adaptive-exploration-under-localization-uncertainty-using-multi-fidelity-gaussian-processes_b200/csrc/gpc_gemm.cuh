// gpc_gemm.cuh -- FP64 tensor-core (DMMA.8x8x4) tile contraction shared by every dense step:
// Cholesky panel solve + trailing update, triangular inverse, L^-1 K* (variance / IG) and the
// full-covariance SYRK.
//
// One CTA = 512 threads = 16 warps (4 x 4), CTA tile 128 x 128, warp tile 32 x 32, BK = 16.
// Operands are staged global -> shared with 16-byte cp.async in a 4-stage ring; shared tiles are
// padded so that every DMMA fragment load (one double per lane) is bank-conflict free:
//   k-contiguous tile  [128 rows][LDT = 20]   lane (r = lane/4, c = lane%4) reads row*20 + c
//   n-contiguous tile  [ 16 k   ][LDN = 132]  lane reads k*132 + n
// Per k4-step a warp issues 4 + 4 LDS.64 and 16 DMMA; at one DMMA per 16 clk per SM sub-partition
// the tensor pipe, not shared memory or issue, is the limiter (see DESIGN.md roofline).
#pragma once
#include "gpc_common.cuh"

namespace gpcg {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int NTHREADS = 512;
constexpr int STAGES = 4;
constexpr int LDT = 20;
constexpr int LDN = 132;
constexpr int A_STAGE = BM * LDT;                   // doubles
constexpr int B_STAGE = BN * LDT;                   // >= 16 * LDN
constexpr int STAGE_DOUBLES = A_STAGE + B_STAGE;    // 5120
constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * 8;  // 163840

// A: pointer to the first row of the 128-row operand tile (k contiguous, leading dim lda).
// B (k-contiguous, B_N = false): pointer to the first row of the 128-row operand tile.
// B (n-contiguous, B_N = true) : pointer to column n0 of row k = 0; rows are k.
template <bool B_N>
__device__ __forceinline__ void load_stage(double* As, double* Bs, const double* __restrict__ A, long lda,
                                           const double* __restrict__ B, long ldb, int k0, int tid) {
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int r = (tid >> 3) + 64 * p, c = tid & 7;
    cp_async16(As + r * LDT + 2 * c, A + (long)r * lda + k0 + 2 * c);
  }
  if (!B_N) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int r = (tid >> 3) + 64 * p, c = tid & 7;
      cp_async16(Bs + r * LDT + 2 * c, B + (long)r * ldb + k0 + 2 * c);
    }
  } else {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int kr = (tid >> 6) + 8 * p, c = tid & 63;
      cp_async16(Bs + kr * LDN + 2 * c, B + (long)(k0 + kr) * ldb + 2 * c);
    }
  }
}

template <bool B_N>
__device__ __forceinline__ void compute_stage(const double* As, const double* Bs, double (&acc)[4][4][2],
                                              int wm, int wn, int lane) {
  const int lr = lane >> 2, lc = lane & 3;
  const double* Ap = As + (wm * 32 + lr) * LDT + lc;
  const double* Bp = B_N ? (Bs + lc * LDN + wn * 32 + lr) : (Bs + (wn * 32 + lr) * LDT + lc);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    double a[4], b[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) a[f] = Ap[f * 8 * LDT + kk * 4];
#pragma unroll
    for (int g = 0; g < 4; ++g) b[g] = B_N ? Bp[kk * 4 * LDN + g * 8] : Bp[g * 8 * LDT + kk * 4];
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
      for (int g = 0; g < 4; ++g) dmma884(acc[f][g][0], acc[f][g][1], a[f], b[g]);
  }
}

// acc += A[:, kbeg:kend] * op(B)[kbeg:kend, :]; kbeg/kend multiples of 16.  On return all async
// copies have landed and the CTA is synchronised, so `smem` may be reused by the epilogue.
template <bool B_N>
__device__ __forceinline__ void mainloop(const double* __restrict__ A, long lda, const double* __restrict__ B,
                                         long ldb, int kbeg, int kend, double (&acc)[4][4][2], double* smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int nk = (kend - kbeg) / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage<B_N>(smem + s * STAGE_DOUBLES, smem + s * STAGE_DOUBLES + A_STAGE, A, lda, B, ldb,
                                kbeg + s * BK, tid);
    cp_async_commit();
  }
  for (int it = 0; it < nk; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nx = it + STAGES - 1;
    if (nx < nk) {
      double* st = smem + (nx % STAGES) * STAGE_DOUBLES;
      load_stage<B_N>(st, st + A_STAGE, A, lda, B, ldb, kbeg + nx * BK, tid);
    }
    cp_async_commit();
    const double* st = smem + (it % STAGES) * STAGE_DOUBLES;
    compute_stage<B_N>(st, st + A_STAGE, acc, wm, wn, lane);
  }
  cp_async_wait<0>();
  __syncthreads();
}

__device__ __forceinline__ void zero_acc(double (&acc)[4][4][2]) {
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int g = 0; g < 4; ++g) acc[f][g][0] = acc[f][g][1] = 0.0;
}

// Accumulator element (f, g, e) of this lane sits at tile row/col:
__device__ __forceinline__ int acc_row(int f) { return ((threadIdx.x >> 5) >> 2) * 32 + f * 8 + ((threadIdx.x & 31) >> 2); }
__device__ __forceinline__ int acc_col(int g) { return ((threadIdx.x >> 5) & 3) * 32 + g * 8 + 2 * (threadIdx.x & 3); }

// C[tile] = beta * C[tile] + alpha * acc   (C row-major, ldc; tile origin already applied)
__device__ __forceinline__ void store_tile(double* __restrict__ C, long ldc, const double (&acc)[4][4][2],
                                           double alpha, double beta) {
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const int r = acc_row(f);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      double2* p = reinterpret_cast<double2*>(C + (long)r * ldc + acc_col(g));
      double2 v;
      if (beta != 0.0) {
        v = *p;
        v.x = fma(alpha, acc[f][g][0], beta * v.x);
        v.y = fma(alpha, acc[f][g][1], beta * v.y);
      } else {
        v.x = alpha * acc[f][g][0];
        v.y = alpha * acc[f][g][1];
      }
      *p = v;
    }
  }
}

// rowsum[r] = sum_c acc(r, c)^2 over the 128 columns of the tile; result valid in threads 0..127
// (thread t returns row t).  Uses `smem` (>= 512 doubles) as scratch; CTA must be converged.
__device__ __forceinline__ double rowsumsq_tile(const double (&acc)[4][4][2], double* smem) {
  const int lane = threadIdx.x & 31, wn = (threadIdx.x >> 5) & 3;
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    double s = 0.0;
#pragma unroll
    for (int g = 0; g < 4; ++g) s = fma(acc[f][g][0], acc[f][g][0], fma(acc[f][g][1], acc[f][g][1], s));
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if ((lane & 3) == 0) smem[wn * 128 + acc_row(f)] = s;
  }
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 128) t = (smem[threadIdx.x] + smem[128 + threadIdx.x]) + (smem[256 + threadIdx.x] + smem[384 + threadIdx.x]);
  return t;
}

}  // namespace gpcg

// ------------------------------------------------------------------------------------------
// The small-tile shape: 64 x 64 CTA tile, 4 warps of 32 x 32, BK = 16, two cp.async stages,
// 40 KB of shared memory -> 4 CTAs per SM.  Measured 35.0 TFLOP/s against 30.2 for the shape
// above (profiles/r01/gemm_variants_r01.txt): several small CTAs per SM cover each other's
// barrier and pipeline-fill bubbles.  Used by the factorisation kernels; k_vt has its own copy
// of this loop because it chains two k-segments through one pipeline.
// ------------------------------------------------------------------------------------------
namespace gpc64 {

constexpr int BM = 64, BN = 64, BK = 16, NT = 128;
constexpr int LDT = 20, LDN = 68;
constexpr int A_STAGE = BM * LDT;                 // 1280 doubles
constexpr int STAGE = A_STAGE + BN * LDT;         // B slot holds 64 x 20 (k-contiguous) or 16 x 68 (n-contiguous)
constexpr int SMEM_BYTES = 2 * STAGE * 8;         // 40960
constexpr int SMEM_BYTES_DEEP = 4 * STAGE * 8;    // 81920: the four-stage ring of the latency-critical launches

template <bool B_N>
__device__ __forceinline__ void load_stage(double* As, double* Bs, const double* __restrict__ A, long lda,
                                           const double* __restrict__ B, long ldb, int k0, int tid) {
#pragma unroll
  for (int c = tid; c < BM * (BK / 2); c += NT) {
    const int r = c >> 3, q = c & 7;
    cp_async16(As + r * LDT + 2 * q, A + (long)r * lda + k0 + 2 * q);
  }
  if (!B_N) {
#pragma unroll
    for (int c = tid; c < BN * (BK / 2); c += NT) {
      const int r = c >> 3, q = c & 7;
      cp_async16(Bs + r * LDT + 2 * q, B + (long)r * ldb + k0 + 2 * q);
    }
  } else {
#pragma unroll
    for (int c = tid; c < BK * (BN / 2); c += NT) {
      const int kr = c >> 5, q = c & 31;
      cp_async16(Bs + kr * LDN + 2 * q, B + (long)(k0 + kr) * ldb + 2 * q);
    }
  }
}

// acc += A[:, kbeg:kend] * op(B)[kbeg:kend, :]; kbeg / kend multiples of 16.  A: first row of the
// 64-row tile (k contiguous).  B (B_N = false): first row of the 64-row tile (k contiguous);
// B (B_N = true): column n0 of row k = 0 (rows are k).  On return the CTA is synchronised.
// NS = ring depth: 2 (40 KB, 4 CTAs per SM) for the throughput-bound launches that fill the GPU; 4 (80 KB) for the
// small launches on the critical path of the factorisation (a panel solve is 30 CTAs on 148 SMs: nothing else on
// the SM covers the load latency, so three k-steps are kept in flight instead of one).
template <bool B_N, int NS = 2>
__device__ __forceinline__ void mainloop(const double* __restrict__ A, long lda, const double* __restrict__ B,
                                         long ldb, int kbeg, int kend, double (&acc)[4][4][2], double* smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1, lr = lane >> 2, lc = lane & 3;
  const int nk = (kend - kbeg) / BK;
#pragma unroll
  for (int s = 0; s < NS - 1; ++s) {
    if (s < nk) load_stage<B_N>(smem + s * STAGE, smem + s * STAGE + A_STAGE, A, lda, B, ldb, kbeg + s * BK, tid);
    cp_async_commit();
  }
  for (int it = 0; it < nk; ++it) {
    cp_async_wait<NS - 2>();
    __syncthreads();
    if (it + NS - 1 < nk) {
      double* st = smem + ((it + NS - 1) % NS) * STAGE;
      load_stage<B_N>(st, st + A_STAGE, A, lda, B, ldb, kbeg + (it + NS - 1) * BK, tid);
    }
    cp_async_commit();
    const double* As = smem + (it % NS) * STAGE;
    const double* Ap = As + (wm * 32 + lr) * LDT + lc;
    const double* Bp = B_N ? (As + A_STAGE + lc * LDN + wn * 32 + lr) : (As + A_STAGE + (wn * 32 + lr) * LDT + lc);
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) a[f] = Ap[f * 8 * LDT + kk * 4];
#pragma unroll
      for (int g = 0; g < 4; ++g) b[g] = B_N ? Bp[kk * 4 * LDN + g * 8] : Bp[g * 8 * LDT + kk * 4];
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int g = 0; g < 4; ++g) dmma884(acc[f][g][0], acc[f][g][1], a[f], b[g]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

__device__ __forceinline__ int acc_row(int f) { return (threadIdx.x >> 6) * 32 + f * 8 + ((threadIdx.x & 31) >> 2); }
__device__ __forceinline__ int acc_col(int g) { return ((threadIdx.x >> 5) & 1) * 32 + g * 8 + 2 * (threadIdx.x & 3); }

// C[tile] = beta * C[tile] + alpha * acc  (tile origin already applied)
__device__ __forceinline__ void store_tile(double* __restrict__ C, long ldc, const double (&acc)[4][4][2],
                                           double alpha, double beta) {
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const int r = acc_row(f);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      double2* p = reinterpret_cast<double2*>(C + (long)r * ldc + acc_col(g));
      double2 v;
      if (beta != 0.0) {
        v = *p;
        v.x = fma(alpha, acc[f][g][0], beta * v.x);
        v.y = fma(alpha, acc[f][g][1], beta * v.y);
      } else {
        v.x = alpha * acc[f][g][0];
        v.y = alpha * acc[f][g][1];
      }
      *p = v;
    }
  }
}

}  // namespace gpc64
