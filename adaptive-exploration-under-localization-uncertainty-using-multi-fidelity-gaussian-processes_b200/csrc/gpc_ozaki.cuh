// gpc_ozaki.cuh -- the L^-1 K* contraction on the 5th-generation tensor cores (tcgen05, INT8).
//
// FP64 on sm_100a has no tcgen05 kind: DMMA and DFMA share one FP64 pipe that peaks at 37 TFLOP/s
// (profiles/r01/pipes_r01.txt), so k_vt (gpc_predict.cuh) is within 10 % of what FP64 hardware can
// do.  This file evaluates the SAME contraction V = K* X^T exactly-enough on the INT8 tensor cores
// (4.56 POP/s measured, profiles/r01/umma_i8_rate_r01.txt) with the Ozaki splitting scheme:
//
//   x = scale * 2^-48 * v,  v = rint(x / scale * 2^48) = sum_{p<S} d_p 256^(S-1-p),  d_p in [-128, 127]
//   (balanced base-256 digits, S = 6 int8 "slices" per operand; the balanced range is
//   [-0.50196, 0.49804] 2^48, so with |x| / scale <= 0.4975 the bytes of v + 0x808080808080 are the digits + 128 and
//   slicing is one add and byte permutes.  Scales are NOT rounded to powers of two: every factor of two of headroom
//   costs one bit of the result -- the terms dropped below carry 2^-45 sA sB), so that
//   sum_k a_k b_k = sA sB 2^-96 sum_{p,q} 256^(10-p-q) sum_k d^A_pk d^B_qk.
// Every digit GEMM C_pq = D^A_p (D^B_q)^T is exact in int32 (|d| <= 128, up to 6 pairs per
// accumulator: K <= 21845); the pairs with p + q = t share one TMEM accumulator; pairs with
// p + q >= S (below 2^-48 of the operand scales) are dropped.  S (S + 1) / 2 = 21 digit GEMMs replace
// one FP64 GEMM.
//
// Operand images.  Digits are stored in global memory already in the shared-memory image the MMA
// consumes (K-major, no swizzle, 8 x 16-byte core matrices): a (rows x 64-byte) k-block of one
// slice is [rows/8][4][8][16 B] (SBO = 512 B, LBO = 128 B).  All S slices of a k-block are
// contiguous, so one stage of the pipeline is two 1-D bulk copies (TMA, cp.async.bulk).
//   A image  [m_pad/128][n_pad/64][S][8 KB]   test rows  x train index (digits of K* / sA)
//   B image  [n_pad/64 ][n_pad/64][S][4 KB]   rows of X x train index (digits of X_i,: / sB_i)
//
// Kernel k_vt_i8: persistent, one CTA per SM, 320 threads:
//   warp 0   TMA producer (one lane): 3-stage ring of (A k-block, B k-block) = 72 KB per stage
//   warp 1   MMA issuer (one lane): 16 tcgen05.mma.kind::i8 (M = 128, N = 64..256, K = 32) per stage into
//            S accumulators of 64 TMEM columns; tcgen05.commit frees the stage / publishes the tile
//   warps 2-9 epilogue (two per TMEM lane quarter, 32 columns each): tcgen05.ld the S int32 levels (software-pipelined), recombine in int64, scale to FP64 and
//            reduce sum_i V(n, i)^2 per test row (one thread owns a row: no shuffles), store V, or emit V's own
//            digit image for the next INT8 product (information gain).
// Schedules: triangular -- item (mt, p) = train tiles nb2-1-p then p of test tile mt (V = K* X^T, X lower
// triangular) -- or full-K (Gram and cross products of V); see the comment above k_vt_i8.
#pragma once
#include "gpc_common.cuh"

namespace gpoz {

constexpr int S = 6;                 // int8 slices (base-256 digits) per operand
constexpr int BK = 64;               // k-block (bytes = k values)
constexpr int TM = 128, TN = 64;     // MMA tile: 128 test rows x 64 train rows
constexpr int A_SLICE = TM * BK;     // 8192 B
constexpr int B_SLICE = TN * BK;     // 4096 B
constexpr int A_STAGE = S * A_SLICE, B_STAGE = S * B_SLICE;
constexpr int STAGE_BYTES = A_STAGE + B_STAGE;   // 73728
constexpr int STAGES = 3;
constexpr int MAX_K = 16384;         // int32 accumulators: 6 pairs x K x 128^2 < 2^31
constexpr double DIGIT_MUL = 281474976710656.0;           // 2^48
constexpr double SCALE_HEADROOM = 1.0 / 0.4975;           // scale = max|x| * this: |v| <= 0.4975 2^48, inside the balanced range
constexpr long long DIGIT_BIAS = 0x808080808080LL;        // 128 (256^6 - 1) / 255
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + alignment slack
constexpr int NT = 320;               // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)
constexpr uint32_t TMEM_COLS = 512;
constexpr double COMB_SCALE = 1.0 / 72057594037927936.0;  // 2^-56

__device__ __forceinline__ void mbar_wait_guarded(uint64_t* bar, uint32_t parity) {
  // a broken barrier protocol must abort the kernel, never hang the GPU
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    if (spin > (1u << 28)) __trap();
  }
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  // K-major, SWIZZLE_NONE: leading (k-chunk) offset 128 B, stride (8-row group) offset 512 B, version 1
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}

// tcgen05.wait::ld with the loaded registers as in/out operands, so that no use of them can be
// scheduled above the wait.
__device__ __forceinline__ void tmem_wait3(int32_t (&a)[8], int32_t (&b)[8], int32_t (&c)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                 "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                 "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7])
               :
               : "memory");
}

// exact int64 -> double for |x| < 2^51: one 64-bit integer add and one DADD instead of I2F.F64.S64
__device__ __forceinline__ double i64_to_f64(long long x) {
  return __longlong_as_double(x + 0x4338000000000000LL) - 6755399441055744.0;  // 2^52 + 2^51
}

// The S digit bytes of 16 consecutive values for slice p, packed as one 16-byte vector: u[] holds
// v + DIGIT_BIAS (bytes = digits + 128, byte 0 least significant = slice S-1).
__device__ __forceinline__ uint4 pack_slice(const unsigned long long (&u)[16], int p) {
  const int b = S - 1 - p;  // byte index inside the 48-bit value
  uint32_t w[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t x[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const unsigned long long v = u[4 * q + e];
      const uint32_t word = b < 4 ? (uint32_t)v : (uint32_t)(v >> 32);
      x[e] = (word >> (8 * (b & 3))) & 255u;
    }
    w[q] = (x[0] | (x[1] << 8) | (x[2] << 16) | (x[3] << 24)) ^ 0x80808080u;  // digit = byte - 128
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

}  // namespace gpoz

// ------------------------------------------------------------------------------------------
// Digits of the rows of X = L^-1 (lower triangular, row-major, ld).  Row i is scaled by
// sB[i] = max|X(i, :)| / 0.4975.  One warp per row; grid = nrows / 8, 256 threads.  tri = 0: general rows
// (the grid's V for the cross products of the information gain).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_slice_rows(const double* __restrict__ X, long ld, long n_pad,
                                                    int8_t* __restrict__ Bimg, double* __restrict__ sB, long nrows,
                                                    int tri) {
  using namespace gpoz;
  const long i = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= nrows) return;
  const double* row = X + i * ld;
  const long klast = tri ? i : n_pad - 1;      // tri: lower-triangular operand, k-blocks beyond the diagonal stay zero
  double mx = 0.0;
  for (long k = lane; k <= klast; k += 32) mx = fmax(mx, fabs(row[k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const double scale = (mx > 0.0) ? mx * SCALE_HEADROOM : 1.0;   // the top balanced digit cannot overflow
  if (lane == 0) sB[i] = scale;
  const double mul = DIGIT_MUL / scale;
  const long jb = i >> 6, r = i & 63;
  const long nkb = n_pad >> 6;
  // every 16-byte chunk (16 consecutive k) of every slice
  for (long c = lane; c < ((klast >> 6) + 1) * 4; c += 32) {
    const long kb = c >> 2, cc = c & 3;
    unsigned long long uu[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const long k = kb * 64 + cc * 16 + u;
      const double x = (k <= klast) ? row[k] : 0.0;
      uu[u] = (unsigned long long)(__double2ll_rn(x * mul) + DIGIT_BIAS);
    }
    int8_t* base = Bimg + ((jb * nkb + kb) * S) * (long)B_SLICE + (r >> 3) * 512 + cc * 128 + (r & 7) * 16;
#pragma unroll
    for (int p = 0; p < S; ++p) *reinterpret_cast<uint4*>(base + (long)p * B_SLICE) = pack_slice(uu, p);
  }
}

// Byte b (0 = least significant) of eight 48-bit values -> the 8 digit bytes of one slice (PRMT).
template <int B>
__device__ __forceinline__ uint2 pack_bytes8(const unsigned long long (&u)[8]) {
  uint32_t w[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) w[e] = B < 4 ? (uint32_t)u[e] : (uint32_t)(u[e] >> 32);
  constexpr uint32_t sel = (uint32_t)(B & 3) | ((uint32_t)(4 + (B & 3)) << 4);   // bytes B of x, B of y -> low half
  const uint32_t t01 = __byte_perm(w[0], w[1], sel), t23 = __byte_perm(w[2], w[3], sel);
  const uint32_t t45 = __byte_perm(w[4], w[5], sel), t67 = __byte_perm(w[6], w[7], sel);
  return make_uint2(__byte_perm(t01, t23, 0x5410) ^ 0x80808080u, __byte_perm(t45, t67, 0x5410) ^ 0x80808080u);
}

// ------------------------------------------------------------------------------------------
// K* assembly straight into the A digit image (+ posterior mean, + optional mean gradients): the
// FP64 cross-covariance never goes to HBM -- 7 bytes per element are written instead of 8.
// CTA = 128 threads = the 128 test rows of one MMA tile, one row per thread; the CTA covers 256
// train columns (four 64-wide k-blocks) staged by 1-D TMA bulk copies.  All lanes of a warp work
// on the SAME train column at a time (shared-memory reads are broadcasts, the fidelity of the
// column is warp-uniform), and 16 consecutive columns of a row become one 16-byte store per slice
// -- eight lanes fill a 128-byte core matrix.  The mean needs no cross-lane reduction.
//   k = sum_{m <= min(fi, fj)} w_i[m] coef[fj][m] base_m(x - x'),  w_i[m] = coef[fi][m] var[m]
// grid (m_pad / 128, n_pad / 256).   meanpart [n_pad/256][m_pad], gradpart [n_pad/256][3][m_pad].
// ------------------------------------------------------------------------------------------
constexpr int KI_COLS = 256;

template <bool WITH_GRAD>
__global__ void __launch_bounds__(128) k_kstar_i8(const __grid_constant__ GpcHyp h, const double* __restrict__ Xt,
                                                  const double* __restrict__ alpha, long N, long n_pad,
                                                  const double* __restrict__ Xs4, long M, long m_pad, double sA,
                                                  int8_t* __restrict__ Aimg, double* __restrict__ meanpart,
                                                  double* __restrict__ gradpart) {  // grid (m_pad / 128, nchunks)
  using namespace gpoz;
  __shared__ __align__(128) double tr[5][KI_COLS];
  __shared__ double hil[GPC_MAXF][4];
  __shared__ double hcoef[GPC_MAXF][GPC_MAXF];
  __shared__ double wS[GPC_MAXF][128];
  __shared__ double T64[64];                       // 2^(j / 64): the table of gpc_exp_neg_tab_w
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31;
  const int mt = blockIdx.x;
  if (tid < 64) T64[tid] = exp2((double)tid * 0.015625);
  const long j0 = (long)blockIdx.y * KI_COLS;
  const long nkb = n_pad >> 6;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int ncol = (int)((n_pad - j0) < KI_COLS ? (n_pad - j0) : KI_COLS);  // multiple of 128
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)ncol * 8u;
    mbar_expect_tx(&bar, 5u * bytes);
#pragma unroll
    for (int c = 0; c < 4; ++c) tma_bulk_g2s(&tr[c][0], Xt + (long)c * n_pad + j0, bytes, &bar);
    tma_bulk_g2s(&tr[4][0], alpha + j0, bytes, &bar);
  }
  if (tid < GPC_MAXF * 3) hil[tid / 3][tid % 3] = h.inv_l[tid / 3][tid % 3] * (h.base == 0 ? 0.70710678118654752440 : 1.0);
  if (tid < GPC_MAXF * GPC_MAXF) hcoef[tid >> 2][tid & 3] = h.coef[tid >> 2][tid & 3];
  const long n = (long)mt * TM + tid;
  double ax = 0, ay = 0, az = 0, af = -1.0;
  if (n < M) { ax = Xs4[n * 4]; ay = Xs4[n * 4 + 1]; az = Xs4[n * 4 + 2]; af = Xs4[n * 4 + 3]; }
  const bool live = n < M && af >= 0.0;
  const int fi = gpc_fid(h, af);
  const int F = h.F, base = h.base;
#pragma unroll
  for (int m = 0; m < GPC_MAXF; ++m) wS[m][tid] = (live && m <= fi && m < F) ? h.coef[fi][m] * h.var[m] : 0.0;
  const int fimax = __reduce_max_sync(0xffffffffu, live ? fi : 0);
  __syncthreads();
  mbar_wait(&bar, 0);
  // fidelity of every staged column as an int, in place (low word of the double slot)
  int* fi32 = reinterpret_cast<int*>(&tr[3][0]);
  for (int j = tid; j < ncol; j += 128) {
    const int f = gpc_fid(h, tr[3][j]);
    fi32[2 * j] = (j0 + j < N) ? f : -1;  // -1: padding column, every coefficient is zero
  }
  __syncthreads();
  const double mul = DIGIT_MUL / sA;
  double mu = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;
  const int r = tid;
#pragma unroll 1
  for (int c8 = 0; c8 < (ncol >> 3); ++c8) {  // 8 columns at a time = half of a 16-byte chunk of every slice
    // the AR1 sum runs over m <= min(max test fidelity in the warp, max train fidelity of the 8 columns);
    // coef[fj][m] is zero for m > fj, w[m] is zero for m > fi, so over-running a term adds exact zeros
    int fjmax = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) fjmax = max(fjmax, fi32[2 * (c8 * 8 + u)]);
    const int mmc = fjmax < fimax ? fjmax : fimax;
    double k[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) k[u] = 0.0;
#pragma unroll 1
    for (int m = 0; m <= mmc; ++m) {
      const double c0 = hil[m][0], c1 = hil[m][1], c2 = hil[m][2], wm = wS[m][tid];
      double q[8], e[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = c8 * 8 + u;
        const double sx = (tr[0][j] - ax) * c0, sy = (tr[1][j] - ay) * c1, sz = (tr[2][j] - az) * c2;
        q[u] = fma(sx, sx, fma(sy, sy, sz * sz));
      }
      if (base == 0) {
        gpc_exp_neg_tab_w<4>(q, e, T64);
        gpc_exp_neg_tab_w<4>(q + 4, e + 4, T64);
      } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = 1.7320508075688772 * sqrt(q[u]);
        gpc_exp_neg_tab_w<4>(q, e, T64);
        gpc_exp_neg_tab_w<4>(q + 4, e + 4, T64);
#pragma unroll
        for (int u = 0; u < 8; ++u) e[u] *= 1.0 + q[u];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int fj = fi32[2 * (c8 * 8 + u)];
        k[u] = fma(wm * (fj >= 0 ? hcoef[fj][m] : 0.0), e[u], k[u]);
      }
    }
    unsigned long long v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = c8 * 8 + u;
      const double ka = k[u] * tr[4][j];
      mu += ka;
      if (WITH_GRAD) {
        g0 = fma(ka, tr[0][j] - ax, g0);
        g1 = fma(ka, tr[1][j] - ay, g1);
        g2 = fma(ka, tr[2][j] - az, g2);
      }
      // rint(k * mul) + bias without F2I: adding 2^52 + 2^51 leaves the rounded integer in the mantissa (|k * mul| < 2^48)
      v[u] = (unsigned long long)(__double_as_longlong(fma(k[u], mul, 6755399441055744.0)) + (DIGIT_BIAS - 0x4338000000000000LL));
    }
    const long kb = (j0 >> 6) + (c8 >> 3);
    int8_t* dst = Aimg + (((long)mt * nkb + kb) * S) * (long)A_SLICE + (r >> 3) * 512 + ((c8 >> 1) & 3) * 128 +
                  (r & 7) * 16 + (c8 & 1) * 8;
    // slice p holds byte S - 1 - p of the 48-bit value
    *reinterpret_cast<uint2*>(dst + 0L * A_SLICE) = pack_bytes8<5>(v);
    *reinterpret_cast<uint2*>(dst + 1L * A_SLICE) = pack_bytes8<4>(v);
    *reinterpret_cast<uint2*>(dst + 2L * A_SLICE) = pack_bytes8<3>(v);
    *reinterpret_cast<uint2*>(dst + 3L * A_SLICE) = pack_bytes8<2>(v);
    *reinterpret_cast<uint2*>(dst + 4L * A_SLICE) = pack_bytes8<1>(v);
    *reinterpret_cast<uint2*>(dst + 5L * A_SLICE) = pack_bytes8<0>(v);
  }
  meanpart[(long)blockIdx.y * m_pad + n] = mu;
  if (WITH_GRAD) {
    double* gp = gradpart + (long)blockIdx.y * 3 * m_pad;
    gp[n] = g0 * h.inv_l[0][0] * h.inv_l[0][0];
    gp[m_pad + n] = g1 * h.inv_l[0][1] * h.inv_l[0][1];
    gp[2 * m_pad + n] = g2 * h.inv_l[0][2] * h.inv_l[0][2];
  }
  (void)lane;
}

// ------------------------------------------------------------------------------------------
// The contraction.  grid = min(#SMs, items), 320 threads, dynamic smem gpoz::SMEM_BYTES.
//   D(test tile mt, B row tile jb) = sum_kb A[mt][kb] B[jb][kb]^T  as 21 exact digit GEMMs.
// Schedules:  triangular (FULLK = false): item (mt, p) = B tiles nb2-1-p then p, k-blocks 0..jb (V = K* X^T with
//             X = L^-1 lower triangular; every item has nb2 + 1 k-steps);
//             full (FULLK = true): item (mt, jb), k-blocks 0..nkb-1 (Gram and cross products of V).
// Outputs:    OUT_SUMSQ   sumsq[jb][m_pad] = sum_i D(n, i)^2 per test row (posterior variance);
//             OUT_F64     out[n * ld_out + jb 64 + c] = D                       (full covariance / information gain);
//             OUT_DIGITS  the digit image of D in the A-image layout (scale 2^48 / dig_mul), so that V itself can
//                         be the operand of the next INT8 product without ever existing in FP64 in HBM.
// B addressing is generic (byte strides per test tile / row tile / k-block / digit slice): the standard image
// [jb][kb][S][4 KB] arrives as one 24 KB bulk copy per stage; for self products (Gram of a V tile) the two
// 64-row halves of the A image of the SAME tile serve as B (six 4 KB copies per stage).
// ------------------------------------------------------------------------------------------
enum { OUT_SUMSQ = 0, OUT_F64 = 1, OUT_DIGITS = 2 };

struct VtI8Args {
  const int8_t* Aimg;
  const int8_t* Bimg;
  long b_mt, b_j, b_k, b_p;   // byte strides of the B operand
  const double* sB;           // scale of every B row
  double sA;                  // scale of A
  int nkb;                    // k-blocks (of 64) in the A image
  int nb2;                    // B row tiles (of 64)
  long ld_out, m_pad;
  int n_items;
  double* out;                // OUT_F64
  int8_t* dig;                // OUT_DIGITS: [m_pad / 128][nkb_out][S][8 KB]
  int nkb_out;
  double dig_mul;             // 2^48 / (scale of the emitted digits)
  double* sumsq;              // OUT_SUMSQ
  int nlev;                   // digit levels used (= the kernel's NLEV template argument): 6 = full (21 digit GEMMs, FP64
                              // results), 4 = FP32-tolerance mode (pairs p + q <= 3: 10 GEMMs); the unused low-order slices
                              // are neither copied nor multiplied
};

template <bool FULLK>
__device__ __forceinline__ void vt_item(int item, int nb2, int npair, int& mt, int (&jbs)[2], int& nseg) {
  if (FULLK) {
    mt = item / nb2;
    jbs[0] = item - mt * nb2;
    jbs[1] = 0;
    nseg = 1;
  } else {
    mt = item / npair;
    const int p = item - mt * npair;
    jbs[0] = nb2 - 1 - p;
    jbs[1] = p;
    nseg = (jbs[1] < jbs[0]) ? 2 : 1;
  }
}

template <int OUT, bool FULLK, int NLEV>
__global__ void __launch_bounds__(gpoz::NT, 1) k_vt_i8(const __grid_constant__ VtI8Args a) {
  using namespace gpoz;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar, tmem_empty_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ double sb_tile[2][TN];   // row scales of the tile being drained (x 2^-56 sA), double-buffered
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb2 = a.nb2, n_items = a.n_items;
  const int npair = (nb2 + 1) / 2;
  const long nkb_total = a.nkb;
  const long m_pad = a.m_pad;
  const double sA = a.sA;
  const double* __restrict__ sB = a.sB;
  constexpr int nlev = NLEV;   // compile-time: the MMA issue loop of the one issuing thread must unroll completely

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    mbar_init(&tmem_empty_bar, 8);  // one arrive per epilogue warp
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int mt, jbs[2], nseg;
        vt_item<FULLK>(item, nb2, npair, mt, jbs, nseg);
        for (int sg = 0; sg < nseg; ++sg) {
          const int jb = jbs[sg];
          const int kend = FULLK ? (int)nkb_total : jb + 1;
          for (int kb = 0; kb < kend; ++kb, ++it) {
            const int st = it % STAGES;
            if (it >= STAGES) mbar_wait_guarded(&empty_bar[st], ((it / STAGES) - 1) & 1);
            uint8_t* dst = smem + st * STAGE_BYTES;
            // slice p of a k-block is the p-th most significant digit: the first nlev slices are a contiguous prefix
            mbar_expect_tx(&full_bar[st], (uint32_t)nlev * (A_SLICE + B_SLICE));
            tma_bulk_g2s(dst, a.Aimg + (((long)mt * nkb_total + kb) * S) * (long)A_SLICE, (uint32_t)nlev * A_SLICE, &full_bar[st]);
            const int8_t* bsrc = a.Bimg + (long)mt * a.b_mt + (long)jb * a.b_j + (long)kb * a.b_k;
            if (a.b_p == (long)B_SLICE) {
              tma_bulk_g2s(dst + A_STAGE, bsrc, (uint32_t)nlev * B_SLICE, &full_bar[st]);
            } else {
#pragma unroll
              for (int p = 0; p < nlev; ++p) tma_bulk_g2s(dst + A_STAGE + p * B_SLICE, bsrc + (long)p * a.b_p, B_SLICE, &full_bar[st]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // c = S32 (2) @4, a = b = INT8 (1) @7 / @10, K-major, N >> 3 @17, M >> 4 @24
      const uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TM >> 4) << 24);
      uint32_t it = 0, tile = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int mt, jbs[2], nseg;
        vt_item<FULLK>(item, nb2, npair, mt, jbs, nseg);
        (void)mt;
        for (int sg = 0; sg < nseg; ++sg, ++tile) {
          const int jb = jbs[sg];
          const int kend = FULLK ? (int)nkb_total : jb + 1;
          if (tile > 0) mbar_wait_guarded(&tmem_empty_bar, (tile - 1) & 1);  // epilogue drained the accumulators
          asm volatile("tcgen05.fence::after_thread_sync;");
          for (int kb = 0; kb < kend; ++kb, ++it) {
            const int st = it % STAGES;
            mbar_wait_guarded(&full_bar[st], (it / STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t sa = smem_u32(smem + st * STAGE_BYTES), sb = sa + A_STAGE;
            // Digit pair (pp, qq) accumulates into TMEM region t = pp + qq (64 columns each).  The B
            // slices of a stage are contiguous 64-row blocks with the same row-group stride, so ONE MMA
            // of width 64 n covers the pairs (pp, q0 .. q0 + n - 1) and lands in regions t0 .. t0 + n - 1:
            // 8 wide MMAs per k-step instead of 21 narrow ones -- the A digits are fetched from shared
            // memory 8 times instead of 21 (shared-memory bandwidth is what bounds this kernel).
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
              for (int pp = 0; pp < nlev; ++pp) {
                int q0 = 0;
                while (q0 < nlev - pp) {
                  const int nsl = (nlev - pp - q0) > 4 ? 4 : (nlev - pp - q0);  // slices in this MMA (N <= 256)
                  const uint32_t idesc = idesc0 | ((uint32_t)((nsl * TN) >> 3) << 17);
                  umma_i8(tmem_base + (pp + q0) * TN, umma_desc(sa + pp * A_SLICE + kk * 256),
                          umma_desc(sb + q0 * B_SLICE + kk * 256), idesc, (kb > 0 || kk > 0 || pp > 0) ? 1u : 0u);
                  q0 += nsl;
                }
              }
            }
            umma_commit(&empty_bar[st]);  // the stage may be refilled once these MMAs have read it
          }
          umma_commit(&tmem_full_bar);    // all MMAs of this tile done: accumulators complete
        }
      }
    }
  } else {
    // ===== epilogue warps (2..9): TMEM lane quarter = warp % 4, column half = (warp - 2) / 4; one thread owns one
    // test row of one 32-column half of the tile =====
    // The MMA issuer cannot start the next tile before the accumulators are drained (6 x 64 of the 512
    // TMEM columns: no room for a second set), so this loop is on the critical path: two warps per lane quarter
    // halve it, the row scales are staged while the MMAs still run, the TMEM loads of the next half-group are in
    // flight during the arithmetic of the current one, and the accumulators are released right after the last load.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const int et = tid - 64;
    const int cbeg = half * (TN / 2), cend = cbeg + TN / 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const double cs = COMB_SCALE * sA;
    uint32_t tile = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int mt, jbs[2], nseg;
      vt_item<FULLK>(item, nb2, npair, mt, jbs, nseg);
      for (int sg = 0; sg < nseg; ++sg, ++tile) {
        const int jb = jbs[sg];
        double* sbt = sb_tile[tile & 1];
        if (et < TN) sbt[et] = sB[(long)jb * TN + et] * cs;   // V = comb 2^-56 sA sB[i]
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mbar_wait_guarded(&tmem_full_bar, tile & 1);
        asm volatile("tcgen05.fence::after_thread_sync;");
        double ss = 0.0;
        const long n = (long)mt * TM + row;
        int32_t v0[8], v1[8], v2[8], w0[8], w1[8], w2[8];
        tmem_ld8(lane_addr + 0 * TN + cbeg, v0);
        tmem_ld8(lane_addr + 1 * TN + cbeg, v1);
        tmem_ld8(lane_addr + 2 * TN + cbeg, v2);
#pragma unroll 1
        for (int c0 = cbeg; c0 < cend; c0 += 8) {
          tmem_wait3(v0, v1, v2);
          // levels >= nlev were never accumulated (their TMEM columns hold stale data): they count as zero
          if (nlev > 3) tmem_ld8(lane_addr + 3 * TN + c0, w0);
          if (nlev > 4) tmem_ld8(lane_addr + 4 * TN + c0, w1);
          if (nlev > 5) tmem_ld8(lane_addr + 5 * TN + c0, w2);
          long long hi[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) hi[j] = (long long)v0[j] * 65536LL + (long long)v1[j] * 256LL + (long long)v2[j];
          tmem_wait3(w0, w1, w2);
          if (nlev <= 5) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              w2[j] = 0;
              if (nlev <= 4) w1[j] = 0;
              if (nlev <= 3) w0[j] = 0;
            }
          }
          if (c0 + 8 < cend) {
            tmem_ld8(lane_addr + 0 * TN + c0 + 8, v0);
            tmem_ld8(lane_addr + 1 * TN + c0 + 8, v1);
            tmem_ld8(lane_addr + 2 * TN + c0 + 8, v2);
          } else {
            // every accumulator column of this warp has been read: hand TMEM back to the MMA issuer
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar);
          }
          // comb = sum_t C_t 256^(5-t) = hi 256^3 + lo
          double vv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const long long lo = (long long)w0[j] * 65536LL + (long long)w1[j] * 256LL + (long long)w2[j];
            const double comb = fma(i64_to_f64(hi[j]), 16777216.0, i64_to_f64(lo));
            vv[j] = comb * sbt[c0 + j];
            ss = fma(vv[j], vv[j], ss);
          }
          if (OUT == OUT_F64) {
            double* dst = a.out + n * a.ld_out + (long)jb * TN + c0;
#pragma unroll
            for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(dst + j) = make_double2(vv[j], vv[j + 1]);
          }
          if (OUT == OUT_DIGITS) {
            unsigned long long dv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dv[j] = (unsigned long long)(__double_as_longlong(fma(vv[j], a.dig_mul, 6755399441055744.0)) +
                                           (DIGIT_BIAS - 0x4338000000000000LL));
            int8_t* dst = a.dig + (((long)mt * a.nkb_out + jb) * S) * (long)A_SLICE + (row >> 3) * 512 + (c0 >> 4) * 128 +
                          (row & 7) * 16 + (c0 & 15);
            *reinterpret_cast<uint2*>(dst + 0L * A_SLICE) = pack_bytes8<5>(dv);
            *reinterpret_cast<uint2*>(dst + 1L * A_SLICE) = pack_bytes8<4>(dv);
            *reinterpret_cast<uint2*>(dst + 2L * A_SLICE) = pack_bytes8<3>(dv);
            *reinterpret_cast<uint2*>(dst + 3L * A_SLICE) = pack_bytes8<2>(dv);
            *reinterpret_cast<uint2*>(dst + 4L * A_SLICE) = pack_bytes8<1>(dv);
            *reinterpret_cast<uint2*>(dst + 5L * A_SLICE) = pack_bytes8<0>(dv);
          }
        }
        if (OUT == OUT_SUMSQ) a.sumsq[((long)jb * 2 + half) * m_pad + n] = ss;   // two partial sums per 64-column tile
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}
