// gpc_ozaki.cuh -- the L^-1 K* contraction on the 5th-generation tensor cores (tcgen05, INT8).
//
// FP64 on sm_100a has no tcgen05 kind: DMMA and DFMA share one FP64 pipe that peaks at 37 TFLOP/s
// (profiles/r01/pipes_r01.txt), so k_vt (gpc_predict.cuh) is within 10 % of what FP64 hardware can
// do.  This file evaluates the SAME contraction V = K* X^T exactly-enough on the INT8 tensor cores
// (4.56 POP/s measured, profiles/r01/umma_i8_rate_r01.txt) with the Ozaki splitting scheme:
//
//   x = scale * 2^-49 * v,  v = rint(x / scale * 2^49) = sum_{p<S} d_p 128^(S-1-p),  d_p in [-64, 63]
//   (balanced base-128 digits, S = 7 int8 "slices" per operand), so that
//   sum_k a_k b_k = sA sB 2^-98 sum_{p,q} 128^(12-p-q) sum_k d^A_pk d^B_qk.
// Every digit GEMM C_pq = D^A_p (D^B_q)^T is exact in int32 (|d| <= 64, K <= 65536); the pairs with
// p + q = t share one TMEM accumulator; pairs with p + q >= S (below 2^-49 of the operand scales)
// are dropped.  S (S + 1) / 2 = 28 digit GEMMs replace one FP64 GEMM.
//
// Operand images.  Digits are stored in global memory already in the shared-memory image the MMA
// consumes (K-major, no swizzle, 8 x 16-byte core matrices): a (rows x 64-byte) k-block of one
// slice is [rows/8][4][8][16 B] (SBO = 512 B, LBO = 128 B).  All S slices of a k-block are
// contiguous, so one stage of the pipeline is two 1-D bulk copies (TMA, cp.async.bulk).
//   A image  [m_pad/128][n_pad/64][S][8 KB]   test rows  x train index (digits of K* / sA)
//   B image  [n_pad/64 ][n_pad/64][S][4 KB]   rows of X x train index (digits of X_i,: / sB_i)
//
// Kernel k_vt_i8: persistent, one CTA per SM, 192 threads:
//   warp 0   TMA producer (one lane): 2-stage ring of (A k-block, B k-block) = 84 KB per stage
//   warp 1   MMA issuer (one lane): 56 tcgen05.mma.kind::i8 (M = 128, N = 64, K = 32) per stage into
//            S accumulators of 64 TMEM columns; tcgen05.commit frees the stage / publishes the tile
//   warps 2-5 epilogue: tcgen05.ld the S int32 levels, recombine in int64, scale to FP64 and either
//            reduce sum_i V(n, i)^2 per test row (one thread owns a row: no shuffles) or store V.
// The schedule is the one of k_vt: item (mt, p) = train tiles nb2-1-p then p of test tile mt.
#pragma once
#include "gpc_common.cuh"

namespace gpoz {

constexpr int S = 7;                 // int8 slices per operand
constexpr int BK = 64;               // k-block (bytes = k values)
constexpr int TM = 128, TN = 64;     // MMA tile: 128 test rows x 64 train rows
constexpr int A_SLICE = TM * BK;     // 8192 B
constexpr int B_SLICE = TN * BK;     // 4096 B
constexpr int A_STAGE = S * A_SLICE, B_STAGE = S * B_SLICE;
constexpr int STAGE_BYTES = A_STAGE + B_STAGE;   // 86016
constexpr int STAGES = 2;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + alignment slack
constexpr int NT = 192;
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

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

// Balanced base-128 digits of v (|v| < 2^47, i.e. |x| / scale < 1/4): d[S-1] least significant.
__device__ __forceinline__ void digits7(long long v, int (&d)[S]) {
#pragma unroll
  for (int p = S - 1; p >= 0; --p) {
    const int low = (int)(((v + 64) & 127) - 64);
    d[p] = low;
    v = (v - low) >> 7;
  }
}

}  // namespace gpoz

// ------------------------------------------------------------------------------------------
// Digits of the rows of X = L^-1 (lower triangular, row-major, ld).  Row i is scaled by
// sB[i] = 2^e with |X(i, :)| / sB[i] <= 1/2.  One warp per row; grid = n_pad / 8, 256 threads.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_slice_rows(const double* __restrict__ X, long ld, long n_pad,
                                                    int8_t* __restrict__ Bimg, double* __restrict__ sB) {
  using namespace gpoz;
  const long i = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n_pad) return;
  const double* row = X + i * ld;
  double mx = 0.0;
  for (long k = lane; k <= i; k += 32) mx = fmax(mx, fabs(row[k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  int e = 0;
  frexp(mx, &e);  // mx = f 2^e, f in [0.5, 1)  ->  |x| / 2^(e+2) < 1/4: the top balanced digit cannot overflow
  const double scale = (mx > 0.0) ? ldexp(1.0, e + 2) : 1.0;
  if (lane == 0) sB[i] = scale;
  const double mul = 562949953421312.0 / scale;  // 2^49 / scale
  const long jb = i >> 6, r = i & 63;
  const long nkb = n_pad >> 6;
  // every 16-byte chunk (16 consecutive k) of every slice; k-blocks beyond the diagonal stay zero
  for (long c = lane; c < (jb + 1) * 4; c += 32) {
    const long kb = c >> 2, cc = c & 3;
    int dg[16][S];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const long k = kb * 64 + cc * 16 + u;
      const double x = (k <= i) ? row[k] : 0.0;
      digits7(__double2ll_rn(x * mul), dg[u]);
    }
    int8_t* base = Bimg + ((jb * nkb + kb) * S) * (long)B_SLICE + (r >> 3) * 512 + cc * 128 + (r & 7) * 16;
#pragma unroll
    for (int p = 0; p < S; ++p) {
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        w[q] = (uint32_t)(dg[4 * q][p] & 255) | ((uint32_t)(dg[4 * q + 1][p] & 255) << 8) |
               ((uint32_t)(dg[4 * q + 2][p] & 255) << 16) | ((uint32_t)(dg[4 * q + 3][p] & 255) << 24);
      *reinterpret_cast<uint4*>(base + (long)p * B_SLICE) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// The contraction.  grid = min(#SMs, items), 192 threads, dynamic smem gpoz::SMEM_BYTES.
//   sumsq [nb2][m_pad] (SUMSQ) / Vt [m_pad][ld] (STORE_V), as k_vt.
// ------------------------------------------------------------------------------------------
template <bool STORE_V, bool SUMSQ>
__global__ void __launch_bounds__(gpoz::NT, 1) k_vt_i8(const int8_t* __restrict__ Aimg, const int8_t* __restrict__ Bimg,
                                                      const double* __restrict__ sB, double sA, long ld, int nb2,
                                                      long m_pad, int n_items, double* __restrict__ Vt,
                                                      double* __restrict__ sumsq) {
  using namespace gpoz;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar, tmem_empty_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int npair = (nb2 + 1) / 2;
  const long nkb_total = ld >> 6;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    mbar_init(&tmem_empty_bar, 4);  // one arrive per epilogue warp
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
        const int mt = item / npair, p = item - mt * npair;
        const int jbs[2] = {nb2 - 1 - p, p};
        const int nseg = (jbs[1] < jbs[0]) ? 2 : 1;
        for (int sg = 0; sg < nseg; ++sg) {
          const int jb = jbs[sg];
          for (int kb = 0; kb <= jb; ++kb, ++it) {
            const int st = it & 1;
            if (it >= STAGES) mbar_wait_guarded(&empty_bar[st], ((it >> 1) - 1) & 1);
            uint8_t* dst = smem + st * STAGE_BYTES;
            mbar_expect_tx(&full_bar[st], STAGE_BYTES);
            tma_bulk_g2s(dst, Aimg + (((long)mt * nkb_total + kb) * S) * (long)A_SLICE, A_STAGE, &full_bar[st]);
            tma_bulk_g2s(dst + A_STAGE, Bimg + (((long)jb * nkb_total + kb) * S) * (long)B_SLICE, B_STAGE, &full_bar[st]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // c = S32 (2) @4, a = b = INT8 (1) @7 / @10, K-major, N >> 3 @17, M >> 4 @24
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      uint32_t it = 0, tile = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int mt = item / npair, p = item - mt * npair;
        (void)mt;
        const int jbs[2] = {nb2 - 1 - p, p};
        const int nseg = (jbs[1] < jbs[0]) ? 2 : 1;
        for (int sg = 0; sg < nseg; ++sg, ++tile) {
          const int jb = jbs[sg];
          if (tile > 0) mbar_wait_guarded(&tmem_empty_bar, (tile - 1) & 1);  // epilogue drained the accumulators
          asm volatile("tcgen05.fence::after_thread_sync;");
          for (int kb = 0; kb <= jb; ++kb, ++it) {
            const int st = it & 1;
            mbar_wait_guarded(&full_bar[st], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t sa = smem_u32(smem + st * STAGE_BYTES), sb = sa + A_STAGE;
#pragma unroll
            for (int t = 0; t < S; ++t) {
#pragma unroll
              for (int pp = 0; pp <= t; ++pp) {
                const int qq = t - pp;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk)
                  umma_i8(tmem_base + t * TN, umma_desc(sa + pp * A_SLICE + kk * 256),
                          umma_desc(sb + qq * B_SLICE + kk * 256), idesc, (kb > 0 || pp > 0 || kk > 0) ? 1u : 0u);
              }
            }
            umma_commit(&empty_bar[st]);  // the stage may be refilled once these MMAs have read it
          }
          umma_commit(&tmem_full_bar);    // all MMAs of this tile done: accumulators complete
        }
      }
    }
  } else {
    // ===== epilogue warps (2..5): TMEM lane quarter = warp % 4 =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t tile = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int mt = item / npair, p = item - mt * npair;
      const int jbs[2] = {nb2 - 1 - p, p};
      const int nseg = (jbs[1] < jbs[0]) ? 2 : 1;
      for (int sg = 0; sg < nseg; ++sg, ++tile) {
        const int jb = jbs[sg];
        mbar_wait_guarded(&tmem_full_bar, tile & 1);
        asm volatile("tcgen05.fence::after_thread_sync;");
        double ss = 0.0;
        const long n = (long)mt * TM + row;
#pragma unroll 1
        for (int c0 = 0; c0 < TN; c0 += 16) {
          long long hi[16], lo[16];
          {
            int32_t v[16];
            tmem_ld16(lane_addr + 0 * TN + c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) hi[j] = v[j];
          }
#pragma unroll
          for (int t = 1; t < 4; ++t) {
            int32_t v[16];
            tmem_ld16(lane_addr + t * TN + c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) hi[j] = hi[j] * 128 + v[j];
          }
          {
            int32_t v[16];
            tmem_ld16(lane_addr + 4 * TN + c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) lo[j] = v[j];
          }
#pragma unroll
          for (int t = 5; t < S; ++t) {
            int32_t v[16];
            tmem_ld16(lane_addr + t * TN + c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) lo[j] = lo[j] * 128 + v[j];
          }
          // comb = sum_t C_t 128^(6-t) = hi 128^3 + lo ;  V = comb 2^-56 sA sB[i]
          const double* sb = sB + (long)jb * TN + c0;
          double vv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const double comb = fma((double)hi[j], 2097152.0, (double)lo[j]);
            vv[j] = comb * (COMB_SCALE * sA) * sb[j];
            ss = fma(vv[j], vv[j], ss);
          }
          if (STORE_V) {
            double* dst = Vt + n * ld + (long)jb * TN + c0;
#pragma unroll
            for (int j = 0; j < 16; j += 2) *reinterpret_cast<double2*>(dst + j) = make_double2(vv[j], vv[j + 1]);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar);
        if (SUMSQ) sumsq[(long)jb * m_pad + n] = ss;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}
