// gpc_common.cuh -- shared device-side definitions of the GP core (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GPC_TILE 128      // row/column block of every dense operand (N is padded to it)
#define GPC_MAXF 4        // fidelities supported by the AR1 kernel
#define GPC_MAXK 64       // max points per IG candidate

// Covariance hyper-parameters in evaluation-ready form (built on the host by gpc_set_hypers).
struct GpcHyp {
  int base;                      // 0 = squared exponential, 1 = Matern-3/2
  int F;                         // number of fidelities (1: single fidelity / NIGP)
  double var[GPC_MAXF];          // k_m variance
  double inv_l[GPC_MAXF][3];     // 1 / lengthscale
  double coef[GPC_MAXF][GPC_MAXF];  // coef[i][m] = prod_{l=m}^{i-1} rho_l  (0 for m > i)
  double noise[GPC_MAXF];        // likelihood variance per fidelity
  double kdiag[GPC_MAXF];        // prior variance at fidelity i
  double jitter;                 // added to the training diagonal
};

__device__ __forceinline__ double gpc_base_k(const GpcHyp& h, int m, double dx, double dy, double dz) {
  const double sx = dx * h.inv_l[m][0], sy = dy * h.inv_l[m][1], sz = dz * h.inv_l[m][2];
  const double r2 = fma(sx, sx, fma(sy, sy, sz * sz));
  if (h.base == 0) return h.var[m] * exp(-0.5 * r2);
  const double r = 1.7320508075688772 * sqrt(r2);
  return h.var[m] * (1.0 + r) * exp(-r);
}

__device__ __forceinline__ int gpc_fid(const GpcHyp& h, double f) {
  int i = (int)f;
  return (h.F == 1 || i < 0) ? 0 : (i >= h.F ? h.F - 1 : i);
}

// k((xa, fa), (xb, fb)): stationary for F == 1, Kennedy-O'Hagan AR1 sum otherwise.
__device__ __forceinline__ double gpc_kval(const GpcHyp& h, double ax, double ay, double az, double af,
                                           double bx, double by, double bz, double bf) {
  const double dx = ax - bx, dy = ay - by, dz = az - bz;
  if (h.F == 1) return gpc_base_k(h, 0, dx, dy, dz);
  const int fi = gpc_fid(h, af), fj = gpc_fid(h, bf);   // the C ABI rejects out-of-range labels; never index past coef
  const int mm = fi < fj ? fi : fj;
  double s = 0.0;
  for (int m = 0; m <= mm; ++m) s = fma(h.coef[fi][m] * h.coef[fj][m], gpc_base_k(h, m, dx, dy, dz), s);
  return s;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- exp(-q), q >= 0 ----------------------------------------------------------------------------
// exp(-q) for q >= 0, branch-free so that the eight columns a thread works on interleave in the FP64
// pipe (the library exp() carries special-case branches that serialise them): k = rint(-q log2 e) by the
// magic-number add, two-step Cody-Waite reduction to |r| <= ln2 / 2, degree-13 Taylor polynomial
// (truncation 1.7e-16 relative at the interval edge), 2^k by an exponent-field add.  q is clamped at 700
// (e^-700 = 1e-304 is zero against any digit or mean), so k >= -1010 and the result stays normal.
__device__ __forceinline__ double gpc_exp_neg(double q) {
  q = fmin(q, 700.0);
  const double t = fma(q, -1.4426950408889634074, 6755399441055744.0);
  const int k = __double2loint(t);
  const double kf = t - 6755399441055744.0;
  double r = fma(kf, -6.93147180369123816490e-01, -q);
  r = fma(kf, -1.90821492927058770002e-10, r);
  double p = 1.6059043836821613e-10;            // 1/13!
  p = fma(p, r, 2.08767569878681e-09);          // 1/12!
  p = fma(p, r, 2.505210838544172e-08);         // 1/11!
  p = fma(p, r, 2.755731922398589e-07);         // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);        // 1/9!
  p = fma(p, r, 2.48015873015873e-05);          // 1/8!
  p = fma(p, r, 1.984126984126984e-04);         // 1/7!
  p = fma(p, r, 1.3888888888888889e-03);        // 1/6!
  p = fma(p, r, 8.333333333333333e-03);         // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);        // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);        // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// fma that the compiler may not reorder: the stages of the W-wide exp below are volatile asm so that the W dependency
// chains advance breadth-first (W-way ILP in the FP64 pipe from a single warp; the library exp() carries special-case
// branches that keep the columns of a thread from interleaving).
__device__ __forceinline__ double fma_pinned(double a, double b, double c) {
  double d;
  asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c));
  return d;
}

// exp(-q[u]) for W independent arguments, table-driven: k = rint(-64 q / ln 2) = 64 e + j, exp(-q) = 2^e T[j] exp(r) with T[j] = 2^(j/64)
// (64 doubles in shared memory) and |r| <= ln2 / 128, where a degree-5 polynomial is exact to 3.5e-17.  Eleven FP64-pipe
// operations per value instead of eighteen -- under the board's power cap the assembly kernel's time follows its energy,
// and the exp was half of its FP64 work.  Result within 2 ulp (the digits that are made from it carry 2^-49 of the scale).
template <int W>
__device__ __forceinline__ void gpc_exp_neg_tab_w(const double* __restrict__ qin, double* __restrict__ e,
                                                  const double* __restrict__ T64) {
  double q[W], t[W], r[W], p[W];
#pragma unroll
  for (int u = 0; u < W; ++u) q[u] = fmin(qin[u], 700.0);
#pragma unroll
  for (int u = 0; u < W; ++u) t[u] = fma_pinned(q[u], -92.332482616893653, 6755399441055744.0);   // -64 / ln 2
#pragma unroll
  for (int u = 0; u < W; ++u) r[u] = t[u] - 6755399441055744.0;       // kf
#pragma unroll
  for (int u = 0; u < W; ++u) q[u] = fma_pinned(r[u], -6.93147180369123816490e-01 / 64.0, -q[u]);   // ln2 / 64, high part (exact product)
#pragma unroll
  for (int u = 0; u < W; ++u) r[u] = fma_pinned(r[u], -1.90821492927058770002e-10 / 64.0, q[u]);    // low part
#pragma unroll
  for (int u = 0; u < W; ++u) p[u] = fma_pinned(8.333333333333333e-03, r[u], 4.1666666666666664e-02);   // 1/5!, 1/4!
#define GPC_EXP_STAGE(c)            \
  _Pragma("unroll") for (int u = 0; u < W; ++u) p[u] = fma_pinned(p[u], r[u], c);
  GPC_EXP_STAGE(1.6666666666666666e-01)   // 1/3!
  GPC_EXP_STAGE(0.5)
  GPC_EXP_STAGE(1.0)
  GPC_EXP_STAGE(1.0)
#undef GPC_EXP_STAGE
#pragma unroll
  for (int u = 0; u < W; ++u) {
    const int k = __double2loint(t[u]);
    const double v = p[u] * T64[k & 63];
    e[u] = __hiloint2double(__double2hiint(v) + ((k >> 6) << 20), __double2loint(v));
  }
}


// ---- mbarrier / bulk-copy (TMA) helpers ----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk async copy global -> shared through the TMA unit (SASS: UBLKCP), completion on mbarrier.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- cp.async (LDGSTS) helpers -----------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// FP64 tensor-core MMA, the only native shape on sm_100a (SASS: DMMA.8x8x4).
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
