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

// k((xa, fa), (xb, fb)): stationary for F == 1, Kennedy-O'Hagan AR1 sum otherwise.
__device__ __forceinline__ double gpc_kval(const GpcHyp& h, double ax, double ay, double az, double af,
                                           double bx, double by, double bz, double bf) {
  const double dx = ax - bx, dy = ay - by, dz = az - bz;
  if (h.F == 1) return gpc_base_k(h, 0, dx, dy, dz);
  const int fi = (int)af, fj = (int)bf;
  const int mm = fi < fj ? fi : fj;
  double s = 0.0;
  for (int m = 0; m <= mm; ++m) s = fma(h.coef[fi][m] * h.coef[fj][m], gpc_base_k(h, m, dx, dy, dz), s);
  return s;
}

__device__ __forceinline__ int gpc_fid(const GpcHyp& h, double f) {
  int i = (int)f;
  return (h.F == 1 || i < 0) ? 0 : (i >= h.F ? h.F - 1 : i);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
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
