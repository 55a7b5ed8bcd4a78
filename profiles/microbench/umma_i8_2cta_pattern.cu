// What a cta_group::2 operand pipeline could buy k_vt_i8: the MMA issue pattern of one k-step (64 k values, 21 digit
// pairs, 128 test rows per CTA x 64 rows of L^-1) run from STATIC shared-memory operands (random digit bytes, no TMA),
// persistent, sustained for seconds under the board's power cap:
//   1cta  the product kernel's pattern: per 32-k half-step 8 MMAs of M = 128, N = 256,128,256,64,256,192,128,64
//         (A slice read 8x, B slices stacked along N)
//   2cta  a CTA pair (cluster of 2): per half-step 9 MMAs of M = 256 issued by the leader; each CTA supplies its own
//         128 A rows and HALF of every B operand (N / 2 rows), so the B bytes read from shared memory per CTA halve
//         (layout R1..R6 of profiles/r02/README.md: N = 256,128 | 256,64 | 256 | 128,64 | 128 | 64)
// Both do 21 x 128 x 64 x 64 multiply-adds per CTA per k-step.  Prints one JSON line: burst and sustained TOP/s of each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_2cta_pattern umma_i8_2cta_pattern.cu
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // K-major, no swizzle: LBO 128 B, SBO 512 B
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46);
}
constexpr int A_SLICE = 128 * 64, B_SLICE = 64 * 64;   // bytes per k-block of one digit slice
constexpr int SMEM = 6 * A_SLICE + 6 * B_SLICE + 2048;

template <int CTAS>
__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (CTAS == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da),
                 "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da),
                 "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // a broken protocol must abort the kernel, never hang the GPU
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    if (spin > (1u << 26)) __trap();
  }
}

template <int CTAS>
__global__ void __launch_bounds__(128, 1) pattern(int ksteps, int32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t rank = 0;
  if (CTAS == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  uint32_t x = 0x9E3779B9u * (blockIdx.x * 128 + tid + 1);
  for (int i = tid; i < (6 * A_SLICE + 6 * B_SLICE) / 4; i += 128) {
    x = x * 1664525u + 1013904223u;
    reinterpret_cast<uint32_t*>(smem)[i] = x;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CTAS == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (CTAS == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;");
    asm volatile("barrier.cluster.wait.acquire.aligned;");
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t sa = smem_u32(smem), sb = sa + 6 * A_SLICE;
  if (tid == 0 && rank == 0) {
    const uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((CTAS * 128) >> 4) << 24);
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint32_t acc0 = (ks & 63) ? 1u : 0u;   // restart the int32 accumulators now and then
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        if (CTAS == 1) {
          // (pp, q0, nsl): D region pp + q0, A slice pp, B slices q0 .. q0 + nsl - 1 stacked along N
          const int pat[8][3] = {{0, 0, 4}, {0, 4, 2}, {1, 0, 4}, {1, 4, 1}, {2, 0, 4}, {3, 0, 3}, {4, 0, 2}, {5, 0, 1}};
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            const int pp = pat[m][0], q0 = pat[m][1], nsl = pat[m][2];
            const uint32_t idesc = idesc0 | ((uint32_t)((nsl * 64) >> 3) << 17);
            mma_i8<1>(tmem_base + (pp + q0) * 64, make_desc(sa + pp * A_SLICE + kk * 256), make_desc(sb + q0 * B_SLICE + kk * 256), idesc,
                      (pp > 0 || acc0) ? 1u : 0u);
          }
        } else {
          // (pp, t0, N, B region offset in 4 KB units): each CTA holds N / 2 rows of the operand at that offset
          // R1 @0 (2 slices), R2 @2, R3 @3 (half slice), R4 @4, R5 @5 (half), R6 @5.5 (half)
          const int pat[9][4] = {{0, 0, 256, 0}, {0, 4, 128, 8}, {1, 0, 256, 0}, {1, 4, 64, 12}, {2, 0, 256, 0},
                                 {3, 0, 128, 16}, {3, 2, 64, 20}, {4, 0, 128, 16}, {5, 0, 64, 22}};
#pragma unroll
          for (int m = 0; m < 9; ++m) {
            const int pp = pat[m][0], t0 = pat[m][1], N = pat[m][2], off = pat[m][3];
            const uint32_t idesc = idesc0 | ((uint32_t)(N >> 3) << 17);
            mma_i8<2>(tmem_base + (pp + t0) * 64, make_desc(sa + pp * A_SLICE + kk * 256), make_desc(sb + off * 1024 + kk * 256), idesc,
                      (pp > 0 || acc0) ? 1u : 0u);
          }
        }
      }
    }
    if (CTAS == 1)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                   "h"((uint16_t)3)
                   : "memory");
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid == 0) out[blockIdx.x] = 1;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (CTAS == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;");
    asm volatile("barrier.cluster.wait.acquire.aligned;");
  }
  if (warp == 0) {
    if (CTAS == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

template <int CTAS>
void launch(int sms, int ksteps, int32_t* out) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sms);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = SMEM;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CTAS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, pattern<CTAS>, ksteps, out);
}

template <int CTAS>
void run(int sms, int32_t* out, double seconds, double* burst, double* sustained) {
  const int ksteps = 2000;
  cudaFuncSetAttribute(pattern<CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  cudaEvent_t e0, e1, h0;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&h0);
  const double ops = 2.0 * 21.0 * 128 * 64 * 64 * (double)ksteps * sms;
  launch<CTAS>(sms, ksteps, out);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); launch<CTAS>(sms, ksteps, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  *burst = ops / best * 1e-9;
  int launches = 0, half_at = 0;
  bool half = false;
  auto t0 = std::chrono::steady_clock::now();
  while (true) {
    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (el > seconds) break;
    if (!half && el > seconds / 2) { cudaEventRecord(h0); half = true; half_at = launches; }
    for (int i = 0; i < 20; ++i) launch<CTAS>(sms, ksteps, out);
    launches += 20;
    cudaStreamSynchronize(0);
  }
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, h0, e1);
  *sustained = ops * (launches - half_at) / ms * 1e-9;
}

int main(int argc, char** argv) {
  const double seconds = argc > 1 ? atof(argv[1]) : 4.0;
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount & ~1;
  int32_t* out; cudaMalloc(&out, 4 * sms);
  double b1 = 0, s1 = 0, b2 = 0, s2 = 0;
  run<1>(sms, out, seconds, &b1, &s1);
  cudaError_t e1 = cudaGetLastError();
  run<2>(sms, out, seconds, &b2, &s2);
  cudaError_t e2 = cudaDeviceSynchronize();
  printf("{\"pattern\": \"one k-step of k_vt_i8 (21 digit pairs, 128 x 64 tile per CTA), static smem operands, %d CTAs\", "
         "\"cta_group1\": {\"burst_tops\": %.1f, \"sustained_tops\": %.1f, \"error\": \"%s\"}, "
         "\"cta_group2\": {\"burst_tops\": %.1f, \"sustained_tops\": %.1f, \"error\": \"%s\"}}\n",
         sms, b1, s1, cudaGetErrorString(e1), b2, s2, cudaGetErrorString(e2));
  return 0;
}
