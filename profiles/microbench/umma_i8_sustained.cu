// Sustained vs burst issue rate of tcgen05.mma kind::i8 (M = 128, N = 256, K = 32, operands in shared memory, random
// digit bytes so that the data path toggles like real operands): one short launch (burst, ~2 ms) and back-to-back
// launches for `seconds` (sustained: the power cap, not the issue rate, sets the clock).  Prints one JSON line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_sustained umma_i8_sustained.cu && ./umma_i8_sustained 4
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <chrono>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
constexpr int N = 256, NACC = 2;
__global__ void __launch_bounds__(128) rate(int iters, int32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t x = 0x9E3779B9u * (blockIdx.x * 128 + tid + 1);
  for (int i = tid; i < (16384 + N * 128) * 2 / 4; i += 128) { x = x * 1664525u + 1013904223u; reinterpret_cast<uint32_t*>(smem)[i] = x; }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (tid == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t da = make_desc(smem_u32(smem) + ((it & 1) * (16384 + N * 128)) + kk * 256, 128, 1024);
          const uint64_t db = make_desc(smem_u32(smem) + ((it & 1) * (16384 + N * 128)) + 16384 + kk * 256, 128, 1024);
          // accumulate = 0 on the first pass keeps the int32 accumulators from saturating
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_base + a * N), "l"(da), "l"(db), "r"(idesc), "r"((it & 63) ? 1u : 0u) : "memory");
        }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(&bar)), "r"(0u) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid == 0) out[blockIdx.x] = 1;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}
int main(int argc, char** argv) {
  const double seconds = argc > 1 ? atof(argv[1]) : 4.0;
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  int32_t* out; cudaMalloc(&out, 4 * sms);
  const int smem = (16384 + N * 128) * 2 + 1024, iters = 4000;
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double ops = 2.0 * 128 * N * 32 * 4.0 * NACC * iters * sms;
  rate<<<sms, 128, smem>>>(iters, out); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {   // burst: single launches with idle gaps in between
    cudaEventRecord(e0); rate<<<sms, 128, smem>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  const double burst = ops / best * 1e-9;   // TOP/s
  // sustained: back-to-back launches for `seconds`, rate over the LAST half of the window
  int launches = 0, half_at = 0; auto t0 = std::chrono::steady_clock::now();
  cudaEvent_t h0; cudaEventCreate(&h0);
  bool half_marked = false;
  while (true) {
    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (el > seconds) break;
    if (!half_marked && el > seconds / 2) { cudaEventRecord(h0); half_marked = true; half_at = launches; }
    for (int i = 0; i < 20; ++i) rate<<<sms, 128, smem>>>(iters, out);
    launches += 20;
    cudaStreamSynchronize(0);
  }
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, h0, e1);
  const double sustained = ops * (launches - half_at) / ms * 1e-9;
  printf("{\"kernel\": \"tcgen05.mma.cta_group::1.kind::i8 M=128 N=256 K=32, smem operands, random bytes, %d CTAs\", "
         "\"burst_tops\": %.1f, \"sustained_tops\": %.1f, \"sustained_window_s\": %.2f, \"error\": \"%s\"}\n",
         sms, burst, sustained, ms * 1e-3, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
