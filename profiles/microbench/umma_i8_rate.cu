// Issue-rate of tcgen05.mma kind::i8 from fixed smem operands: M=128, N in {64,128,256}, K=32 per instruction.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
template <int N, int NACC>
__global__ void __launch_bounds__(128) rate(int iters, int32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + N * 128) * 2 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_base + a * N), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
        }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid == 0) out[blockIdx.x] = 1;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}
template <int N, int NACC> void run(int sms, int32_t* out) {
  const int smem = (16384 + N * 128) * 2 + 1024, iters = 4000;
  cudaFuncSetAttribute(rate<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate<N, NACC><<<sms, 128, smem>>>(iters, out); cudaDeviceSynchronize();
  cudaEventRecord(e0); rate<N, NACC><<<sms, 128, smem>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError();
  double ops = 2.0 * 128 * N * 32 * 4.0 * NACC * iters * sms;
  printf("M=128 N=%3d accs=%d : %.3f ms  %.2f POPS  (%s)\n", N, NACC, ms, ops / ms * 1e-12, cudaGetErrorString(e));
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int32_t* out; cudaMalloc(&out, 4 * p.multiProcessorCount);
  run<64, 7>(p.multiProcessorCount, out);
  run<64, 1>(p.multiProcessorCount, out);
  run<128, 4>(p.multiProcessorCount, out);
  run<128, 1>(p.multiProcessorCount, out);
  run<256, 2>(p.multiProcessorCount, out);
  run<256, 1>(p.multiProcessorCount, out);
  return 0;
}
