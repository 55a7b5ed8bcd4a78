// tma_multicast_bw.cu -- does cp.async.bulk multicast lift the L2 -> SM delivery cap?
// Each CTA consumes, per k-step, one 48 KB "A" block that the whole cluster shares plus a 24 KB "B"
// block of its own (the operand pattern of k_vt_i8).  unicast: every CTA fetches all 72 KB itself;
// multicast: every CTA fetches 48/csz KB of A and multicasts it to the cluster (+ its own B).
// No math: the consumer waits for the stage and releases it (to every CTA of the cluster).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_multicast_bw tma_multicast_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

constexpr int A_BYTES = 49152, B_BYTES = 24576, STAGE = A_BYTES + B_BYTES, STAGES = 3;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(bar)), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(s32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}

template <int CSZ, bool MC>
__global__ void __launch_bounds__(64, 1) k(const uint8_t* A, const uint8_t* B, long a_blocks, long b_blocks, int ksteps, unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
  uint32_t rank = 0;
  if (CSZ > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int cluster = blockIdx.x / CSZ;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], MC ? CSZ : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (CSZ > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); } else __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int it = 0; it < ksteps; ++it) {
      const int st = it % STAGES;
      if (it >= STAGES) mbar_wait(&empty_bar[st], ((it / STAGES) - 1) & 1);
      uint8_t* dst = smem + st * STAGE;
      expect_tx(&full_bar[st], STAGE);
      const long ab = ((long)cluster * 977 + it) % a_blocks;       // the cluster walks A blocks
      const long bb = ((long)blockIdx.x * 131 + it) % b_blocks;    // every CTA its own B blocks
      if (MC) {
        const uint32_t part = A_BYTES / CSZ;
        bulk_mc(dst + rank * part, A + ab * A_BYTES + rank * part, part, &full_bar[st], (uint16_t)((1u << CSZ) - 1));
      } else {
        bulk(dst, A + ab * A_BYTES, A_BYTES, &full_bar[st]);
      }
      bulk(dst + A_BYTES, B + bb * B_BYTES, B_BYTES, &full_bar[st]);
    }
  } else if (warp == 1 && lane == 0) {
    unsigned acc = 0;
    for (int it = 0; it < ksteps; ++it) {
      const int st = it % STAGES;
      mbar_wait(&full_bar[st], (it / STAGES) & 1);
      acc += smem[st * STAGE + (it & 1023)];
      if (MC) { for (uint32_t r = 0; r < CSZ; ++r) arrive_remote(&empty_bar[st], r); }
      else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty_bar[st])) : "memory");
    }
    if (acc == 0xffffffffu) *sink = acc;
  }
  if (CSZ > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); } else __syncthreads();
}

template <int CSZ, bool MC>
void run(const uint8_t* A, const uint8_t* B, long ab, long bb, unsigned* sink, int nsm) {
  const int ksteps = 2000;
  const int grid = (nsm / CSZ) * CSZ;
  auto kern = k<CSZ, MC>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * STAGE);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = STAGES * STAGE;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CSZ; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, A, B, ab, bb, ksteps, sink);
    cudaEventRecord(e1);
    if (e != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) { printf("csz=%d mc=%d: %s\n", CSZ, (int)MC, cudaGetErrorString(cudaGetLastError())); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
  }
  const double delivered = (double)grid * ksteps * STAGE;
  const double fetched = (double)grid * ksteps * (B_BYTES + (MC ? A_BYTES / CSZ : A_BYTES));
  printf("cluster=%d %s grid=%d: %.3f ms  delivered to smem %.2f TB/s  fetched from L2 %.2f TB/s  (%.3f us per 72 KB k-step)\n", CSZ,
         MC ? "multicast" : "unicast  ", grid, best, delivered / best / 1e9, fetched / best / 1e9, 1e3 * best / ksteps);
}

int main() {
  int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const long ab = 1200, bb = 800;   // 59 MB + 20 MB: L2-resident
  uint8_t *A, *B; unsigned* sink;
  cudaMalloc(&A, ab * A_BYTES); cudaMalloc(&B, bb * B_BYTES); cudaMalloc(&sink, 4);
  cudaMemset(A, 1, ab * A_BYTES); cudaMemset(B, 1, bb * B_BYTES);
  run<1, false>(A, B, ab, bb, sink, nsm);
  run<2, false>(A, B, ab, bb, sink, nsm);
  run<2, true>(A, B, ab, bb, sink, nsm);
  run<4, false>(A, B, ab, bb, sink, nsm);
  run<4, true>(A, B, ab, bb, sink, nsm);
  run<8, true>(A, B, ab, bb, sink, nsm);
  return 0;
}
