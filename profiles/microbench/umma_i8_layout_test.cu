// Validate tcgen05.mma kind::i8 (s8 x s8 -> s32) with SWIZZLE_NONE K-major smem descriptors.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
template <int N>
__global__ void __launch_bounds__(128) umma_test(const int8_t* __restrict__ Aimg, const int8_t* __restrict__ Bimg, int kblocks, int32_t* __restrict__ C) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                 // 128 x 128 B
  uint8_t* sB = smem + 16384;         // N x 128 B
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_s;
  // instruction descriptor: c=S32(2)@4, a=INT8(1)@7, b=INT8(1)@10, K-major both, n_dim=N>>3 @17, m_dim=128>>4 @24
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  for (int kb = 0; kb < kblocks; ++kb) {
    if (tid == 0) {
      mbar_expect_tx(&bar_load, 16384 + N * 128);
      tma_bulk_g2s(sA, Aimg + (size_t)kb * 16384, 16384, &bar_load);
      tma_bulk_g2s(sB, Bimg + (size_t)kb * N * 128, N * 128, &bar_load);
    }
    mbar_wait(&bar_load, kb & 1);
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (tid == 0) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t da = make_desc(smem_u32(sA) + kk * 256, 128, 1024);
        const uint64_t db = make_desc(smem_u32(sB) + kk * 256, 128, 1024);
        const uint32_t acc = (kb > 0 || kk > 0) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    mbar_wait(&bar_mma, kb & 1);   // smem reusable, accumulator updated
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  // epilogue: warp w reads TMEM lanes 32w..32w+31, 32 columns at a time
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) C[(size_t)row * N + c0 + j] = (int32_t)v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)N));
}
// host: tile image, layout [rb][c][8 rows][16 B] per k-block of 128 bytes
static void make_image(const std::vector<int8_t>& M, int rows, int K, std::vector<int8_t>& img) {
  img.assign((size_t)rows * K, 0);
  const int kblocks = K / 128;
  for (int kb = 0; kb < kblocks; ++kb)
    for (int r = 0; r < rows; ++r)
      for (int k = 0; k < 128; ++k) {
        size_t off = (size_t)kb * rows * 128 + (size_t)(r / 8) * 1024 + (size_t)(k / 16) * 128 + (size_t)(r % 8) * 16 + (k % 16);
        img[off] = M[(size_t)r * K + kb * 128 + k];
      }
}
template <int N> int run(int K) {
  std::vector<int8_t> A((size_t)128 * K), B((size_t)N * K), Ai, Bi;
  srand(1);
  for (auto& v : A) v = (int8_t)(rand() % 255 - 127);
  for (auto& v : B) v = (int8_t)(rand() % 255 - 127);
  make_image(A, 128, K, Ai); make_image(B, N, K, Bi);
  int8_t *dA, *dB; int32_t* dC;
  cudaMalloc(&dA, Ai.size()); cudaMalloc(&dB, Bi.size()); cudaMalloc(&dC, (size_t)128 * N * 4);
  cudaMemcpy(dA, Ai.data(), Ai.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, Bi.data(), Bi.size(), cudaMemcpyHostToDevice);
  cudaMemset(dC, 0xff, (size_t)128 * N * 4);
  const int smem = 16384 + N * 128 + 1024;
  cudaFuncSetAttribute(umma_test<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_test<N><<<1, 128, smem>>>(dA, dB, K / 128, dC);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d K=%d: CUDA error %s\n", N, K, cudaGetErrorString(e)); return 1; }
  std::vector<int32_t> C((size_t)128 * N);
  cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
  long bad = 0; int first = -1;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < N; ++j) {
      long s = 0;
      for (int k = 0; k < K; ++k) s += (long)A[(size_t)i * K + k] * B[(size_t)j * K + k];
      if (s != C[(size_t)i * N + j]) { if (first < 0) first = i * N + j; ++bad; }
    }
  printf("N=%d K=%d: mismatches %ld of %d", N, K, bad, 128 * N);
  if (bad) printf("  first at (%d,%d): got %d", first / N, first % N, C[first]);
  printf("\n");
  return bad != 0;
}
int main() {
  int rc = 0;
  rc |= run<128>(128);
  rc |= run<128>(512);
  rc |= run<256>(256);
  rc |= run<64>(256);
  return rc;
}
