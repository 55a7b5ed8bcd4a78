// Micro-benchmark: FP64 DFMA vs DMMA.8x8x4 issue-rate peaks on B200 (sm_100a).
// Establishes the FP64 roofline denominator that MEASURED_PEAKS.json lacks.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) exp_loop(double* out, int iters) {
  double x = -1e-3 * (threadIdx.x + 1), s = 0;
  for (int it = 0; it < iters; it++) { s += exp(x); x -= 1e-3; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("device %s SMs %d\n", p.name, sms);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  const int iters = 20000;
  for (int bps = 1; bps <= 4; bps *= 2) {
    int grid = sms * bps;
    float ms = time_ms([&] { dmma_loop<16><<<grid, 256>>>(out, iters); });
    double fl = 2.0 * 8 * 8 * 4 * 16.0 * iters * 8 * grid;
    printf("DMMA.8x8x4 NACC=16 ctas/SM=%d : %.2f TFLOP/s (%.3f ms)\n", bps, fl / ms * 1e-9, ms);
    ms = time_ms([&] { dmma_loop<4><<<grid, 256>>>(out, iters); });
    fl = 2.0 * 8 * 8 * 4 * 4.0 * iters * 8 * grid;
    printf("DMMA.8x8x4 NACC=4  ctas/SM=%d : %.2f TFLOP/s (%.3f ms)\n", bps, fl / ms * 1e-9, ms);
    ms = time_ms([&] { dfma_loop<16><<<grid, 256>>>(out, iters); });
    fl = 2.0 * 16.0 * iters * 256 * grid;
    printf("DFMA       NACC=16 ctas/SM=%d : %.2f TFLOP/s (%.3f ms)\n", bps, fl / ms * 1e-9, ms);
  }
  {
    int grid = sms * 4;
    float ms = time_ms([&] { exp_loop<<<grid, 256>>>(out, 4000); });
    printf("exp(double): %.1f Gexp/s (%.3f ms)\n", 4000.0 * 256 * grid / ms * 1e-6, ms);
  }
  return 0;
}
