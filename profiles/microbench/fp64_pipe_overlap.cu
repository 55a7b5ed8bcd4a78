// Do DMMA and DFMA overlap?  Every SM sub-partition gets the same mix: warps 0..7 DMMA, 8..15 DFMA.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512) mix(double* out, int it_dmma, int it_dfma) {
  const int warp = threadIdx.x >> 5;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; i++) { c[i][0] = i; c[i][1] = 0; }
  if (warp >= 8) {
    for (int it = 0; it < it_dfma; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) { c[i][0] = fma(c[i][0], a, b); c[i][1] = fma(c[i][1], a, b); }
    }
  } else {
    for (int it = 0; it < it_dmma; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
float run(double* out, int sms, int a, int b) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  mix<<<sms, 512>>>(out, a, b); cudaDeviceSynchronize();
  cudaEventRecord(e0); mix<<<sms, 512>>>(out, a, b); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 512);
  const int I = 20000;
  float t_m = run(out, sms, I, 0);
  printf("DMMA alone (8 warps/SM): %.3f ms  %.2f TF\n", t_m, 512.0 * 16 * I * 8 * sms / t_m * 1e-9);
  for (int k = 1; k <= 8; k *= 2) {
    int J = I * k;
    float t_f = run(out, sms, 0, J), t_b = run(out, sms, I, J);
    printf("DFMA iters x%d: DFMA alone %.3f ms (%.2f TF)  both %.3f ms  [sum %.3f, max %.3f]  combined %.2f TF\n", k, t_f,
           2.0 * 32 * J * 32.0 * 8 * sms / t_f * 1e-9, t_b, t_m + t_f, t_m > t_f ? t_m : t_f,
           (512.0 * 16 * I * 8 * sms + 2.0 * 32 * J * 32.0 * 8 * sms) / t_b * 1e-9);
  }
  return 0;
}
