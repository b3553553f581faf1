// fp64_peak.cu -- what the B200 fp64 pipe can do, measured on the box (SURVEY.md 8d asks for it):
//   (1) DFMA throughput with 8 independent chains per thread, full occupancy   -> TFLOP/s
//   (2) latency of ONE dependent DFMA chain per warp (cycles per DFMA), 1 and 14 warps per SM
//   (3) the same for a dependent chain of fp64 divisions and of psd-style exp() calls is left to
//       tools/prof_timing.py (it times the product's own routines).
// The DP kernel is a chain of dependent fp64 operations, so (2) is its real roofline.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_throughput(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void dfma_latency(double* out, long long* cycles, int iters, double a, double b) {
  double x = threadIdx.x;
  const long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; i++) x = fma(x, a, b);
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void ddiv_latency(double* out, long long* cycles, int iters, double a) {
  double x = 1.0 + threadIdx.x;
  const long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < iters; i++) x = a / x + 1.0;
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * sms * 2048); cudaMalloc(&cyc, sizeof(long long) * sms * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1 << 16;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    dfma_throughput<<<sms * 2, 1024>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8 * (double)iters * sms * 2 * 1024;
    if (rep == 2) printf("{\"dfma_tflops\": %.2f, \"sms\": %d", flops / (ms * 1e-3) / 1e12, sms);
  }
  long long h[8];
  for (int warps = 1; warps <= 14; warps += 13) {
    dfma_latency<<<sms, 32 * warps>>>(out, cyc, 1 << 14, 1.0000001, 1e-9);
    dfma_latency<<<sms, 32 * warps>>>(out, cyc, 1 << 14, 1.0000001, 1e-9);
    cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
    printf(", \"dfma_dependent_cycles_%dwarps\": %.2f", warps, (double)h[0] / (1 << 14));
    ddiv_latency<<<sms, 32 * warps>>>(out, cyc, 1 << 12, 3.0);
    ddiv_latency<<<sms, 32 * warps>>>(out, cyc, 1 << 12, 3.0);
    cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
    printf(", \"ddiv_plus_dadd_dependent_cycles_%dwarps\": %.2f", warps, (double)h[0] / (1 << 12));
  }
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf(", \"sm_clock_mhz_nominal\": %d}\n", khz / 1000);
  return cudaGetLastError() != cudaSuccess;
}
