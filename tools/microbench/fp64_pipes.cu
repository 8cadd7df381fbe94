// Microbenchmarks (development tool): FP64 FMA throughput / latency, rsqrt latency, DMMA throughput.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_tput(double* out, int iters) {
  double a[8]; for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dfma_lat(double* out, long long* cyc, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) a = fma(a, b, c);
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void rsqrt_lat(double* out, long long* cyc, int iters) {
  double a = 1.5 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) a = rsqrt(a) + 1.25;
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void exp_lat(double* out, long long* cyc, int iters) {
  double a = -0.5 - threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) a = exp(a) - 1.3;
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void exp_tput(double* out, int iters) {
  double a[4]; for (int i = 0; i < 4; ++i) a[i] = -0.5 - threadIdx.x * 1e-3 - i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = exp(a[i]) - 1.3;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a[0] + a[1] + a[2] + a[3];
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void dmma_tput(double* out, int iters) {
  double c[16][2]; for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0; for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma_lat(double* out, long long* cyc, int iters) {
  double c0 = 0, c1 = 0, a = threadIdx.x * 1e-3, b = 1.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) dmma(c0, c1, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = c0 + c1; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8 * 8); cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms; long long h;
  int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    dfma_tput<<<148 * 2, threads>>>(out, 100);
    cudaEventRecord(e0); dfma_tput<<<148 * 2, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DFMA tput  %4d thr x2 CTA/SM: %.2f TFLOP/s\n", threads, 2.0 * 148 * 2 * threads * 8.0 * iters / ms / 1e9);
  }
  for (int threads : {128, 256, 512}) {
    dmma_tput<<<148, threads>>>(out, 100);
    cudaEventRecord(e0); dmma_tput<<<148, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DMMA tput  %4d thr x1 CTA/SM: %.2f TFLOP/s\n", threads, 2.0 * 148 * (threads / 32) * 16.0 * 256 * iters / ms / 1e9);
  }
  for (int threads : {256, 1024}) {
    exp_tput<<<148 * 2, threads>>>(out, 10);
    cudaEventRecord(e0); exp_tput<<<148 * 2, threads>>>(out, 2000); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("exp  tput  %4d thr x2 CTA/SM: %.2f Gexp/s\n", threads, 148.0 * 2 * threads * 4.0 * 2000 / ms / 1e6);
  }
  dfma_lat<<<1, 32>>>(out, cyc, 10000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("DFMA latency  %.1f cyc\n", h / 10000.0);
  dmma_lat<<<1, 32>>>(out, cyc, 10000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("DMMA latency  %.1f cyc\n", h / 10000.0);
  rsqrt_lat<<<1, 32>>>(out, cyc, 10000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("rsqrt+add latency %.1f cyc\n", h / 10000.0);
  exp_lat<<<1, 32>>>(out, cyc, 10000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("exp+add latency %.1f cyc\n", h / 10000.0);
  return 0;
}
