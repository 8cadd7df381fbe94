// Cycle-level micro-benchmark of the 64x64 finalisers used on the Cholesky chain (development tool).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I emulator_b200/csrc -o tools/microbench/tile_ops tools/microbench/tile_ops.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include "tile_ops.cuh"
using namespace eb;
constexpr int PS = TPS;
__global__ void __launch_bounds__(256, 2) bench(const double* tile, double* out, long long* cyc, int reps) {
  extern __shared__ __align__(16) double sm[];
  double* sC = sm;
  double* sL = sm + 64 * PS;
  double* rd = sm + 2 * 64 * PS;
  double* sOut = sm + 2 * 64 * PS + 64;
  __shared__ int info;
  __shared__ int flags[16];
  int* gfl = reinterpret_cast<int*>(cyc + 4);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < 4096; e += 256) sC[(e >> 6) * PS + (e & 63)] = tile[e];
  if (tid == 0) info = 0;
  if (tid < 16) flags[tid] = 0;
  __syncthreads();
  double v[2][8];
  long long t_chol = 0, t_trsm = 0;
  int epoch = 0;
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
      for (int c = 0; c < 8; ++c) v[sl][c] = sC[(lane + 32 * sl) * PS + 8 * warp + c];
    __syncthreads();
    long long t0 = clock64();
    tile_cholesky(v, sL, rd, &info, 0, flags, ++epoch, out, 64, out + 4096 + 4096, gfl);
    long long t1 = clock64();
    t_chol += t1 - t0;
  }
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
      for (int c = 0; c < 8; ++c) v[sl][c] = sC[(lane + 32 * sl) * PS + 8 * warp + c];
    __syncthreads();
    long long t0 = clock64();
    tile_trsm(v, sL, rd, sOut);
    long long t1 = clock64();
    t_trsm += t1 - t0;
  }
  if (tid == 255) { cyc[0] = t_chol / reps; cyc[1] = t_trsm / reps; cyc[2] = info; }
  for (int e = tid; e < 4096; e += 256) out[e] = sL[(e >> 6) * PS + (e & 63)];
  for (int e = tid; e < 4096; e += 256) out[4096 + e] = sOut[(e >> 6) * PS + (e & 63)];
}
int main() {
  std::vector<double> a(4096), g(4096);
  for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) g[i * 64 + j] = std::sin(0.37 * i + 0.11 * j * j);
  for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) { double s = (i == j) ? 8.0 : 0.0; for (int k = 0; k < 64; ++k) s += g[i * 64 + k] * g[j * 64 + k] / 64.0; a[i * 64 + j] = s; }
  double *d_t, *d_o; long long* d_c;
  cudaMalloc(&d_t, 4096 * 8); cudaMalloc(&d_o, (2 * 4096 + 64) * 8); cudaMalloc(&d_c, 16 * 8); cudaMemset(d_c, 0, 16 * 8);
  cudaMemcpy(d_t, a.data(), 4096 * 8, cudaMemcpyHostToDevice);
  const int smem = (3 * 64 * PS + 128) * 8;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<<<1, 256, smem>>>(d_t, d_o, d_c, 50);
  cudaError_t e = cudaDeviceSynchronize();
  long long c[3]; cudaMemcpy(c, d_c, 24, cudaMemcpyDeviceToHost);
  std::vector<double> o(8192); cudaMemcpy(o.data(), d_o, 8192 * 8, cudaMemcpyDeviceToHost);
  double err = 0, err2 = 0, up = 0;
  for (int i = 0; i < 64; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int k = 0; k <= j; ++k) s += o[i * 64 + k] * o[j * 64 + k]; err = fmax(err, fabs(s - a[i * 64 + j])); }
  for (int i = 0; i < 64; ++i) for (int j = i + 1; j < 64; ++j) up = fmax(up, fabs(o[i * 64 + j]));
  for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) { double s = 0; for (int k = 0; k <= j; ++k) s += o[4096 + i * 64 + k] * o[j * 64 + k]; err2 = fmax(err2, fabs(s - a[i * 64 + j])); }
  printf("%s  tile_cholesky %lld cycles, tile_trsm %lld cycles, info %lld, |LL^T-A| %.2e (upper %.1e), |YL^T-A| %.2e\n", cudaGetErrorString(e), c[0], c[1], c[2], err, up, err2);
  return 0;
}
