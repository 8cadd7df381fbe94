// Phase timing inside the warp-synchronous 64x64 Cholesky (development tool; copy of tile_cholesky with probes).
#include <cstdio>
#include <vector>
#include <cmath>
#include "factor_dataflow.cuh"
using namespace eb;
__device__ long long g_probe[8][8];
__device__ __forceinline__ void tile_cholesky_probe(double (&v)[2][8], double* sL, double* rd, int* info, int col0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int p = 0; p < 8; ++p) {
    long long t0 = clock64();
    if (warp == p) {
      const int sj = p >> 2;
      const int ob = 8 * (p & 3);
      double dloc = sj ? v[1][0] : v[0][0];
      double d = __shfl_sync(0xffffffffu, dloc, ob);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (!(d > 0.0) && lane == 0) atomicCAS(info, 0, col0 + 8 * p + c + 1);
        const double rinv = rsqrt(d);
        if (lane == ob + c) rd[8 * p + c] = rinv;
        const double l0 = v[0][c] * rinv, l1 = v[1][c] * rinv;
        v[0][c] = l0; v[1][c] = l1;
        const double lsel = sj ? l1 : l0;
        if (c < 7) {
          const double nloc = fma(-lsel, lsel, sj ? v[1][c + 1] : v[0][c + 1]);
          d = __shfl_sync(0xffffffffu, nloc, ob + c + 1);
        }
#pragma unroll
        for (int c2 = c + 1; c2 < 8; ++c2) {
          const double lk = __shfl_sync(0xffffffffu, lsel, ob + c2);
          v[0][c2] = fma(-l0, lk, v[0][c2]);
          v[1][c2] = fma(-l1, lk, v[1][c2]);
        }
      }
      long long t1 = clock64();
#pragma unroll
      for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int row = lane + 32 * sl, col = 8 * p + c;
          sL[row * PS + col] = (row >= col) ? v[sl][c] : 0.0;
        }
      long long t2 = clock64();
      if (lane == 0) { g_probe[p][0] = t1 - t0; g_probe[p][1] = t2 - t1; }
    }
    long long t3 = clock64();
    __syncthreads();
    long long t4 = clock64();
    if (warp > p) {
      double ar[2][8];
#pragma unroll
      for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int k = 0; k < 8; ++k) ar[sl][k] = sL[(lane + 32 * sl) * PS + 8 * p + k];
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const double b = sL[(8 * warp + c) * PS + 8 * p + k];
          v[0][c] = fma(-ar[0][k], b, v[0][c]);
          v[1][c] = fma(-ar[1][k], b, v[1][c]);
        }
    }
    long long t5 = clock64();
    if (warp == p + 1 && lane == 0) { g_probe[p][2] = t4 - t3; g_probe[p][3] = t5 - t4; }
  }
}
__global__ void __launch_bounds__(256, 2) bench(const double* tile, long long* cyc) {
  extern __shared__ __align__(16) double sm[];
  double* sC = sm; double* sL = sm + 64 * PS; double* rd = sm + 2 * 64 * PS;
  __shared__ int info;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < 4096; e += 256) sC[(e >> 6) * PS + (e & 63)] = tile[e];
  if (tid == 0) info = 0;
  __syncthreads();
  double v[2][8];
  for (int r = 0; r < 3; ++r) {
    for (int sl = 0; sl < 2; ++sl) for (int c = 0; c < 8; ++c) v[sl][c] = sC[(lane + 32 * sl) * PS + 8 * warp + c];
    __syncthreads();
    long long t0 = clock64();
    tile_cholesky_probe(v, sL, rd, &info, 0);
    if (tid == 0) cyc[0] = clock64() - t0;
  }
}
int main() {
  std::vector<double> a(4096), g(4096);
  for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) g[i * 64 + j] = std::sin(0.37 * i + 0.11 * j * j);
  for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) { double s = (i == j) ? 8.0 : 0.0; for (int k = 0; k < 64; ++k) s += g[i * 64 + k] * g[j * 64 + k] / 64.0; a[i * 64 + j] = s; }
  double* d_t; long long* d_c; cudaMalloc(&d_t, 4096 * 8); cudaMalloc(&d_c, 8);
  cudaMemcpy(d_t, a.data(), 4096 * 8, cudaMemcpyHostToDevice);
  const int smem = (3 * 64 * PS + 128) * 8;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<<<1, 256, smem>>>(d_t, d_c);
  cudaDeviceSynchronize();
  long long c, pr[8][8]; cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost); cudaMemcpyFromSymbol(pr, g_probe, sizeof(pr));
  printf("total %lld\n", c);
  for (int p = 0; p < 8; ++p) printf("panel %d: factor %lld  write %lld | next warp: barrier-wait %lld  update %lld\n", p, pr[p][0], pr[p][1], pr[p][2], pr[p][3]);
  return 0;
}
