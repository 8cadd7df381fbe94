// Per-warp timeline of tile_cholesky (development tool): instrumented copy.
#include <cstdio>
#include <vector>
#include <cmath>
#include "tile_ops.cuh"
using namespace eb;
__device__ long long g_t[8][6];
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
__device__ __forceinline__ void chol_probe(double (&v)[2][8], double* sL, double* rd, int* info, volatile int* flags, int epoch, long long t00) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int p = 0; p + 1 < warp; ++p) { panel_wait(flags + p, epoch); stand_back(); if (warp >= 4) apply_panel<true>(v, sL, sL, p, lane, warp); else apply_panel<false>(v, sL, sL, p, lane, warp); }
  long long t1 = clk();
  if (warp > 0) { panel_wait(flags + warp - 1, epoch); }
  long long t2 = clk();
  if (warp > 0) { if (warp >= 4) apply_panel<true>(v, sL, sL, warp - 1, lane, warp); else apply_panel<false>(v, sL, sL, warp - 1, lane, warp); }
  long long t3 = clk();
  {
    const int p = warp; const int sj = p >> 2; const int ob = 8 * (p & 3);
    double d = shfl_idx_f64(sj ? v[1][0] : v[0][0], ob);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const double rinv = rsqrt_refined(d);
      const double l0 = v[0][c] * rinv, l1 = v[1][c] * rinv;
      v[0][c] = l0; v[1][c] = l1;
      const double lsel = sj ? l1 : l0;
      if (c < 7) { const double nloc = fma(-lsel, lsel, sj ? v[1][c + 1] : v[0][c + 1]); d = shfl_idx_f64(nloc, ob + c + 1); }
      if (lane == ob + c) rd[8 * p + c] = rinv;
#pragma unroll
      for (int c2 = c + 1; c2 < 8; ++c2) { const double lk = shfl_idx_f64(lsel, ob + c2); v[0][c2] = fma(-l0, lk, v[0][c2]); v[1][c2] = fma(-l1, lk, v[1][c2]); }
    }
    long long t4 = clk();
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
      for (int c = 0; c < 8; ++c) { const int row = lane + 32 * sl, col = 8 * p + c; sL[row * TPS + col] = (row >= col) ? v[sl][c] : 0.0; }
    panel_signal(flags + p, epoch);
    long long t5 = clk();
    if (lane == 0) { g_t[warp][0] = t1 - t00; g_t[warp][1] = t2 - t00; g_t[warp][2] = t3 - t00; g_t[warp][3] = t4 - t00; g_t[warp][4] = t5 - t00; }
  }
  __syncthreads();
}
__global__ void __launch_bounds__(256, 2) bench(const double* tile) {
  extern __shared__ __align__(16) double sm[];
  double* sC = sm; double* sL = sm + 64 * TPS; double* rd = sm + 2 * 64 * TPS;
  __shared__ int info; __shared__ int flags[8]; __shared__ long long t00s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < 4096; e += 256) sC[(e >> 6) * TPS + (e & 63)] = tile[e];
  if (tid < 8) flags[tid] = 0;
  __syncthreads();
  double v[2][8];
  for (int r = 1; r <= 3; ++r) {
    for (int sl = 0; sl < 2; ++sl) for (int c = 0; c < 8; ++c) v[sl][c] = sC[(lane + 32 * sl) * TPS + 8 * warp + c];
    __syncthreads();
    if (tid == 0) t00s = clk();
    __syncthreads();
    chol_probe(v, sL, rd, &info, flags, r, t00s);
  }
}
int main() {
  std::vector<double> a(4096), g(4096);
  for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) g[i * 64 + j] = std::sin(0.37 * i + 0.11 * j * j);
  for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) { double s = (i == j) ? 8.0 : 0.0; for (int k = 0; k < 64; ++k) s += g[i * 64 + k] * g[j * 64 + k] / 64.0; a[i * 64 + j] = s; }
  double* d_t; cudaMalloc(&d_t, 4096 * 8); cudaMemcpy(d_t, a.data(), 4096 * 8, cudaMemcpyHostToDevice);
  const int smem = (3 * 64 * TPS + 128) * 8;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<<<1, 256, smem>>>(d_t); cudaDeviceSynchronize();
  long long t[8][6]; cudaMemcpyFromSymbol(t, g_t, sizeof(t));
  printf("warp: earlier-updates-done  panel(w-1)-seen  update-done  factor-done  signalled   (cycles since start)\n");
  for (int w = 0; w < 8; ++w) printf("%d: %6lld %6lld %6lld %6lld %6lld   | wait->update %lld, factor %lld, write+signal %lld\n", w, t[w][0], t[w][1], t[w][2], t[w][3], t[w][4], t[w][2]-t[w][1], t[w][3]-t[w][2], t[w][4]-t[w][3]);
  return 0;
}
