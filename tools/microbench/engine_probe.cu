// Where does the 128x128 DMMA tile loop lose its last 12 %?  (development tool)
// Shared-memory-resident operands (no global loads): the inner loop of engine.cuh with its fragment loads and
// DMMAs, varying the number of warps, the barrier, and the order of loads and math.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o engine_probe engine_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int KC = 16, KS = KC + 8;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// BM x BN tile, warps arranged WARPS_M x WARPS_N, each WM x WN.  STAGES resident stages are cycled through.
template <int BM, int BN, int WM, int WN, bool BARRIER, bool PIPELINED, int MINB>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, MINB) probe(double* out, int pairs) {
  extern __shared__ __align__(16) double sm[];
  constexpr int WARPS_N = BN / WN, MI = WM / 8, NI = WN / 8, THREADS = (BM / WM) * (BN / WN) * 32;
  constexpr int STAGE = (BM + BN) * KS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp / WARPS_N) * WM, wn0 = (warp % WARPS_N) * WN, g = lane >> 2, t = lane & 3;
  for (int e = tid; e < 4 * STAGE; e += THREADS) sm[e] = 1e-3 * (e % 97);
  __syncthreads();
  double acc[MI][NI][2];
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  for (int p = 0; p < pairs; ++p) {
    if (BARRIER) __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const double* sA = sm + ((2 * p + h) & 3) * STAGE;
      const double* sB = sA + BM * KS;
      if (!PIPELINED) {
#pragma unroll
        for (int kk = 0; kk < KC; kk += 8) {
          double af[2][MI], bf[2][NI];
#pragma unroll
          for (int i = 0; i < MI; ++i) {
            const double2 v = *reinterpret_cast<const double2*>(sA + (wm0 + g + 8 * i) * KS + kk + 2 * t);
            af[0][i] = v.x; af[1][i] = v.y;
          }
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            const double2 v = *reinterpret_cast<const double2*>(sB + (wn0 + g + 8 * i) * KS + kk + 2 * t);
            bf[0][i] = v.x; bf[1][i] = v.y;
          }
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
              for (int ni = 0; ni < NI; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], af[h2][mi], bf[h2][ni]);
        }
      } else {
        // fragments of block kk+8 are requested before the math of block kk (register double buffering)
        double af[2][2][MI], bf[2][2][NI];
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const double2 v = *reinterpret_cast<const double2*>(sA + (wm0 + g + 8 * i) * KS + 2 * t);
          af[0][0][i] = v.x; af[0][1][i] = v.y;
        }
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const double2 v = *reinterpret_cast<const double2*>(sB + (wn0 + g + 8 * i) * KS + 2 * t);
          bf[0][0][i] = v.x; bf[0][1][i] = v.y;
        }
#pragma unroll
        for (int blk = 0; blk < KC / 8; ++blk) {
          const int cur = blk & 1, nxt = cur ^ 1;
          if (blk + 1 < KC / 8) {
#pragma unroll
            for (int i = 0; i < MI; ++i) {
              const double2 v = *reinterpret_cast<const double2*>(sA + (wm0 + g + 8 * i) * KS + (blk + 1) * 8 + 2 * t);
              af[nxt][0][i] = v.x; af[nxt][1][i] = v.y;
            }
#pragma unroll
            for (int i = 0; i < NI; ++i) {
              const double2 v = *reinterpret_cast<const double2*>(sB + (wn0 + g + 8 * i) * KS + (blk + 1) * 8 + 2 * t);
              bf[nxt][0][i] = v.x; bf[nxt][1][i] = v.y;
            }
          }
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
              for (int ni = 0; ni < NI; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], af[cur][h2][mi], bf[cur][h2][ni]);
        }
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) s += acc[mi][ni][0] + acc[mi][ni][1];
  out[(size_t)blockIdx.x * THREADS + tid] = s;
}

template <int BM, int BN, int WM, int WN, bool BARRIER, bool PIPELINED, int MINB>
void run(const char* name, double* out, int ctas_per_sm) {
  auto k = probe<BM, BN, WM, WN, BARRIER, PIPELINED, MINB>;
  const int smem = 4 * (BM + BN) * KS * 8;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int threads = (BM / WM) * (BN / WN) * 32, grid = 148 * ctas_per_sm, pairs = 4000;
  k<<<grid, threads, smem>>>(out, 50);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); k<<<grid, threads, smem>>>(out, pairs); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
  }
  cudaError_t err = cudaGetLastError();
  const double flops = 2.0 * BM * BN * KC * 2.0 * pairs * grid;
  int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, threads, smem);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
  printf("%-58s %6.2f TFLOP/s  (%d thr, %d CTA/SM launched, occupancy %d, %d regs)%s\n", name, flops / best / 1e9, threads, ctas_per_sm, nb,
         fa.numRegs, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
  double* out; cudaMalloc(&out, (size_t)148 * 2 * 1024 * 8);
  run<128, 128, 64, 32, true, false, 1>("128x128, 8 warps 64x32, barrier/pair (engine as is)", out, 1);
  run<128, 128, 64, 32, false, false, 1>("128x128, 8 warps 64x32, no barrier", out, 1);
  run<128, 128, 64, 32, true, true, 1>("128x128, 8 warps 64x32, barrier, register double buffer", out, 1);
  run<128, 128, 32, 32, true, false, 1>("128x128, 16 warps 32x32, barrier", out, 1);
  run<128, 128, 32, 32, false, false, 1>("128x128, 16 warps 32x32, no barrier", out, 1);
  run<128, 128, 32, 32, true, true, 1>("128x128, 16 warps 32x32, barrier, register double buffer", out, 1);
  run<128, 128, 32, 64, true, false, 1>("128x128, 8 warps 32x64, barrier", out, 1);
  run<128, 64, 32, 32, true, false, 2>("128x64, 8 warps 32x32, barrier, 2 CTA/SM", out, 2);
  run<128, 64, 32, 32, true, true, 2>("128x64, 8 warps 32x32, barrier, reg dbuf, 2 CTA/SM", out, 2);
  run<128, 64, 64, 32, true, false, 2>("128x64, 4 warps 64x32, barrier, 2 CTA/SM", out, 2);
  run<128, 64, 64, 32, true, false, 3>("128x64, 4 warps 64x32, barrier, 3 CTA/SM", out, 3);
  run<64, 64, 32, 16, true, false, 2>("64x64, 8 warps 32x16, barrier, 2 CTA/SM (factor loop)", out, 2);
  run<64, 64, 32, 16, true, true, 2>("64x64, 8 warps 32x16, barrier, reg dbuf, 2 CTA/SM", out, 2);
  run<64, 64, 32, 32, true, false, 2>("64x64, 4 warps 32x32, barrier, 2 CTA/SM", out, 2);
  run<64, 64, 32, 32, true, false, 4>("64x64, 4 warps 32x32, barrier, 4 CTA/SM", out, 4);
  run<256, 128, 64, 64, true, false, 1>("256x128, 8 warps 64x64, barrier", out, 1);
  return 0;
}
