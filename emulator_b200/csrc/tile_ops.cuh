// Warp-synchronous 64x64 tile finalisers for the Cholesky's dependency chain (sm_100a, fp64).
//
// Layout: warp w (of 8) owns columns 8w..8w+7 of the tile, lane owns rows lane and lane+32, so a
// thread holds v[2][8].  Both routines are right-looking by 8-column panels: the owning warp
// finishes its panel with shuffles only, writes it to shared memory and raises a per-panel flag;
// warps to its right wait for that flag (not for a CTA barrier: only the next warp is on the
// critical path, the others may lag) and apply the panel to their columns.
//
// Latency matters here, not throughput (measured on B200: DFMA 8.8, shuffle ~25, rsqrt 66 cycles):
//   * the shuffle that carries the next pivot is issued before the bulk shuffles of the current
//     column (volatile asm keeps ptxas from sinking it -- the generic version lost ~100 cycles per
//     column to that);
//   * 1/sqrt(pivot) is MUFU.RSQ64H plus one third-order refinement, without the special-case
//     branches of rsqrt() (a non-positive pivot just produces NaN and is reported after the panel).
#pragma once
#include <cuda_runtime.h>

namespace eb {

constexpr int TPS = 65;  // padded stride of 64x64 tiles in shared memory

__device__ __forceinline__ double shfl_idx_f64(double x, int src) {
  int lo = __double2loint(x), hi = __double2hiint(x);
  asm volatile(
      "shfl.sync.idx.b32 %0, %0, %2, 0x1f, 0xffffffff;\n\t"
      "shfl.sync.idx.b32 %1, %1, %2, 0x1f, 0xffffffff;"
      : "+r"(lo), "+r"(hi)
      : "r"(src));
  return __hiloint2double(hi, lo);
}

__device__ __forceinline__ double rsqrt_refined(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-(y * y), d, 1.0);  // 1 - d y^2
  return fma(fma(e, 0.375, 0.5), y * e, y);  // y (1 + e/2 + 3 e^2 / 8)
}

__device__ __forceinline__ void panel_signal(volatile int* flag, int epoch) {
  __threadfence_block();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) *flag = epoch;
}
__device__ __forceinline__ void panel_wait(volatile int* flag, int epoch) {
  while (*flag != epoch) {
  }
}

// v[s][c] -= sum_k rows[row_s][8p+k] * cols[8*warp+c][8p+k]   (rank-8 update with a finished panel).
// SKIP0: row slot 0 (rows 0..31) is not needed (Cholesky, columns >= 32: entirely above the diagonal).
template <bool SKIP0>
__device__ __forceinline__ void apply_panel(double (&v)[2][8], const double* rows, const double* cols, int p, int lane,
                                            int warp) {
  double ar[2][8];
#pragma unroll
  for (int sl = SKIP0 ? 1 : 0; sl < 2; ++sl)
#pragma unroll
    for (int k = 0; k < 8; ++k) ar[sl][k] = rows[(lane + 32 * sl) * TPS + 8 * p + k];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    // two partial sums per entry halve the dependent-FMA chain
    double s0a = 0.0, s0b = 0.0, s1a = 0.0, s1b = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      const double b0 = cols[(8 * warp + c) * TPS + 8 * p + k], b1 = cols[(8 * warp + c) * TPS + 8 * p + k + 1];
      if (!SKIP0) {
        s0a = fma(ar[0][k], b0, s0a);
        s0b = fma(ar[0][k + 1], b1, s0b);
      }
      s1a = fma(ar[1][k], b0, s1a);
      s1b = fma(ar[1][k + 1], b1, s1b);
    }
    if (!SKIP0) v[0][c] -= s0a + s0b;
    v[1][c] -= s1a + s1b;
  }
}

// A warp that is not next in line gives the critical warp a head start on the shared FP64 pipe;
// it has at least one whole panel factorisation (~1000 cycles of mostly idle pipe) to catch up.
__device__ __forceinline__ void stand_back() { __nanosleep(400); }

// In: v = lower triangle of an SPD tile (rows lane, lane+32; columns 8*warp..).  Out: L in sL
// (zero strict upper), reciprocal pivots in rd[64].  Returns (to every thread of the panel's warp)
// nothing; a non-positive pivot is recorded in *info as 1-based global column.  flags: 8 ints in
// shared memory; epoch: a value unique to this call.  Ends with a CTA barrier.
//
// Streaming: as soon as panel p is in shared memory, a warp that has nothing else to do (warp p-1,
// already finished; warp 7 for panel 0) copies it -- 64 rows x 8 columns of L plus 8 reciprocal
// pivots -- to global memory (Lg, ld; rd_g) and publishes the global flag gflags[p], so that the
// tiles below the diagonal start their substitution against panel p while the chain is still
// factorising panel p+1 (see tile_trsm_streamed).
__device__ __forceinline__ void stream_panel(const double* sL, const double* rd, int p, double* Lg, int ld, double* rd_g,
                                             int* gflags, int lane) {
#pragma unroll
  for (int sl = 0; sl < 2; ++sl) {
    const int row = lane + 32 * sl;
    const double* src = sL + row * TPS + 8 * p;
    double* dst = Lg + (size_t)row * ld + 8 * p;
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
      double2 val;
      val.x = src[c];
      val.y = src[c + 1];
      *reinterpret_cast<double2*>(dst + c) = val;
    }
  }
  if (lane < 8) rd_g[8 * p + lane] = rd[8 * p + lane];
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(gflags + p), "r"(1) : "memory");
  }
}

__device__ __forceinline__ void tile_cholesky(double (&v)[2][8], double* sL, double* rd, int* info, int col0,
                                              volatile int* flags, int epoch, double* Lg, int ld, double* rd_g,
                                              int* gflags) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 7) {  // nothing to do until panel 0 exists: stream it out first
    panel_wait(flags + 0, epoch);
    stream_panel(sL, rd, 0, Lg, ld, rd_g, gflags, lane);
  }
  // panels 0 .. warp-2: apply as they appear (this warp is not yet on the critical path)
  for (int p = 0; p + 1 < warp; ++p) {
    panel_wait(flags + p, epoch);
    stand_back();
    if (warp >= 4)
      apply_panel<true>(v, sL, sL, p, lane, warp);
    else
      apply_panel<false>(v, sL, sL, p, lane, warp);
  }
  // panel warp-1 is applied in the same basic block as the factorisation of the own panel, so that
  // the scheduler starts the pivot chain as soon as column 0 is up to date and hides the rest of
  // the update behind the chain's latency
  if (warp > 0) {
    panel_wait(flags + warp - 1, epoch);
    if (warp >= 4)
      apply_panel<true>(v, sL, sL, warp - 1, lane, warp);
    else
      apply_panel<false>(v, sL, sL, warp - 1, lane, warp);
  }
  {
    const int p = warp;
    const int sj = p >> 2;       // row slot of this panel's pivot rows
    const int ob = 8 * (p & 3);  // lane of the first pivot row
    double d = shfl_idx_f64(sj ? v[1][0] : v[0][0], ob);
    bool bad = false;
    int bad_col = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (!(d > 0.0) && !bad) {
        bad = true;
        bad_col = c;
      }
      const double rinv = rsqrt_refined(d);
      const double l0 = v[0][c] * rinv, l1 = v[1][c] * rinv;
      v[0][c] = l0;
      v[1][c] = l1;
      const double lsel = sj ? l1 : l0;
      if (c < 7) {
        // next pivot, computed by its own lane and fetched first
        const double nloc = fma(-lsel, lsel, sj ? v[1][c + 1] : v[0][c + 1]);
        d = shfl_idx_f64(nloc, ob + c + 1);
      }
      if (lane == ob + c) rd[8 * p + c] = rinv;
#pragma unroll
      for (int c2 = c + 1; c2 < 8; ++c2) {
        const double lk = shfl_idx_f64(lsel, ob + c2);
        v[0][c2] = fma(-l0, lk, v[0][c2]);
        v[1][c2] = fma(-l1, lk, v[1][c2]);
      }
    }
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int row = lane + 32 * sl, col = 8 * p + c;
        sL[row * TPS + col] = (row >= col) ? v[sl][c] : 0.0;
      }
    panel_signal(flags + p, epoch);
    if (bad && lane == 0) atomicCAS(info, 0, col0 + 8 * p + bad_col + 1);
    if (warp < 7) {  // done with everything of its own: stream the next panel out when it appears
      panel_wait(flags + warp + 1, epoch);
      stream_panel(sL, rd, warp + 1, Lg, ld, rd_g, gflags, lane);
    }
  }
  __syncthreads();
}

// Y = V L^-T for the tile held in v (rows independent), L in sL, reciprocal pivots in rd; forward
// substitution by 8-column panels with a CTA barrier per panel (measured: the flag-synchronised
// variant is slower here -- 9.8k vs 7.6k cycles -- because every warp has the same amount of
// update work and nothing to hide it behind).  Result in sOut ([64][TPS]).  The kernels use the streamed
// variant below; this one is the reference point of tools/microbench/tile_ops.cu.
__device__ __forceinline__ void tile_trsm(double (&v)[2][8], const double* sL, const double* rd, double* sOut) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int p = 0; p < 8; ++p) {
    if (warp == p) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double r = rd[8 * p + c];
        const double x0 = v[0][c] * r, x1 = v[1][c] * r;
        v[0][c] = x0;
        v[1][c] = x1;
#pragma unroll
        for (int c2 = c + 1; c2 < 8; ++c2) {
          const double lv = sL[(8 * p + c2) * TPS + 8 * p + c];
          v[0][c2] = fma(-x0, lv, v[0][c2]);
          v[1][c2] = fma(-x1, lv, v[1][c2]);
        }
      }
#pragma unroll
      for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int c = 0; c < 8; ++c) sOut[(lane + 32 * sl) * TPS + 8 * p + c] = v[sl][c];
    }
    __syncthreads();
    if (warp > p) {
      double ar[2][8];
#pragma unroll
      for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int k = 0; k < 8; ++k) ar[sl][k] = sOut[(lane + 32 * sl) * TPS + 8 * p + k];
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const double b = sL[(8 * warp + c) * TPS + 8 * p + k];
          v[0][c] = fma(-ar[0][k], b, v[0][c]);
          v[1][c] = fma(-ar[1][k], b, v[1][c]);
        }
    }
  }
}

// tile_trsm against a diagonal tile that is still being factorised by another CTA: panel p of L
// (and its reciprocal pivots) is fetched from global memory as soon as its flag is published.
// Lg/ld/rd_g/gflags describe the diagonal tile; returns false if the watchdog of the poll fired.
template <class WaitFn>
__device__ __forceinline__ bool tile_trsm_streamed(double (&v)[2][8], double* sL, double* rd, double* sOut, const double* Lg,
                                                    int ld, const double* rd_g, const int* gflags, WaitFn wait_flag) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  bool ok = true;
  for (int p = 0; p < 8; ++p) {
    ok = wait_flag(gflags + p) && ok;
    {
      // 64 rows x 8 columns: thread t loads row t/4, columns 2*(t%4), +1
      const int row = tid >> 2, c2 = (tid & 3) * 2;
      const double2 val = __ldcg(reinterpret_cast<const double2*>(Lg + (size_t)row * ld + 8 * p + c2));
      sL[row * TPS + 8 * p + c2] = val.x;
      sL[row * TPS + 8 * p + c2 + 1] = val.y;
      if (tid < 8) rd[8 * p + tid] = __ldcg(rd_g + 8 * p + tid);
    }
    __syncthreads();
    if (warp == p) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double r = rd[8 * p + c];
        const double x0 = v[0][c] * r, x1 = v[1][c] * r;
        v[0][c] = x0;
        v[1][c] = x1;
#pragma unroll
        for (int c2 = c + 1; c2 < 8; ++c2) {
          const double lv = sL[(8 * p + c2) * TPS + 8 * p + c];
          v[0][c2] = fma(-x0, lv, v[0][c2]);
          v[1][c2] = fma(-x1, lv, v[1][c2]);
        }
      }
#pragma unroll
      for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int c = 0; c < 8; ++c) sOut[(lane + 32 * sl) * TPS + 8 * p + c] = v[sl][c];
    }
    __syncthreads();
    if (warp > p) {
      double ar[2][8];
#pragma unroll
      for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int k = 0; k < 8; ++k) ar[sl][k] = sOut[(lane + 32 * sl) * TPS + 8 * p + k];
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const double b = sL[(8 * warp + c) * TPS + 8 * p + k];
          v[0][c] = fma(-ar[0][k], b, v[0][c]);
          v[1][c] = fma(-ar[1][k], b, v[1][c]);
        }
    }
  }
  return ok;
}

}  // namespace eb
