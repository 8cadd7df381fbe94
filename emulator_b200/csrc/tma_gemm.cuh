// TMA-fed variant of the 128x128 fp64 tile loop (sm_100a): one producer warp moves the operand slabs with
// cp.async.bulk.tensor (one instruction per 128x16 slab, no per-thread address arithmetic), eight consumer warps
// do nothing but fragment loads and DMMA.8x8x4, and the two sides meet at per-stage mbarriers instead of a
// CTA-wide barrier.
//
// Why: with the operands already in shared memory the loop of engine.cuh reaches 36.6 of the 37.1 TFLOP/s DMMA
// peak (tools/microbench/engine_probe.cu); fed through per-thread cp.async it reached 31.6 (ncu: 88 % of the
// tensor pipe) -- the same eight warps that issue DMMAs spent a sixth of their instruction stream on LDGSTS
// address arithmetic and met at a barrier every two chunks, and with two warps per scheduler nothing covers for a
// warp that is busy issuing loads.
//
// Shared-memory layout: dense 128-byte rows (16 doubles of k) written by TMA with the 128-byte swizzle: the
// 16-byte chunk c of row R sits at chunk c ^ (R & 7).  Lane (g, t) of a warp feeds k = 4t, 4t+1 (chunk 2t) to the
// first two DMMAs of a slab and k = 4t+2, 4t+3 (chunk 2t+1) to the last two -- any assignment of k to lanes is
// legal as long as A and B agree -- so the eight lanes of an LDS.128 phase (two rows, four chunks each) hit the
// even chunk positions for one row and the odd ones for the other: conflict-free without padding.
#pragma once
#include <cuda.h>

#include "engine.cuh"

namespace eb {

constexpr int TMA_STAGES = 6;                         // 6 x 32 KB ring
constexpr int TMA_SLAB_BYTES = 128 * KC * 8;          // one operand slab: 128 rows x 16 doubles
constexpr int TMA_STAGE_BYTES = 2 * TMA_SLAB_BYTES;   // A and B
constexpr int TMA_SMEM_BYTES = TMA_STAGES * TMA_STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr int TMA_CONSUMER_WARPS = 8;
constexpr int TMA_THREADS = (TMA_CONSUMER_WARPS + 1) * 32;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
// Wait for a phase of the barrier.  A stage that never completes (a transfer that cannot be made) would hang the
// launch: after 20 s without progress the kernel traps instead, which the host sees as a launch failure.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned spins = 0;
  unsigned long long t0 = 0ull;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xfffffu) == 0u) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0ull) t0 = now;
      else if (now - t0 > 20ull * 1000ull * 1000ull * 1000ull) __trap();
    }
  }
}
// 2-D tiled TMA load: box (16 doubles of k, 128 rows) at coordinates (k0, row0) -> dst, completion on `bar`.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int k0, int row0, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(k0), "r"(row0)
               : "memory");
}

// Fragments of N 8-row blocks for two consecutive DMMAs out of a swizzled slab: rows row0 + 8i (row0 % 8 == g),
// 16-byte chunk `chunk` (before the swizzle).
template <int N>
__device__ __forceinline__ void load_fragments_swizzled(double (&f)[2][N], const double* slab, int row0, int g, int chunk) {
  const char* base = reinterpret_cast<const char*>(slab) + row0 * 128 + ((chunk ^ g) << 4);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double2 v = *reinterpret_cast<const double2*>(base + i * 8 * 128);
    f[0][i] = v.x;
    f[1][i] = v.y;
  }
}

// The shared state of one CTA's pipeline.
struct TmaPipe {
  double* slabs;                // [TMA_STAGES][2][128][16], 1024-byte aligned
  unsigned long long* full;     // [TMA_STAGES] bytes have landed
  unsigned long long* empty;    // [TMA_STAGES] every consumer warp is done with the stage

  __device__ __forceinline__ void carve(double* dyn_smem) {
    const unsigned long long a = (reinterpret_cast<unsigned long long>(dyn_smem) + 1023ull) & ~1023ull;
    slabs = reinterpret_cast<double*>(a);
    full = reinterpret_cast<unsigned long long*>(a + (size_t)TMA_STAGES * TMA_STAGE_BYTES);
    empty = full + TMA_STAGES;
  }
  // (full_arrivals: 1 = the TMA-issuing thread; more when other warps write part of a stage themselves)
  __device__ __forceinline__ void init(unsigned full_arrivals = 1) {  // one thread; followed by a CTA barrier in the caller
    for (int s = 0; s < TMA_STAGES; ++s) {
      mbar_init(full + s, full_arrivals);
      mbar_init(empty + s, TMA_CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
};

// Producer side (one lane): slabs of A at rows a_row, of B at rows b_row, k chunks kt0 .. kt0 + nk - 1 starting at
// columns a_col / b_col.  `it` counts the chunks this CTA has pushed through the ring so far (persistent use).
__device__ __forceinline__ void tma_produce(const TmaPipe& pipe, const CUtensorMap* mapA, const CUtensorMap* mapB, int a_row, int a_col,
                                            int b_row, int b_col, int nk, unsigned& it) {
  for (int kt = 0; kt < nk; ++kt, ++it) {
    const unsigned stage = it % TMA_STAGES, round = it / TMA_STAGES;
    if (round > 0) mbar_wait(pipe.empty + stage, (round - 1) & 1);
    mbar_expect_tx(pipe.full + stage, TMA_STAGE_BYTES);
    double* dst = pipe.slabs + (size_t)stage * (TMA_STAGE_BYTES / 8);
    tma_load_2d(dst, mapA, a_col + kt * KC, a_row, pipe.full + stage);
    tma_load_2d(dst + TMA_SLAB_BYTES / 8, mapB, b_col + kt * KC, b_row, pipe.full + stage);
  }
}

// Consumer side (8 warps, 64 x 32 sub-tiles of the 128 x 128 tile): acc += A B^T over nk chunks.
// `active`: whole warps may skip the math (strictly-upper quadrant of a diagonal tile) while still taking part in
// the pipeline's bookkeeping.
__device__ __forceinline__ void tma_consume(const TmaPipe& pipe, double (&acc)[Cfg128::MI][Cfg128::NI][2], int nk, unsigned& it,
                                            bool active = true) {
  using Cfg = Cfg128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WM, wn0 = (warp % Cfg::WARPS_N) * Cfg::WN;
  const int g = lane >> 2, t = lane & 3;
  for (int kt = 0; kt < nk; ++kt, ++it) {
    const unsigned stage = it % TMA_STAGES, round = it / TMA_STAGES;
    mbar_wait(pipe.full + stage, round & 1);
    const double* sA = pipe.slabs + (size_t)stage * (TMA_STAGE_BYTES / 8);
    const double* sB = sA + TMA_SLAB_BYTES / 8;
    if (active) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        double af[2][Cfg::MI], bf[2][Cfg::NI];
        load_fragments_swizzled<Cfg::MI>(af, sA, wm0 + g, g, 2 * t + half);
        load_fragments_swizzled<Cfg::NI>(bf, sB, wn0 + g, g, 2 * t + half);
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
          for (int mi = 0; mi < Cfg::MI; ++mi)
#pragma unroll
            for (int ni = 0; ni < Cfg::NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[h2][mi], bf[h2][ni]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(pipe.empty + stage);
  }
}

__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(TMA_CONSUMER_WARPS * 32) : "memory"); }

// ---------------------------------------------------------------------------------------
// Predictive variance through the TMA pipeline (see var_gemm_kernel in kernels.cuh for the mathematics):
// one CTA per 128 x 128 tile of V = K* X^T, reduced to per-row sums of squares in the epilogue.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TMA_THREADS, 1) var_gemm_tma_kernel(const TileTask* __restrict__ tasks,
                                                                      const __grid_constant__ CUtensorMap mapK,
                                                                      const __grid_constant__ CUtensorMap mapX,
                                                                      double* __restrict__ rowss, int ld_rowss) {
  extern __shared__ __align__(16) double gsm[];
  __shared__ double sRow[4 * 128];
  using Cfg = Cfg128;
  TmaPipe pipe;
  pipe.carve(gsm);
  if (threadIdx.x == 0) pipe.init();
  __syncthreads();
  const TileTask tk = tasks[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned it = 0;
  if (warp == TMA_CONSUMER_WARPS) {
    if (lane == 0) tma_produce(pipe, &mapK, &mapX, tk.a_row, tk.a_col, tk.b_row, tk.b_col, tk.nk, it);
    return;
  }
  double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
  for (int mi = 0; mi < Cfg::MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < Cfg::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  tma_consume(pipe, acc, tk.nk, it);
  const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WM;
  const int g = lane >> 2, t = lane & 3;
  const int wn = warp % Cfg::WARPS_N;
#pragma unroll
  for (int mi = 0; mi < Cfg::MI; ++mi) {
    double v = 0.0;
#pragma unroll
    for (int ni = 0; ni < Cfg::NI; ++ni) {
      v = fma(acc[mi][ni][0], acc[mi][ni][0], v);
      v = fma(acc[mi][ni][1], acc[mi][ni][1], v);
    }
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (t == 0) sRow[wn * 128 + wm0 + mi * 8 + g] = v;
  }
  consumer_barrier();
  if (threadIdx.x < 128) {
    const double v = sRow[threadIdx.x] + sRow[128 + threadIdx.x] + sRow[256 + threadIdx.x] + sRow[384 + threadIdx.x];
    rowss[(size_t)(tk.c_col / 128) * ld_rowss + tk.c_row + threadIdx.x] = v;
  }
}

// ---------------------------------------------------------------------------------------
// Predictive variance of a prediction GRID with the cross-covariance built and consumed in one kernel: K* never
// exists in global memory.  The kernel factorises over its axes (kernels.cuh, predict_grid_mean_kernel):
// K*[q][j] = E1[b(q)][j] * E2[r(q)][j] with q = r * n_ind + b, E1 a table of n_ind x N entries and E2 one row of N
// entries per parameter row.  Four generator warps (one lane per query row of the 128-row tile) form each 128 x 16
// slab of K* from the two tables straight into the pipeline stage, in the layout TMA would have written (dense
// 128-byte rows, 16-byte chunk c of row R at position c ^ (R & 7)); the slab of X = L^-1 comes by TMA; the eight
// DMMA warps consume both.  Per 16384-row chunk the DRAM traffic falls from 537 MB written + several GB re-read
// (K*) to the 16 MB of E2.
// ---------------------------------------------------------------------------------------
constexpr int FUSED_GEN_WARPS = 4;
constexpr int FUSED_THREADS = (TMA_CONSUMER_WARPS + FUSED_GEN_WARPS) * 32;

__global__ void __launch_bounds__(FUSED_THREADS, 1) var_grid_fused_kernel(const TileTask* __restrict__ tasks,
                                                                          const __grid_constant__ CUtensorMap mapX,
                                                                          const double* __restrict__ e1, const double* __restrict__ e2,
                                                                          int np, size_t q0, int m_rows, int n_ind,
                                                                          double* __restrict__ rowss, int ld_rowss) {
  extern __shared__ __align__(16) double gsm[];
  __shared__ double sRow[4 * 128];
  using Cfg = Cfg128;
  TmaPipe pipe;
  pipe.carve(gsm);
  if (threadIdx.x == 0) pipe.init(1 + FUSED_GEN_WARPS);
  __syncthreads();
  const TileTask tk = tasks[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned it = 0;
  if (warp >= TMA_CONSUMER_WARPS) {
    // ---- generators: lane <-> query row of the tile; warp 8's lane 0 also issues the TMA of the X slab ----
    const int row = (warp - TMA_CONSUMER_WARPS) * 32 + lane;
    const size_t q = q0 + (size_t)tk.a_row + row;
    const bool live = tk.a_row + row < m_rows;
    const size_t r_first = q0 / n_ind;
    const double* e1p = e1 + (size_t)(q % n_ind) * np + tk.a_col;
    const double* e2p = e2 + (size_t)(q / n_ind - r_first) * np + tk.a_col;
    const bool tma_lane = warp == TMA_CONSUMER_WARPS && lane == 0;
    // the products of the NEXT chunk are formed while this one's stage is still in use (the table loads are L2
    // round trips): when the stage is released, the slab is written at once
    double2 nxt[8];
    auto form = [&](int kt) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        nxt[c] = make_double2(0.0, 0.0);
        if (live && kt < tk.nk) {
          const double2 a = __ldg(reinterpret_cast<const double2*>(e1p + kt * KC + 2 * c));
          const double2 b = __ldg(reinterpret_cast<const double2*>(e2p + kt * KC + 2 * c));
          nxt[c].x = a.x * b.x;
          nxt[c].y = a.y * b.y;
        }
      }
    };
    form(0);
    for (int kt = 0; kt < tk.nk; ++kt, ++it) {
      const unsigned stage = it % TMA_STAGES, round = it / TMA_STAGES;
      if (round > 0) mbar_wait(pipe.empty + stage, (round - 1) & 1);
      double* dst = pipe.slabs + (size_t)stage * (TMA_STAGE_BYTES / 8);
      if (tma_lane) {
        mbar_expect_tx(pipe.full + stage, TMA_SLAB_BYTES);
        tma_load_2d(dst + TMA_SLAB_BYTES / 8, &mapX, tk.b_col + kt * KC, tk.b_row, pipe.full + stage);
      }
      char* out = reinterpret_cast<char*>(dst) + row * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) *reinterpret_cast<double2*>(out + ((c ^ (row & 7)) << 4)) = nxt[c];
      __syncwarp();
      if (lane == 0) mbar_arrive(pipe.full + stage);
      form(kt + 1);
    }
    return;
  }
  double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
  for (int mi = 0; mi < Cfg::MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < Cfg::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  tma_consume(pipe, acc, tk.nk, it);
  const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WM;
  const int g = lane >> 2, t = lane & 3;
  const int wn = warp % Cfg::WARPS_N;
#pragma unroll
  for (int mi = 0; mi < Cfg::MI; ++mi) {
    double v = 0.0;
#pragma unroll
    for (int ni = 0; ni < Cfg::NI; ++ni) {
      v = fma(acc[mi][ni][0], acc[mi][ni][0], v);
      v = fma(acc[mi][ni][1], acc[mi][ni][1], v);
    }
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (t == 0) sRow[wn * 128 + wm0 + mi * 8 + g] = v;
  }
  consumer_barrier();
  if (threadIdx.x < 128) {
    const double v = sRow[threadIdx.x] + sRow[128 + threadIdx.x] + sRow[256 + threadIdx.x] + sRow[384 + threadIdx.x];
    rowss[(size_t)(tk.c_col / 128) * ld_rowss + tk.c_row + threadIdx.x] = v;
  }
}

// Plain tile GEMM through the same pipeline (micro-benchmark: C = A B^T, both k-contiguous).
__global__ void __launch_bounds__(TMA_THREADS, 1) tile_gemm_tma_kernel(const TileTask* __restrict__ tasks,
                                                                       const __grid_constant__ CUtensorMap mapA,
                                                                       const __grid_constant__ CUtensorMap mapB, double* C, int ldc) {
  extern __shared__ __align__(16) double gsm[];
  using Cfg = Cfg128;
  TmaPipe pipe;
  pipe.carve(gsm);
  if (threadIdx.x == 0) pipe.init();
  __syncthreads();
  const TileTask tk = tasks[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned it = 0;
  if (warp == TMA_CONSUMER_WARPS) {
    if (lane == 0) tma_produce(pipe, &mapA, &mapB, tk.a_row, tk.a_col, tk.b_row, tk.b_col, tk.nk, it);
    return;
  }
  double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
  for (int mi = 0; mi < Cfg::MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < Cfg::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  tma_consume(pipe, acc, tk.nk, it);
  const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WM, wn0 = (warp % Cfg::WARPS_N) * Cfg::WN;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mi = 0; mi < Cfg::MI; ++mi) {
    double* crow = C + (size_t)(tk.c_row + wm0 + mi * 8 + g) * ldc + tk.c_col + wn0 + 2 * t;
#pragma unroll
    for (int ni = 0; ni < Cfg::NI; ++ni) {
      double2 v;
      v.x = acc[mi][ni][0];
      v.y = acc[mi][ni][1];
      *reinterpret_cast<double2*>(crow + ni * 8) = v;
    }
  }
}

}  // namespace eb
