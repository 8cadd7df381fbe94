// fp64 tile-GEMM engine for sm_100a: cp.async (LDGSTS) multi-stage pipeline feeding
// DMMA.8x8x4 (mma.sync.m8n8k4.f64) -- the only FP64 tensor shape Blackwell has (tcgen05 has
// no f64 kind).  The dense contractions of the GP hot path are tile tasks on this engine:
// K^-1 = X^T X and the predictive variance through `Engine::mainloop` (128x128 tiles), the
// Cholesky / triangular-inverse tile products through its flag-synchronised 64x64 sibling in
// factor_dataflow.cuh.  (Measured: a 16-warp variant of the 128x128 tile and a deeper ring
// bring nothing; two chunks per barrier do: 31.4 -> 32.2 TFLOP/s, 89 % of cuBLAS DGEMM.)
//
// Operand layouts (both live in row-major global matrices):
//   KCONTIG : element (row, k) at base[row*ld + k]      ("N" for A, "T" for B)
//   MCONTIG : element (row, k) at base[k*ld + row]      ("T" for A, "N" for B)
// The k index inside a block of 8 is permuted between the two DMMAs that consume it: lane (g, t)
// feeds k = 2t to the first and k = 2t+1 to the second (any permutation is legal as long as A and B
// agree), so a KCONTIG operand needs ONE 16-byte shared load per two DMMAs.  Staging is padded so
// that those loads are bank-conflict free:
//   KCONTIG : [rows][KC+8]   -> lane address row*24 + 2t  (LDS.128: a quarter-warp covers 8 distinct 16B slots)
//   MCONTIG : [KC][rows+2]   -> lane address (2t | 2t+1)*(rows+2) + row  (LDS.64: a half-warp covers 16 distinct 8B slots)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace eb {

constexpr int KC = 16;      // k extent of one pipeline stage
constexpr int STAGES = 4;   // cp.async ring depth: two pairs of chunks

enum Layout : int { KCONTIG = 0, MCONTIG = 1 };

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// D(8x8) += A(8x4) * B(4x8).  Lane l holds A[l/4][l%4], B[l%4][l/4], C[l/4][2*(l%4)+{0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int ROWS, int LAYOUT, bool PERMK = true>
struct OperandSmem;

// Fragments of N 8-row blocks for two consecutive DMMAs: f[0][i] = op(row0 + 8i, k0), f[1][i] = op(row0 + 8i, k0 + 1).
template <int LAYOUT, int N>
__device__ __forceinline__ void load_fragments(double (&f)[2][N], const double* s, int stride, int row0, int k0) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (LAYOUT == 0) {  // KCONTIG
      const double2 v = *reinterpret_cast<const double2*>(s + (row0 + 8 * i) * stride + k0);
      f[0][i] = v.x;
      f[1][i] = v.y;
    } else {
      f[0][i] = s[k0 * stride + row0 + 8 * i];
      f[1][i] = s[(k0 + 1) * stride + row0 + 8 * i];
    }
  }
}

// PERMK = false keeps the plain k order (lane t feeds k = t): strides [rows][KC+4] / [KC][rows+4], one
// 8-byte load per DMMA operand.  Measured on B200: the permutation gains 12 % on the 64x64 KCONTIG loop
// (fewer shared loads per DMMA), nothing on the 128x128 KCONTIG tile, and costs the MCONTIG x MCONTIG
// tile of the K^-1 pass 4 % (no 16-byte loads to gain, coarser interleaving), which therefore opts out.
template <int ROWS, int LAYOUT, bool PERMK>
struct OperandSmem {
  static constexpr int STRIDE = PERMK ? ((LAYOUT == KCONTIG) ? (KC + 8) : (ROWS + 2)) : ((LAYOUT == KCONTIG) ? (KC + 4) : (ROWS + 4));
  static constexpr int DOUBLES = (LAYOUT == KCONTIG) ? ROWS * STRIDE : KC * STRIDE;
};

// One CTA tile configuration.
template <int BM_, int BN_, int WM_, int WN_>
struct TileCfg {
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_;
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int THREADS = 32 * WARPS_M * WARPS_N;
  static constexpr int MI = WM / 8, NI = WN / 8;
};
using Cfg128 = TileCfg<128, 128, 64, 32>;  // 256 threads, 64 accumulator doubles / thread

template <class Cfg, int LA, int LB, bool PERMK = true>
struct Engine {
  using SA = OperandSmem<Cfg::BM, LA, PERMK>;
  using SB = OperandSmem<Cfg::BN, LB, PERMK>;
  static constexpr int STAGE_DOUBLES = SA::DOUBLES + SB::DOUBLES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * (int)sizeof(double);

  // Stage one ROWS x KC operand slab into shared memory with 16-byte cp.async.
  template <int ROWS, int LAYOUT>
  static __device__ __forceinline__ void load_operand(double* s, const double* g, int ld, int tid) {
    constexpr int STRIDE = OperandSmem<ROWS, LAYOUT, PERMK>::STRIDE;
    if (LAYOUT == KCONTIG) {
      constexpr int CHUNKS = ROWS * (KC / 2);
#pragma unroll
      for (int c = tid; c < CHUNKS; c += Cfg::THREADS) {
        int row = c >> 3, kc = (c & 7) * 2;
        cp_async16(s + row * STRIDE + kc, g + (size_t)row * ld + kc);
      }
    } else {
      constexpr int CHUNKS = KC * (ROWS / 2);
#pragma unroll
      for (int c = tid; c < CHUNKS; c += Cfg::THREADS) {
        int k = c / (ROWS / 2), r2 = (c % (ROWS / 2)) * 2;
        cp_async16(s + k * STRIDE + r2, g + (size_t)k * ld + r2);
      }
    }
  }

  static __device__ __forceinline__ void load_stage(double* stage, const double* Ag, int lda,
                                                    const double* Bg, int ldb, int kt, int tid) {
    const double* a = (LA == KCONTIG) ? Ag + (size_t)kt * KC : Ag + (size_t)kt * KC * lda;
    const double* b = (LB == KCONTIG) ? Bg + (size_t)kt * KC : Bg + (size_t)kt * KC * ldb;
    load_operand<Cfg::BM, LA>(stage, a, lda, tid);
    load_operand<Cfg::BN, LB>(stage + SA::DOUBLES, b, ldb, tid);
  }

  // acc[mi][ni][2] += A(BM x nk*KC) * B(BN x nk*KC)^T for this warp's WM x WN sub-tile.
  // `active` lets whole warps skip the math (used for the strictly-upper quadrant of
  // diagonal tiles) while still taking part in loads and barriers.
  static __device__ __forceinline__ void mainloop(double (&acc)[Cfg::MI][Cfg::NI][2], const double* Ag,
                                                  int lda, const double* Bg, int ldb, int nk,
                                                  double* smem, bool active) {
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WM;
    const int wn0 = (warp % Cfg::WARPS_N) * Cfg::WN;
    const int g = lane >> 2, t = lane & 3;

    // Two chunks ("a pair") per barrier, nk even.  While pair kt is in the math, pair kt+2 is in
    // flight into the other two stages; it is issued right after the barrier that retires pair kt-2.
    static_assert(STAGES == 4, "ring logic assumes two pairs");
    if (nk > 0) {
      load_stage(smem, Ag, lda, Bg, ldb, 0, tid);
      load_stage(smem + STAGE_DOUBLES, Ag, lda, Bg, ldb, 1, tid);
    }
    cp_async_commit();
    for (int kt = 0; kt < nk; kt += 2) {
      cp_async_wait<0>();
      __syncthreads();
      if (kt + 2 < nk) {
        load_stage(smem + ((kt + 2) % STAGES) * STAGE_DOUBLES, Ag, lda, Bg, ldb, kt + 2, tid);
        load_stage(smem + ((kt + 3) % STAGES) * STAGE_DOUBLES, Ag, lda, Bg, ldb, kt + 3, tid);
      }
      cp_async_commit();
      if (active) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double* sA = smem + ((kt + h) % STAGES) * STAGE_DOUBLES;
          const double* sB = sA + SA::DOUBLES;
          if (PERMK) {
#pragma unroll
            for (int kk = 0; kk < KC; kk += 8) {
              double af[2][Cfg::MI], bf[2][Cfg::NI];
              load_fragments<LA, Cfg::MI>(af, sA, SA::STRIDE, wm0 + g, kk + 2 * t);
              load_fragments<LB, Cfg::NI>(bf, sB, SB::STRIDE, wn0 + g, kk + 2 * t);
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
                for (int mi = 0; mi < Cfg::MI; ++mi)
#pragma unroll
                  for (int ni = 0; ni < Cfg::NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[h2][mi], bf[h2][ni]);
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < KC; kk += 4) {
              double af[Cfg::MI], bf[Cfg::NI];
#pragma unroll
              for (int mi = 0; mi < Cfg::MI; ++mi) {
                const int row = wm0 + mi * 8 + g;
                af[mi] = (LA == KCONTIG) ? sA[row * SA::STRIDE + kk + t] : sA[(kk + t) * SA::STRIDE + row];
              }
#pragma unroll
              for (int ni = 0; ni < Cfg::NI; ++ni) {
                const int col = wn0 + ni * 8 + g;
                bf[ni] = (LB == KCONTIG) ? sB[col * SB::STRIDE + kk + t] : sB[(kk + t) * SB::STRIDE + col];
              }
#pragma unroll
              for (int mi = 0; mi < Cfg::MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < Cfg::NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
            }
          }
        }
      }
    }
    cp_async_wait<0>();
    __syncthreads();  // pipeline smem may be reused by the caller's epilogue
  }
};

// A tile task: C tile origin plus operand origins (memory row / memory column of the first
// element of the operand slab; for MCONTIG operands the memory row is the k index).
struct TileTask {
  int c_row, c_col;
  int a_row, a_col;
  int b_row, b_col;
  int nk;     // number of KC-chunks in the contraction
  int flags;  // bit0: diagonal 128-tile (skip the strictly-upper 64x64 quadrant)
};

}  // namespace eb
