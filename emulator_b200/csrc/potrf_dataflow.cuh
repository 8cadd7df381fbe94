// Persistent dataflow Cholesky (sm_100a, fp64).
//
// One launch factorises the whole matrix.  The lower triangle is cut into 64x64 tiles; tile
// tasks are handed out in column-major order through an atomic counter, so every claimed
// task is owned by a CTA that is already running and depends only on lower-numbered tasks
// (no co-residency assumption, no deadlock).  The owner of tile (i, j) computes, left-looking,
//     C = K_ij - sum_{k<j} L_ik L_jk^T          (DMMA engine; K_ij is generated in-kernel from
//                                                 the training inputs: the kernel-matrix build
//                                                 is fused into the factorisation)
// streaming the finished row panels of tile rows i and j through the cp.async pipeline as soon
// as their per-tile ready flags are published, then finalises:
//     i == j :  L_jj = chol(C)  (register-blocked, one barrier per column)
//     i  > j :  L_ij = C L_jj^-T (row-parallel forward substitution against L_jj)
// writes the tile once, and publishes its flag (release store after a device-scope fence).
#pragma once
#include "kernels.cuh"

namespace eb {

using CfgP = TileCfg<64, 64, 32, 16>;  // 256 threads: 2 x 4 warps, 16 accumulator doubles / thread

struct PotrfArgs {
  double* A;
  int ld, nt, n;
  const double* X;        // [np][DP] padded training inputs
  const double* noise;    // [np]
  const KernelParams* kp;
  const int2* tiles;      // task list, column-major lower triangle
  int ntasks;
  unsigned* counter;      // task counter (zeroed before the launch)
  int* ready;             // [nt*nt] tile flags (zeroed before the launch)
  double* logdet_part;    // [nt]
  double* rdiag;          // [nt][64] reciprocal pivots 1 / L_jj, published with the diagonal tile
  int* info;
  int* abort_flag;        // watchdog: set when a wait exceeded its budget
  unsigned long long* trace;  // optional [ntasks][6] globaltimer stamps (diagnostics), or nullptr
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define EB_STAMP(slot)                                                            \
  do {                                                                            \
    if (a.trace != nullptr && tid == 0) a.trace[(size_t)task * 6 + (slot)] = gtime(); \
  } while (0)

// Flags are polled with a relaxed gpu-scope load: an acquire load would make ptxas emit
// CCTL.IVALL (L1 invalidate + drain of every outstanding load, including the cp.async ring) on
// every poll.  No L1 invalidation is needed because the tile data guarded by a flag is only ever
// read through L2 (cp.async.cg / ld.global.cg); the producer fences before its release store.
__device__ __forceinline__ int ld_flag(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Lane 0 of every warp polls; returns false if the watchdog fired.
__device__ __forceinline__ bool wait_ready(const int* flag, int* abort_flag) {
  bool ok = true;
  if ((threadIdx.x & 31) == 0) {
    long long spins = 0;
    while (ld_flag(flag) == 0) {
      __nanosleep(20);
      if ((++spins & 0x3ff) == 0) {
        if (spins > (1ll << 24) || ld_flag(abort_flag) != 0) {  // ~1 s
          atomicExch(abort_flag, 1);
          ok = false;
          break;
        }
      }
    }
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  return ok;
}

constexpr int PSTAGES = 4;  // cp.async ring depth of the dataflow kernel
constexpr int PS = 65;  // padded stride of the 64x64 finalisation buffers

template <int DP>
struct PotrfSmem {
  using E = Engine<CfgP, KCONTIG, KCONTIG>;
  static constexpr int PIPE = PSTAGES * E::STAGE_DOUBLES;
  static constexpr int C_OFF = PIPE;                 // [64][65]
  static constexpr int L_OFF = C_OFF + 64 * PS;      // [64][65]
  static constexpr int XI_OFF = L_OFF + 64 * PS;     // [DP][64]
  static constexpr int XJ_OFF = XI_OFF + DP * 64;    // [DP][64]
  static constexpr int COL_OFF = XJ_OFF + DP * 64;   // [2][64]
  static constexpr int RD_OFF = COL_OFF + 128;       // [64]
  static constexpr int DOUBLES = RD_OFF + 64 + 16;
  static constexpr int BYTES = DOUBLES * (int)sizeof(double);
};

template <int DP>
__global__ void __launch_bounds__(256, 1) potrf_dataflow_kernel(const PotrfArgs a) {
  extern __shared__ __align__(16) double gsm[];
  using SM = PotrfSmem<DP>;
  using E = typename SM::E;
  __shared__ unsigned s_task;
  __shared__ int s_ok;
  double* sC = gsm + SM::C_OFF;
  double* sL = gsm + SM::L_OFF;
  double* sXi = gsm + SM::XI_OFF;
  double* sXj = gsm + SM::XJ_OFF;
  double* colbuf = gsm + SM::COL_OFF;
  double* rdiag = gsm + SM::RD_OFF;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp / CfgP::WARPS_N) * CfgP::WM;
  const int wn0 = (warp % CfgP::WARPS_N) * CfgP::WN;
  const int g = lane >> 2, t = lane & 3;
  const int ld = a.ld, nt = a.nt;

  for (;;) {
    __syncthreads();  // previous task fully retired (smem reuse, s_task reuse)
    if (tid == 0) {
      s_task = atomicAdd(a.counter, 1u);
      s_ok = 1;
    }
    __syncthreads();
    const unsigned task = s_task;
    if (task >= (unsigned)a.ntasks) break;
    const int2 tile = a.tiles[task];
    const int ti = tile.x, tj = tile.y;
    const int i0 = ti * 64, j0 = tj * 64;
    bool ok = true;
    EB_STAMP(0);

    // stage the training inputs of the tile's rows / columns for the K_ij generation
    for (int e = tid; e < 64 * DP; e += 256) {
      int r = e / DP, m = e % DP;
      sXi[m * 64 + r] = a.X[(size_t)(i0 + r) * DP + m];
      sXj[m * 64 + r] = a.X[(size_t)(j0 + r) * DP + m];
    }

    __syncthreads();  // sXi / sXj staged

    // ---- acc = -K_ij, generated straight into the accumulator layout (before any panel wait,
    //      so the exp() work overlaps the wait for the producers) ----
    double acc[CfgP::MI][CfgP::NI][2];
    {
      const double amp = a.kp->amp;
      double r2[CfgP::MI][CfgP::NI * 2];
#pragma unroll
      for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
        for (int c = 0; c < CfgP::NI * 2; ++c) r2[mi][c] = 0.0;
#pragma unroll
      for (int m = 0; m < DP; ++m) {
        const double im = a.kp->inv_m[m];
        double xi[CfgP::MI], xj[CfgP::NI * 2];
#pragma unroll
        for (int mi = 0; mi < CfgP::MI; ++mi) xi[mi] = sXi[m * 64 + wm0 + mi * 8 + g];
#pragma unroll
        for (int c = 0; c < CfgP::NI * 2; ++c) xj[c] = sXj[m * 64 + wn0 + (c >> 1) * 8 + 2 * t + (c & 1)];
#pragma unroll
        for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
          for (int c = 0; c < CfgP::NI * 2; ++c) {
            const double dx = xi[mi] - xj[c];
            r2[mi][c] = fma(dx * dx, im, r2[mi][c]);
          }
      }
#pragma unroll
      for (int mi = 0; mi < CfgP::MI; ++mi) {
        const int gi = i0 + wm0 + mi * 8 + g;
#pragma unroll
        for (int c = 0; c < CfgP::NI * 2; ++c) {
          const int gj = j0 + wn0 + (c >> 1) * 8 + 2 * t + (c & 1);
          double kv;
          if (gi < a.n && gj < a.n) {
            kv = amp * exp(-0.5 * r2[mi][c]);
            if (gi == gj) kv += a.noise[gi];
          } else {
            kv = (gi == gj) ? 1.0 : 0.0;
          }
          acc[mi][c >> 1][c & 1] = -kv;
        }
      }
    }
    EB_STAMP(1);

    // ---- left-looking accumulation over the finished panels: acc += sum_k L_ik L_jk^T ----
    {
      const int nk = tj * (64 / KC);
      const double* Ag = a.A + (size_t)i0 * ld;
      const double* Bg = a.A + (size_t)j0 * ld;
      const int* ready_i = a.ready + (size_t)ti * nt;
      const int* ready_j = a.ready + (size_t)tj * nt;
      // leading tile columns whose panels are already published (one coalesced look per warp)
      int known = 0;
      for (int base = 0; base < tj; base += 32) {
        const int kc = base + lane;
        int f = 1;
        if (kc < tj) f = (ld_flag(ready_i + kc) != 0) && (ld_flag(ready_j + kc) != 0);
        const unsigned missing = __ballot_sync(0xffffffffu, f == 0);
        if (missing) {
          known = base + __ffs(missing) - 1;
          break;
        }
        known = min(base + 32, tj);
      }
      auto issue = [&](int kt) {
        if ((kt & 3) == 0) {
          const int kc = kt >> 2;
          if (kc >= known) {
            ok = wait_ready(ready_i + kc, a.abort_flag) && ok;
            if (ti != tj) ok = wait_ready(ready_j + kc, a.abort_flag) && ok;
          }
        }
        E::load_stage(gsm + (kt % PSTAGES) * E::STAGE_DOUBLES, Ag, ld, Bg, ld, kt, tid);
      };
#pragma unroll
      for (int s = 0; s < PSTAGES - 1; ++s) {
        if (s < nk) issue(s);
        cp_async_commit();
      }
      for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<PSTAGES - 2>();
        __syncthreads();
        const double* sA = gsm + (kt % PSTAGES) * E::STAGE_DOUBLES;
        const double* sB = sA + E::SA::DOUBLES;
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) {
          double af[CfgP::MI], bf[CfgP::NI];
#pragma unroll
          for (int mi = 0; mi < CfgP::MI; ++mi) af[mi] = sA[(wm0 + mi * 8 + g) * E::SA::STRIDE + kk + t];
#pragma unroll
          for (int ni = 0; ni < CfgP::NI; ++ni) bf[ni] = sB[(wn0 + ni * 8 + g) * E::SB::STRIDE + kk + t];
#pragma unroll
          for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
            for (int ni = 0; ni < CfgP::NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
        // refill the stage consumed in the previous iteration (every warp is past the barrier
        // above, so nobody still reads it); a flag wait here never delays math on loaded stages
        {
          const int nxt = kt + PSTAGES - 1;
          if (nxt < nk) issue(nxt);
          cp_async_commit();
        }
      }
      cp_async_wait<0>();
    }
    // ---- C = -acc parked in shared memory ----
#pragma unroll
    for (int mi = 0; mi < CfgP::MI; ++mi) {
      const int lr = wm0 + mi * 8 + g;
#pragma unroll
      for (int c = 0; c < CfgP::NI * 2; ++c) sC[lr * PS + wn0 + (c >> 1) * 8 + 2 * t + (c & 1)] = -acc[mi][c >> 1][c & 1];
    }
    __syncthreads();
    EB_STAMP(2);
    double* At = a.A + (size_t)i0 * ld + j0;
    // Finalisation layout: warp w owns columns 8w..8w+7 of the tile, lane owns rows lane, lane+32.
    double v[2][8];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
      for (int c = 0; c < 8; ++c) v[sl][c] = sC[(lane + 32 * sl) * PS + 8 * warp + c];
    if (ti == tj) {
      // ---- diagonal tile: right-looking Cholesky by 8-column panels; the owning warp factorises
      //      its panel warp-synchronously (shuffles only), the warps to its right apply it ----
      for (int p = 0; p < 8; ++p) {
        if (warp == p) {
          const int sj = p >> 2;            // row slot of this panel's pivot rows
          const int ob = 8 * (p & 3);       // lane of the first pivot row
          double dloc = sj ? v[1][0] : v[0][0];
          double d = __shfl_sync(0xffffffffu, dloc, ob);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (!(d > 0.0) && lane == 0) atomicCAS(a.info, 0, j0 + 8 * p + c + 1);
            const double rinv = rsqrt(d);
            if (lane == ob + c) rdiag[8 * p + c] = rinv;
            const double l0 = v[0][c] * rinv, l1 = v[1][c] * rinv;
            v[0][c] = l0;
            v[1][c] = l1;
            const double lsel = sj ? l1 : l0;
            if (c < 7) {
              // next pivot, computed by its own lane (keeps the second shuffle off the chain)
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
#pragma unroll
          for (int sl = 0; sl < 2; ++sl)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int row = lane + 32 * sl, col = 8 * p + c;
              sL[row * PS + col] = (row >= col) ? v[sl][c] : 0.0;
            }
        }
        __syncthreads();
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
      }
      // store L_jj (zero strict upper) and the reciprocal pivots
      for (int e = tid; e < 64 * 64; e += 256) {
        const int r = e >> 6, c = e & 63;
        At[(size_t)r * ld + c] = sL[r * PS + c];
      }
      if (tid < 64) a.rdiag[(size_t)tj * 64 + tid] = rdiag[tid];
    } else {
      // ---- off-diagonal tile: L_ij = C L_jj^-T, forward substitution by 8-column panels ----
      ok = wait_ready(a.ready + (size_t)tj * nt + tj, a.abort_flag) && ok;
      EB_STAMP(3);
      const double* Lg = a.A + (size_t)j0 * ld + j0;
      for (int e = tid; e < 64 * 32; e += 256) {
        const int r = e >> 5, c = (e & 31) * 2;
        const double2 val = __ldcg(reinterpret_cast<const double2*>(Lg + (size_t)r * ld + c));
        sL[r * PS + c] = val.x;
        sL[r * PS + c + 1] = val.y;
      }
      if (tid < 64) rdiag[tid] = __ldcg(a.rdiag + (size_t)tj * 64 + tid);
      __syncthreads();
      for (int p = 0; p < 8; ++p) {
        if (warp == p) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const double rd = rdiag[8 * p + c];
            const double x0 = v[0][c] * rd, x1 = v[1][c] * rd;
            v[0][c] = x0;
            v[1][c] = x1;
#pragma unroll
            for (int c2 = c + 1; c2 < 8; ++c2) {
              const double lv = sL[(8 * p + c2) * PS + 8 * p + c];
              v[0][c2] = fma(-x0, lv, v[0][c2]);
              v[1][c2] = fma(-x1, lv, v[1][c2]);
            }
          }
#pragma unroll
          for (int sl = 0; sl < 2; ++sl)
#pragma unroll
            for (int c = 0; c < 8; ++c) sC[(lane + 32 * sl) * PS + 8 * p + c] = v[sl][c];
        }
        __syncthreads();
        if (warp > p) {
          double ar[2][8];
#pragma unroll
          for (int sl = 0; sl < 2; ++sl)
#pragma unroll
            for (int k = 0; k < 8; ++k) ar[sl][k] = sC[(lane + 32 * sl) * PS + 8 * p + k];
#pragma unroll
          for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const double b = sL[(8 * warp + c) * PS + 8 * p + k];
              v[0][c] = fma(-ar[0][k], b, v[0][c]);
              v[1][c] = fma(-ar[1][k], b, v[1][c]);
            }
        }
      }
      for (int e = tid; e < 64 * 64; e += 256) {
        const int rr = e >> 6, c = e & 63;
        At[(size_t)rr * ld + c] = sC[rr * PS + c];
      }
    }
    // ---- publish ----
    EB_STAMP(4);
    const int all_ok = __syncthreads_and(ok ? 1 : 0);  // every thread's tile stores happen-before ...
    if (tid == 0) {
      __threadfence();                                  // ... this cumulative device-scope fence
      st_release(a.ready + (size_t)ti * nt + tj, 1);
    }
    EB_STAMP(5);
    if (ti == tj) {
      // log-determinant share, off the critical path: sum_j log L_jj = -sum_j log(1 / L_jj)
      if (tid < 64) colbuf[tid] = log(rdiag[tid]);
      __syncthreads();
      if (tid == 0) {
        double sum = 0.0;
        for (int j = 0; j < 64; ++j) sum += colbuf[j];
        a.logdet_part[tj] = -sum;
      }
    }
    if (!all_ok) break;  // watchdog fired somewhere in this CTA: leave (host reports the abort flag)
  }
}

// Batched triangular inverse of the 64x64 diagonal tiles (off the critical path): one CTA per
// tile, X = L_jj^-1 written to dinv[j].
constexpr int DIAG_INV_SMEM_BYTES = 2 * 64 * PS * (int)sizeof(double);
__global__ void __launch_bounds__(256) diag_inverse_kernel(const double* __restrict__ A, int ld,
                                                            double* __restrict__ dinv) {
  extern __shared__ __align__(16) double gsm[];
  double* L = gsm;
  double* X = gsm + 64 * PS;
  const int tj = blockIdx.x, tid = threadIdx.x;
  const double* At = A + (size_t)tj * 64 * ld + (size_t)tj * 64;
  for (int e = tid; e < 64 * 64; e += 256) {
    const int r = e >> 6, c = e & 63;
    L[r * PS + c] = At[(size_t)r * ld + c];
    X[r * PS + c] = 0.0;
  }
  __syncthreads();
  const int c = tid >> 2, q = tid & 3;
  for (int i = 0; i < 64; ++i) {
    double part = 0.0;
    for (int k = c + q; k < i; k += 4) part = fma(L[i * PS + k], X[k * PS + c], part);
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if (q == 0 && c <= i) {
      const double di = 1.0 / L[i * PS + i];
      X[i * PS + c] = (c == i) ? di : -part * di;
    }
    __syncthreads();
  }
  for (int e = tid; e < 64 * 64; e += 256) dinv[(size_t)tj * 64 * 64 + e] = X[(e >> 6) * PS + (e & 63)];
}

}  // namespace eb
