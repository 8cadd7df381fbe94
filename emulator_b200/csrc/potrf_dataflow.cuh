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
  int* info;
  int* abort_flag;        // watchdog: set when a wait exceeded its budget
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
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
    while (ld_acquire(flag) == 0) {
      __nanosleep(40);
      if ((++spins & 0x3ff) == 0) {
        if (spins > (1ll << 24) || ld_acquire(abort_flag) != 0) {  // ~1 s
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

constexpr int PS = 65;  // padded stride of the 64x64 finalisation buffers

template <int DP>
struct PotrfSmem {
  using E = Engine<CfgP, KCONTIG, KCONTIG>;
  static constexpr int PIPE = STAGES * E::STAGE_DOUBLES;
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

    // stage the training inputs of the tile's rows / columns for the K_ij generation
    for (int e = tid; e < 64 * DP; e += 256) {
      int r = e / DP, m = e % DP;
      sXi[m * 64 + r] = a.X[(size_t)(i0 + r) * DP + m];
      sXj[m * 64 + r] = a.X[(size_t)(j0 + r) * DP + m];
    }

    // ---- left-looking accumulation over the finished panels: acc = sum_k L_ik L_jk^T ----
    double acc[CfgP::MI][CfgP::NI][2];
#pragma unroll
    for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < CfgP::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    {
      const int nk = tj * (64 / KC);
      const double* Ag = a.A + (size_t)i0 * ld;
      const double* Bg = a.A + (size_t)j0 * ld;
      const int* ready_i = a.ready + (size_t)ti * nt;
      const int* ready_j = a.ready + (size_t)tj * nt;
      auto issue = [&](int kt) {
        if ((kt & 3) == 0) {
          const int kc = kt >> 2;
          ok = wait_ready(ready_i + kc, a.abort_flag) && ok;
          if (ti != tj) ok = wait_ready(ready_j + kc, a.abort_flag) && ok;
        }
        E::load_stage(gsm + (kt % STAGES) * E::STAGE_DOUBLES, Ag, ld, Bg, ld, kt, tid);
      };
#pragma unroll
      for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) issue(s);
        cp_async_commit();
      }
      for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
          const int nxt = kt + STAGES - 1;
          if (nxt < nk) issue(nxt);
          cp_async_commit();
        }
        const double* sA = gsm + (kt % STAGES) * E::STAGE_DOUBLES;
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
      }
      cp_async_wait<0>();
    }
    __syncthreads();  // sXi / sXj staged (also when nk == 0)

    // ---- C = K_ij - acc, generated in the accumulator layout and parked in shared memory ----
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
        const int lr = wm0 + mi * 8 + g, gi = i0 + lr;
#pragma unroll
        for (int c = 0; c < CfgP::NI * 2; ++c) {
          const int lc = wn0 + (c >> 1) * 8 + 2 * t + (c & 1), gj = j0 + lc;
          double kv;
          if (gi < a.n && gj < a.n) {
            kv = amp * exp(-0.5 * r2[mi][c]);
            if (gi == gj) kv += a.noise[gi];
          } else {
            kv = (gi == gj) ? 1.0 : 0.0;
          }
          sC[lr * PS + lc] = kv - acc[mi][c >> 1][c & 1];
        }
      }
    }
    __syncthreads();

    double* At = a.A + (size_t)i0 * ld + j0;
    if (ti == tj) {
      // ---- diagonal tile: Cholesky, every thread keeps a 4x4 strided block in registers ----
      const int ty = tid >> 4, tx = tid & 15;
      double v[4][4];
#pragma unroll
      for (int aa = 0; aa < 4; ++aa)
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) v[aa][bb] = sC[(ty + 16 * aa) * PS + tx + 16 * bb];
      if (tx == 0) {
#pragma unroll
        for (int aa = 0; aa < 4; ++aa) colbuf[ty + 16 * aa] = v[aa][0];
      }
      __syncthreads();
      for (int j = 0; j < 64; ++j) {
        const double* col = colbuf + (j & 1) * 64;
        const double d = col[j];
        if (!(d > 0.0) && tid == 0) atomicCAS(a.info, 0, j0 + j + 1);
        const double rinv = rsqrt(d);
        double li[4], lk[4];
#pragma unroll
        for (int aa = 0; aa < 4; ++aa) li[aa] = col[ty + 16 * aa] * rinv;
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) lk[bb] = col[tx + 16 * bb] * rinv;
#pragma unroll
        for (int aa = 0; aa < 4; ++aa)
#pragma unroll
          for (int bb = 0; bb < 4; ++bb) {
            const int row = ty + 16 * aa, cc = tx + 16 * bb;
            if (cc > j && cc <= row) v[aa][bb] = fma(-li[aa], lk[bb], v[aa][bb]);
          }
        if (tid < 64) sL[tid * PS + j] = (tid > j) ? col[tid] * rinv : ((tid == j) ? d * rinv : 0.0);
        // owners of column j+1 publish it for the next step
        if (j + 1 < 64 && tx == ((j + 1) & 15)) {
          const int bb = (j + 1) >> 4;
          double* nxt = colbuf + ((j + 1) & 1) * 64;
#pragma unroll
          for (int aa = 0; aa < 4; ++aa) {
            double val = (bb == 0) ? v[aa][0] : (bb == 1) ? v[aa][1] : (bb == 2) ? v[aa][2] : v[aa][3];
            nxt[ty + 16 * aa] = val;
          }
        }
        __syncthreads();
      }
      // store L_jj (zero strict upper), log-determinant share
      for (int e = tid; e < 64 * 64; e += 256) {
        const int r = e >> 6, c = e & 63;
        At[(size_t)r * ld + c] = (c <= r) ? sL[r * PS + c] : 0.0;
      }
      if (tid == 0) {
        double s = 0.0;
        for (int j = 0; j < 64; ++j) s += log(sL[j * PS + j]);
        a.logdet_part[tj] = s;
      }
    } else {
      // ---- off-diagonal tile: L_ij = C L_jj^-T by forward substitution, 4 lanes per row ----
      ok = wait_ready(a.ready + (size_t)tj * nt + tj, a.abort_flag) && ok;
      const double* Lg = a.A + (size_t)j0 * ld + j0;
      for (int e = tid; e < 64 * 64; e += 256) {
        const int r = e >> 6, c = e & 63;
        sL[r * PS + c] = __ldcg(Lg + (size_t)r * ld + c);
      }
      __syncthreads();
      if (tid < 64) rdiag[tid] = 1.0 / sL[tid * PS + tid];
      __syncthreads();
      const int r = tid >> 2, q = tid & 3;
      double x[16];
#pragma unroll
      for (int m = 0; m < 16; ++m) x[m] = sC[r * PS + q + 4 * m];
      const int src_base = lane & ~3;
#pragma unroll
      for (int k = 0; k < 64; ++k) {
        const int qk = k & 3, mk = k >> 2;
        double xk = (q == qk) ? x[mk] * rdiag[k] : 0.0;
        xk = __shfl_sync(0xffffffffu, xk, src_base + qk);
        if (q == qk) x[mk] = xk;
#pragma unroll
        for (int m = mk; m < 16; ++m) {
          const int col = q + 4 * m;
          if (col > k) x[m] = fma(-xk, sL[col * PS + k], x[m]);
        }
      }
#pragma unroll
      for (int m = 0; m < 16; ++m) sC[r * PS + q + 4 * m] = x[m];
      __syncthreads();
      for (int e = tid; e < 64 * 64; e += 256) {
        const int rr = e >> 6, c = e & 63;
        At[(size_t)rr * ld + c] = sC[rr * PS + c];
      }
    }
    // ---- publish ----
    __threadfence();
    const int all_ok = __syncthreads_and(ok ? 1 : 0);
    if (tid == 0) st_release(a.ready + (size_t)ti * nt + tj, 1);
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
