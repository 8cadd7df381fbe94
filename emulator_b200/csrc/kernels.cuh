// CUDA kernels of the GP hot path (fp64, sm_100a).  See DESIGN.md for the data layout.
//
// Matrix storage: two row-major (Np x Np) fp64 buffers per GP handle, Np = N rounded up to a
// multiple of 128; rows/columns >= N carry the identity so every kernel works on whole tiles.
// `A` receives the Cholesky factor L (factor_dataflow.cuh); `W` receives the inverse in TRANSPOSED form,
// Y = L^-T (upper triangle), which the gradient consumes as it is (K^-1 = Y Y^T); before the first variance
// prediction that follows an evaluation X = L^-1 is written row-major into `A` (L is dead by then).
#pragma once
#include <math.h>
#include "engine.cuh"
#include "tma_gemm.cuh"

namespace eb {

constexpr int MAXD = 16;   // maximum kernel dimension supported
constexpr int TB = 64;     // base (panel) tile edge
constexpr double TWO_PI_LOG = 1.8378770664093454835606594728112;  // log(2*pi)

// Per-evaluation kernel parameters derived on the device from the hyperparameter vector
// p = [log c, log M_0, ..., log M_{d-1}]  (george: `ConstantKernel * ExpSquaredKernel`,
// reference call site swiftemulator/emulators/gaussian_process.py:181-183).
struct KernelParams {
  double amp;           // d * exp(p0)   (george sums the constant kernel over its d axes)
  double inv_m[MAXD];   // 1 / exp(p_{i+1}); zero for padded dimensions
};

// Several independent problems of one shape can share launches (factor_dataflow.cuh; the stages after
// the factorisation below): block index -> problem through a pointer table passed by value.
constexpr int MAX_BATCH = 8;

// One block per problem of a batch: problem b reads p + b * p_stride and writes out[b].
__global__ void prep_params_kernel(const double* __restrict__ p, int d, KernelParams* __restrict__ out, int p_stride = 0) {
  if (threadIdx.x == 0) {
    p += (size_t)blockIdx.x * p_stride;
    out += blockIdx.x;
    out->amp = (double)d * exp(p[0]);
    for (int i = 0; i < MAXD; ++i) out->inv_m[i] = (i < d) ? 1.0 / exp(p[i + 1]) : 0.0;
  }
}

// The same for problems whose parameter blocks live in different handles.
struct PrepBatch {
  KernelParams* kp[MAX_BATCH];
};
__global__ void prep_params_group_kernel(const double* __restrict__ p, int p_stride, int d, const PrepBatch out) {
  if (threadIdx.x == 0) {
    p += (size_t)blockIdx.x * p_stride;
    KernelParams* o = out.kp[blockIdx.x];
    o->amp = (double)d * exp(p[0]);
    for (int i = 0; i < MAXD; ++i) o->inv_m[i] = (i < d) ? 1.0 / exp(p[i + 1]) : 0.0;
  }
}

// Per-problem pointers of the stages that follow the factorisation (alpha, K^-1 traces, final sum).
struct PostArgs {
  const double* W;        // X = L^-1
  const double* r;        // residual
  double* z;              // z = X r
  double* tpartial;       // [nT][np] partial sums of alpha = X^T z
  double* alpha;
  const double* Xtr;      // [np][DP] padded training inputs
  const KernelParams* kp;
  double* trace_partial;  // [n_lauum][DP+1]
  const double* logdet_part;
  const double* zfin;     // the z the final sum reads (value-only: the residual row of the factorised matrix)
  const int* info;
  const int* abort_flag;
  double* out;            // [d + 5] result vector
};
struct PostBatch {
  int nb;
  PostArgs p[MAX_BATCH];
};

// ---------------------------------------------------------------------------------------
// K1: kernel-matrix build.  One CTA per 64x64 tile of the lower triangle (tile list from the
// host); inputs staged in shared memory, stores are 128-byte coalesced row segments.
//   K_ij = amp * exp(-0.5 * sum_m (x_im - x_jm)^2 / M_m) + delta_ij * noise_i
// (george BasicSolver.compute; reference call site gaussian_process.py:203-206).
// Tiles listed with tj > ti are zero-filled (strictly-upper 64-blocks of diagonal 128-tiles,
// which later hold zeros of the triangular factors).
// ---------------------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(256) kbuild_kernel(const int2* __restrict__ tiles, const double* __restrict__ X,
                                                      const double* __restrict__ noise,
                                                      const KernelParams* __restrict__ kp, int n, int ld,
                                                      double* __restrict__ A) {
  __shared__ double sXi[DP][TB];
  __shared__ double sXj[DP][TB];
  __shared__ double sInv[DP];
  const int2 tile = tiles[blockIdx.x];
  const int i0 = tile.x * TB, j0 = tile.y * TB;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  if (tile.y > tile.x) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) A[(size_t)(i0 + ty + 16 * a) * ld + j0 + tx + 16 * b] = 0.0;
    return;
  }
  for (int e = tid; e < TB * DP; e += 256) {
    int r = e / DP, m = e % DP;
    sXi[m][r] = X[(size_t)(i0 + r) * DP + m];
    sXj[m][r] = X[(size_t)(j0 + r) * DP + m];
  }
  if (tid < DP) sInv[tid] = kp->inv_m[tid];
  __syncthreads();
  const double amp = kp->amp;
  double r2[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) r2[a][b] = 0.0;
#pragma unroll
  for (int m = 0; m < DP; ++m) {
    double xi[4], xj[4];
    const double im = sInv[m];
#pragma unroll
    for (int a = 0; a < 4; ++a) xi[a] = sXi[m][ty + 16 * a];
#pragma unroll
    for (int b = 0; b < 4; ++b) xj[b] = sXj[m][tx + 16 * b];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        double dx = xi[a] - xj[b];
        r2[a][b] = fma(dx * dx, im, r2[a][b]);
      }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int gi = i0 + ty + 16 * a;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int gj = j0 + tx + 16 * b;
      double v;
      if (gi < n && gj < n) {
        v = amp * exp(-0.5 * r2[a][b]);
        if (gi == gj) v += noise[gi];
      } else {
        v = (gi == gj) ? 1.0 : 0.0;
      }
      A[(size_t)gi * ld + gj] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Generic tile GEMM:  C_tile = beta * C_tile + alpha * A_slab * B_slab^T  over a task list.
// ---------------------------------------------------------------------------------------
template <class Cfg, int LA, int LB>
__global__ void __launch_bounds__(Cfg::THREADS) tile_gemm_kernel(const TileTask* __restrict__ tasks, double* C,
                                                                  int ldc, const double* A, int lda,
                                                                  const double* B, int ldb, double alpha,
                                                                  double beta) {
  extern __shared__ __align__(16) double gsm[];
  using E = Engine<Cfg, LA, LB>;
  const TileTask tk = tasks[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WM;
  const int wn0 = (warp % Cfg::WARPS_N) * Cfg::WN;
  // strictly-upper quadrant of a diagonal 128-tile: nothing to compute or store
  const bool active = !((tk.flags & 1) && (wm0 + Cfg::WM <= Cfg::BM / 2) && (wn0 >= Cfg::BN / 2));
  double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
  for (int mi = 0; mi < Cfg::MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < Cfg::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  E::mainloop(acc, A + (size_t)tk.a_row * lda + tk.a_col, lda, B + (size_t)tk.b_row * ldb + tk.b_col, ldb, tk.nk,
              gsm, active);
  if (!active) return;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mi = 0; mi < Cfg::MI; ++mi) {
    double* crow = C + (size_t)(tk.c_row + wm0 + mi * 8 + g) * ldc + tk.c_col + wn0 + 2 * t;
#pragma unroll
    for (int ni = 0; ni < Cfg::NI; ++ni) {
      double2* pc = reinterpret_cast<double2*>(crow + ni * 8);
      double2 v;
      if (beta != 0.0) {
        v = *pc;
        v.x = fma(alpha, acc[mi][ni][0], beta * v.x);
        v.y = fma(alpha, acc[mi][ni][1], beta * v.y);
      } else {
        v.x = alpha * acc[mi][ni][0];
        v.y = alpha * acc[mi][ni][1];
      }
      *pc = v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// The inverse is kept transposed: Y = X^T = L^-T, upper triangular, row-major (factor_dataflow.cuh).
// z = X r = Y^T r in two deterministic passes, stage 1: partial[rc][i] = sum_{j in row chunk rc, j <= i} Y[j][i] r[j].
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) trmv_t_partial_kernel(const PostBatch pb, int ld, int np) {
  const PostArgs& a = pb.p[blockIdx.z];
  const double* __restrict__ Y = a.W;
  const double* __restrict__ r = a.r;
  double* __restrict__ partial = a.tpartial;
  const int cb = blockIdx.x, rc = blockIdx.y;  // 128-column block (i), 128-row chunk (j)
  const int i = cb * 128 + threadIdx.x;
  double s = 0.0;
  if (rc <= cb) {
    const int j0 = rc * 128;
    __shared__ double sr[128];
    sr[threadIdx.x] = r[j0 + threadIdx.x];
    __syncthreads();
    const double* yp = Y + (size_t)j0 * ld + i;
#pragma unroll 8
    for (int jj = 0; jj < 128; ++jj) {
      const int j = j0 + jj;
      double v = (j <= i) ? yp[(size_t)jj * ld] : 0.0;
      s = fma(v, sr[jj], s);
    }
  }
  partial[(size_t)rc * np + i] = s;
}
// stage 2: z[i] = sum over row chunks in fixed order.
__global__ void __launch_bounds__(128) trmv_t_reduce_kernel(const PostBatch pb, int np) {
  const PostArgs& a = pb.p[blockIdx.y];
  const double* __restrict__ partial = a.tpartial;
  double* __restrict__ z = a.z;
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= np) return;
  double s = 0.0;
  for (int rc = 0; rc <= i / 128; ++rc) s += partial[(size_t)rc * np + i];
  z[i] = s;
}

// alpha = X^T z = Y z  (Y upper triangular, row-major): one warp per row, coalesced.
__global__ void __launch_bounds__(256) trmv_upper_kernel(const PostBatch pb, int ld, int np) {
  const PostArgs& a = pb.p[blockIdx.y];
  const double* __restrict__ Y = a.W;
  const double* __restrict__ z = a.z;
  double* __restrict__ alpha = a.alpha;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= np) return;
  const double* yr = Y + (size_t)row * ld;
  double s = 0.0;
  for (int i = (row & ~31) + lane; i < np; i += 32)
    if (i >= row) s = fma(yr[i], z[i], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) alpha[row] = s;
}

// X = Y^T written row-major into `out` (lower triangle; zeros above): what the predictive variance contracts with.
__global__ void __launch_bounds__(256) transpose_upper_kernel(const double* __restrict__ Y, int ld, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;  // block of `out`: rows by.., columns bx..
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (bx > by) {  // strictly upper block of X: zeros
    for (int r = ty; r < 32; r += 8) out[(size_t)(by + r) * ld + bx + tx] = 0.0;
    return;
  }
  for (int r = ty; r < 32; r += 8) tile[r][tx] = Y[(size_t)(bx + r) * ld + by + tx];  // Y rows bx.., columns by..
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int gi = by + r, gj = bx + tx;
    out[(size_t)gi * ld + gj] = (gj <= gi) ? tile[tx][r] : 0.0;
  }
}

// ---------------------------------------------------------------------------------------
// K3: K^-1 tile = sum_k Y[i][k] Y[j][k] (Y = L^-T; never written to HBM) fused with the gradient traces
//   g_0     = 1/2 sum_ij (a_i a_j - Kinv_ij) k_ij
//   g_{m+1} = 1/2 sum_ij (a_i a_j - Kinv_ij) k_ij * 1/2 (x_im - x_jm)^2 / M_m
// (george GP.grad_log_likelihood: A = alpha alpha^T - K^-1, einsum("ijk,ij", dK, A); reference
// call site gaussian_process.py:212-214).  dK is regenerated in registers; per-CTA partial sums
// go to `partial[task][DP+1]` and are combined in fixed order by finalize_kernel.
// The contraction runs through the TMA pipeline of tma_gemm.cuh (one producer warp, eight DMMA warps).
// If kinv_out != nullptr the lower-triangle tile is also stored (debug / tests).
// ---------------------------------------------------------------------------------------
constexpr int LAUUM_CS = 129;  // odd stride of the staged K^-1 tile
template <int DP>
struct LauumSmem {
  static constexpr int XS = DP | 1;  // odd stride: conflict-free row reads
  static constexpr int C_OFF = 0;
  static constexpr int XI_OFF = 128 * LAUUM_CS;
  static constexpr int XJT_OFF = XI_OFF + 128 * XS;
  static constexpr int AI_OFF = XJT_OFF + DP * 128;
  static constexpr int AJ_OFF = AI_OFF + 128;
  static constexpr int INV_OFF = AJ_OFF + 128;
  static constexpr int RED_OFF = INV_OFF + DP;
  static constexpr int DOUBLES = RED_OFF + 8 * (DP + 1);
  static constexpr int EPI_BYTES = DOUBLES * (int)sizeof(double);
  static constexpr int BYTES = EPI_BYTES > TMA_SMEM_BYTES ? EPI_BYTES : TMA_SMEM_BYTES;
};

// TMA views (128-row boxes) of the Y matrices of the problems of a batch.
struct LauumMaps {
  CUtensorMap y[MAX_BATCH];
};

template <int DP>
__global__ void __launch_bounds__(TMA_THREADS, 1) lauum_trace_kernel(const TileTask* __restrict__ tasks, const PostBatch pb,
                                                                     const __grid_constant__ LauumMaps maps, int ld, int n,
                                                                     double* kinv_out) {
  extern __shared__ __align__(16) double gsm[];
  using Cfg = Cfg128;
  using SM = LauumSmem<DP>;
  constexpr int XS = SM::XS;
  // block b works on task b / nb of problem b % nb: the task list is sorted longest first, so the blocks of
  // a batch still start in that order (and the many short tiles of all problems fill the tail together)
  const unsigned task_id = blockIdx.x / (unsigned)pb.nb;
  const unsigned prob = blockIdx.x % (unsigned)pb.nb;
  const PostArgs& a = pb.p[prob];
  const double* __restrict__ Xtr = a.Xtr;
  const double* __restrict__ alpha = a.alpha;
  const KernelParams* __restrict__ kp = a.kp;
  double* __restrict__ partial = a.trace_partial;
  const TileTask tk = tasks[task_id];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  TmaPipe pipe;
  pipe.carve(gsm);
  if (threadIdx.x == 0) pipe.init();
  __syncthreads();
  unsigned it = 0;
  if (warp == TMA_CONSUMER_WARPS) {
    if (lane == 0) tma_produce(pipe, &maps.y[prob], &maps.y[prob], tk.a_row, tk.a_col, tk.b_row, tk.b_col, tk.nk, it);
    return;
  }
  const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WM;
  const int wn0 = (warp % Cfg::WARPS_N) * Cfg::WN;
  const bool diag = tk.flags & 1;
  // bit 1: this K-slice carries the alpha_i alpha_j term (the trace is linear in K^-1, so a tile's
  // contraction may be split into several K-slices, each traced on its own)
  const double alpha_on = (tk.flags & 2) ? 1.0 : 0.0;
  const bool active = !(diag && (wm0 + Cfg::WM <= 64) && (wn0 >= 64));
  double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
  for (int mi = 0; mi < Cfg::MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < Cfg::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  tma_consume(pipe, acc, tk.nk, it, active);
  consumer_barrier();  // every warp is done with the ring: the epilogue buffers alias it

  // ---- epilogue: park the K^-1 tile in shared memory, stage the tile's inputs ----
  double* sC = gsm + SM::C_OFF;
  double* sXi = gsm + SM::XI_OFF;
  double* sXjT = gsm + SM::XJT_OFF;
  double* sAi = gsm + SM::AI_OFF;
  double* sAj = gsm + SM::AJ_OFF;
  double* sInv = gsm + SM::INV_OFF;
  double* sRed = gsm + SM::RED_OFF;
  const int i0 = tk.c_row, j0 = tk.c_col;
  {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mi = 0; mi < Cfg::MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < Cfg::NI; ++ni) {
        double* pc = sC + (wm0 + mi * 8 + g) * LAUUM_CS + wn0 + ni * 8 + 2 * t;
        pc[0] = acc[mi][ni][0];
        pc[1] = acc[mi][ni][1];
      }
  }
  for (int e = threadIdx.x; e < 128 * DP; e += 256) {
    int r = e / DP, m = e % DP;
    sXi[r * XS + m] = Xtr[(size_t)(i0 + r) * DP + m];
    sXjT[m * 128 + r] = Xtr[(size_t)(j0 + r) * DP + m];
  }
  if (threadIdx.x < 128) {
    sAi[threadIdx.x] = alpha[i0 + threadIdx.x];
    sAj[threadIdx.x] = alpha[j0 + threadIdx.x];
  }
  if (threadIdx.x < DP) sInv[threadIdx.x] = kp->inv_m[threadIdx.x];
  consumer_barrier();
  const double amp = kp->amp;
  double s[DP + 1];
#pragma unroll
  for (int m = 0; m <= DP; ++m) s[m] = 0.0;
  // warp w owns rows 16w..16w+15; lanes own columns lane + 32*cb
  double aj[4];
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) aj[cb] = sAj[lane + 32 * cb];
#pragma unroll 1
  for (int rr = 0; rr < 16; ++rr) {
    const int li = warp * 16 + rr;
    const int gi = i0 + li;
    if (gi >= n) break;
    const double ai = sAi[li];
    double r2[4];
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) r2[cb] = 0.0;
#pragma unroll
    for (int m = 0; m < DP; ++m) {
      const double xi = sXi[li * XS + m], im = sInv[m];
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        const double dx = xi - sXjT[m * 128 + lane + 32 * cb];
        r2[cb] = fma(dx * dx, im, r2[cb]);
      }
    }
    double wa[4];
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      const int lj = lane + 32 * cb;
      const int gj = j0 + lj;
      wa[cb] = 0.0;
      if (gj <= gi) {
        const double kinv = sC[li * LAUUM_CS + lj];
        if (kinv_out != nullptr) kinv_out[(size_t)gi * ld + gj] = kinv;
        const double w = (gi == gj) ? 1.0 : 2.0;
        wa[cb] = w * (alpha_on * ai * aj[cb] - kinv) * (amp * exp(-0.5 * r2[cb]));
      }
      s[0] += wa[cb];
    }
#pragma unroll
    for (int m = 0; m < DP; ++m) {
      const double xi = sXi[li * XS + m], hm = 0.5 * sInv[m];
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        const double dx = xi - sXjT[m * 128 + lane + 32 * cb];
        s[m + 1] = fma(wa[cb], dx * dx * hm, s[m + 1]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m <= DP; ++m) {
    double v = s[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sRed[warp * (DP + 1) + m] = v;
  }
  consumer_barrier();
  if (threadIdx.x <= DP) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += sRed[w * (DP + 1) + threadIdx.x];
    partial[(size_t)task_id * (DP + 1) + threadIdx.x] = v;
  }
}

// ---------------------------------------------------------------------------------------
// Final scalar assembly (one CTA):  out = [lml, grad_0..grad_{P-1}, info, logdet, quad].
//   lml = -1/2 (N log 2pi + 2 sum log L_jj) - 1/2 z^T z      (george GP.log_likelihood)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) finalize_kernel(const PostBatch pb, int nt, int n, int ntask, int dp, int d,
                                                        int want_grad) {
  const PostArgs& a = pb.p[blockIdx.x];
  const double* __restrict__ logdet_part = a.logdet_part;
  const double* __restrict__ z = a.zfin;
  const double* __restrict__ partial = a.trace_partial;
  const int* __restrict__ info = a.info;
  const int* __restrict__ abort_flag = a.abort_flag;
  double* __restrict__ out = a.out;
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s = fma(z[i], z[i], s);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double ld = 0.0;
    for (int t = 0; t < nt; ++t) ld += logdet_part[t];
    ld *= 2.0;
    const double quad = red[0];
    out[0] = -0.5 * ((double)n * TWO_PI_LOG + ld) - 0.5 * quad;
    out[1 + d + 1] = (*abort_flag != 0) ? -1.0 : (double)(*info);  // -1: device watchdog fired
    out[1 + d + 2] = ld;
    out[1 + d + 3] = quad;
  }
  // gradient: component c is summed by 8 threads over interleaved tasks, then combined in fixed
  // order (deterministic; 528 serial dependent loads by one thread took 30 us)
  __syncthreads();
  const int c = threadIdx.x >> 3, part = threadIdx.x & 7;
  double v = 0.0;
  if (want_grad && c <= d)
    for (int k = part; k < ntask; k += 8) v += partial[(size_t)k * (dp + 1) + c];
  red[threadIdx.x] = v;
  __syncthreads();
  if (want_grad && part == 0 && c <= d) {
    double sum = 0.0;
    for (int q = 0; q < 8; ++q) sum += red[threadIdx.x + q];
    out[1 + c] = 0.5 * sum;
  }
}

// The value-only reduction for a batch: block b finalises problem b (z rows live in different matrices).
struct ValueBatchPtrs {
  const double* z[8];
};
__global__ void __launch_bounds__(256) finalize_value_batch_kernel(const double* __restrict__ logdet_all, int nt,
                                                                    const ValueBatchPtrs zs, int n, int d,
                                                                    const int* __restrict__ info_all,
                                                                    double* __restrict__ out_all, int out_stride,
                                                                    int info_stride) {
  __shared__ double red[256];
  const int b = blockIdx.x;
  const double* z = zs.z[b];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s = fma(z[i], z[i], s);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double* out = out_all + (size_t)b * out_stride;
    double ld = 0.0;
    for (int t = 0; t < nt; ++t) ld += logdet_all[(size_t)b * nt + t];
    ld *= 2.0;
    const double quad = red[0];
    out[0] = -0.5 * ((double)n * TWO_PI_LOG + ld) - 0.5 * quad;  // same expression and order as finalize_kernel
    for (int i = 0; i <= d; ++i) out[1 + i] = 0.0;
    const int* info = info_all + (size_t)b * info_stride;  // {info, abort flag}
    out[1 + d + 1] = (info[1] != 0) ? -1.0 : (double)info[0];
    out[1 + d + 2] = ld;
    out[1 + d + 3] = quad;
  }
}

// ---------------------------------------------------------------------------------------
// K4a: prediction mean fused with the cross-covariance build.  One warp per 4 query rows;
// lanes stride over training points.  mu[q] = sum_j k(t_q, x_j) alpha_j; when ks != nullptr the
// K* rows are also written (row-major, ld = ldk) for the variance GEMM.
// (george GP.predict: Kxs = kernel.get_value(xs, x); mu = Kxs @ alpha; reference call site
// gaussian_process.py:285-287.)
// ---------------------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(256) predict_mean_kernel(const double* __restrict__ T, int m_rows,
                                                            const double* __restrict__ Xtr,
                                                            const double* __restrict__ alpha,
                                                            const KernelParams* __restrict__ kp, int n, int np,
                                                            double* __restrict__ mu, double* ks, int ldk) {
  constexpr int RQ = 4;
  __shared__ double sT[8][RQ][DP];
  __shared__ double sInv[DP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = (blockIdx.x * 8 + warp) * RQ;
  if (threadIdx.x < DP) sInv[threadIdx.x] = kp->inv_m[threadIdx.x];
  for (int e = lane; e < RQ * DP; e += 32) {
    int q = e / DP, m = e % DP;
    sT[warp][q][m] = (q0 + q < m_rows) ? T[(size_t)(q0 + q) * DP + m] : 0.0;
  }
  __syncthreads();
  if (q0 >= m_rows) {
    return;
  }
  const double amp = kp->amp;
  double accq[RQ];
#pragma unroll
  for (int q = 0; q < RQ; ++q) accq[q] = 0.0;
  for (int j = lane; j < np; j += 32) {
    double r2[RQ];
#pragma unroll
    for (int q = 0; q < RQ; ++q) r2[q] = 0.0;
    const double* xj = Xtr + (size_t)j * DP;
#pragma unroll
    for (int m = 0; m < DP; ++m) {
      const double x = xj[m], im = sInv[m];
#pragma unroll
      for (int q = 0; q < RQ; ++q) {
        const double dx = sT[warp][q][m] - x;
        r2[q] = fma(dx * dx, im, r2[q]);
      }
    }
    const double aj = (j < n) ? alpha[j] : 0.0;
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
      const double kv = (j < n) ? amp * exp(-0.5 * r2[q]) : 0.0;
      accq[q] = fma(kv, aj, accq[q]);
      if (ks != nullptr) ks[(size_t)(q0 + q) * ldk + j] = kv;  // rows padded to a tile multiple by the host
    }
  }
#pragma unroll
  for (int q = 0; q < RQ; ++q) {
    double v = accq[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && q0 + q < m_rows) mu[q0 + q] = v;
  }
}

// ---------------------------------------------------------------------------------------
// K4a for prediction grids (every parameter row theta_r at every independent value ind_b; query row
// q = r * n_ind + b is (ind_b, theta_r), the rows of mock_sweep / mock_hypercube / the Sobol sweep,
// mocking/__init__.py:84-95).  The squared-exponential kernel factorises over its axes,
//     k(q, x_j) = amp * exp(-(ind_b - x_j0)^2 / 2 M_0) * exp(-sum_{m>=1} (theta_rm - x_jm)^2 / 2 M_m)
//               = amp * E1[b][j] * E2[r][j],
// so a grid needs n_ind * N + n_theta * N exponentials instead of n_theta * n_ind * N: E1 is a small table
// (grid_axis_table_kernel), E2 is formed once per (r, j) in registers, and every K* entry is two multiplications.
// The kernel is then bound by the K* store (when the variance is wanted) instead of by exp().
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grid_axis_table_kernel(const double* __restrict__ ind, int n_ind,
                                                               const double* __restrict__ Xtr, int dp, int n, int np,
                                                               const KernelParams* __restrict__ kp, int round_f32,
                                                               double* __restrict__ e1) {
  const size_t e = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= (size_t)n_ind * np) return;
  const int b = (int)(e / np), j = (int)(e % np);
  double v = ind[b];
  if (round_f32) v = (double)(float)v;
  const double dx = v - Xtr[(size_t)j * dp];
  e1[e] = (j < n) ? exp(-0.5 * (dx * dx * kp->inv_m[0])) : 0.0;
}

template <int DP>
__global__ void __launch_bounds__(256) predict_grid_mean_kernel(const double* __restrict__ theta, size_t q0, int m_rows,
                                                                 int n_ind, int d, int round_f32,
                                                                 const double* __restrict__ e1,
                                                                 const double* __restrict__ Xtr,
                                                                 const double* __restrict__ alpha,
                                                                 const KernelParams* __restrict__ kp, int n, int np,
                                                                 double* __restrict__ mu, double* ks, int ldk,
                                                                 double* e2_out) {
  // (e2_out != nullptr: the per-row factor E2[r - r_first][j] = amp * exp(..) is kept for the fused variance kernel,
  //  var_grid_fused_kernel, which rebuilds K* = E1 * E2 tile by tile in shared memory; no K* row is stored then)
  // One CTA per parameter row; its 256 threads stride over the training points (coalesced E1 loads and K* stores),
  // NI independent values per pass: NI partial sums per thread, reduced over the CTA in fixed order at the end.
  constexpr int NI = 16;
  __shared__ double sTheta[DP];
  __shared__ double sInv[DP];
  __shared__ double sRed[8][NI];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t r = q0 / n_ind + blockIdx.x;  // parameter rows touched by this chunk: q0 / n_ind ..
  const size_t q_end = q0 + m_rows;
  if (threadIdx.x < DP) {
    sInv[threadIdx.x] = kp->inv_m[threadIdx.x];
    double v = 0.0;
    if (threadIdx.x >= 1 && threadIdx.x < d) {
      v = theta[r * (size_t)(d - 1) + (threadIdx.x - 1)];
      if (round_f32) v = (double)(float)v;
    }
    sTheta[threadIdx.x] = v;
  }
  __syncthreads();
  // the independent values of row r that fall into [q0, q_end)
  const size_t qr = r * n_ind;
  const int b_lo = (int)(qr < q0 ? q0 - qr : 0);
  const int b_hi = (int)(qr + n_ind > q_end ? q_end - qr : n_ind);
  const double amp = kp->amp;
  for (int b0 = b_lo; b0 < b_hi; b0 += NI) {
    const int nb = min(NI, b_hi - b0);
    double acc[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) acc[i] = 0.0;
    for (int j = threadIdx.x; j < np; j += 256) {
      double r2 = 0.0;
      const double* xj = Xtr + (size_t)j * DP;
#pragma unroll
      for (int m = 1; m < DP; ++m) {
        const double dx = sTheta[m] - xj[m];
        r2 = fma(dx * dx, sInv[m], r2);
      }
      const double e2 = (j < n) ? amp * exp(-0.5 * r2) : 0.0;
      if (e2_out != nullptr && b0 == b_lo) e2_out[(size_t)blockIdx.x * np + j] = e2;
      const double aj = (j < n) ? alpha[j] : 0.0;
      const double* e1p = e1 + (size_t)b0 * np + j;
      double* kp_out = (ks != nullptr) ? ks + (size_t)(qr + b0 - q0) * ldk + j : nullptr;
#pragma unroll
      for (int i = 0; i < NI; ++i)
        if (i < nb) {
          const double kv = e2 * e1p[(size_t)i * np];
          acc[i] = fma(kv, aj, acc[i]);
          if (kp_out != nullptr) kp_out[(size_t)i * ldk] = kv;
        }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      double v = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) sRed[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < nb) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += sRed[w][threadIdx.x];
      mu[qr + b0 + threadIdx.x - q0] = v;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// K4b: predictive variance.  V = K* X^T (X = L^-1 lower, so only k <= i contributes) through
// the tile engine; V is never stored: each CTA reduces its 128x128 tile to per-row sums of
// squares, rowss[itile][q].   var[q] = amp - sum_itile rowss[itile][q]
// (george: var = k(t,t) - sum(Kxs^T * K^-1 Kxs^T) = amp - |L^-1 k*|^2).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) var_gemm_kernel(const TileTask* __restrict__ tasks, const double* Ks, int ldk,
                                                        const double* X, int ld, double* __restrict__ rowss,
                                                        int ld_rowss) {
  extern __shared__ __align__(16) double gsm[];
  using Cfg = Cfg128;
  using E = Engine<Cfg, KCONTIG, KCONTIG>;
  const TileTask tk = tasks[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WM;
  double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
  for (int mi = 0; mi < Cfg::MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < Cfg::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  E::mainloop(acc, Ks + (size_t)tk.a_row * ldk + tk.a_col, ldk, X + (size_t)tk.b_row * ld + tk.b_col, ld, tk.nk, gsm,
              true);
  double* sRow = gsm;  // [4 warps_n][128 rows]
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
  __syncthreads();
  if (threadIdx.x < 128) {
    double v = sRow[threadIdx.x] + sRow[128 + threadIdx.x] + sRow[256 + threadIdx.x] + sRow[384 + threadIdx.x];
    rowss[(size_t)(tk.c_col / 128) * ld_rowss + tk.c_row + threadIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256) var_finish_kernel(const double* __restrict__ rowss, int ld_rowss, int nitile,
                                                          const KernelParams* __restrict__ kp, int m_rows,
                                                          double* __restrict__ var) {
  const int q = blockIdx.x * 256 + threadIdx.x;
  if (q >= m_rows) return;
  double s = 0.0;
  for (int it = 0; it < nitile; ++it) s += rowss[(size_t)it * ld_rowss + q];
  var[q] = kp->amp - s;
}

}  // namespace eb
