// Persistent dataflow factorisation (sm_100a, fp64): Cholesky K = L L^T and the triangular
// inverse X = L^-1 in ONE launch.
//
// The lower triangle is cut into 64x64 tiles.  Static task lists (built on the host) are handed out to
// the CTAs of a persistent launch: the Cholesky's queues in order through atomic counters, the
// inverse's queues ready-first with a ticket fallback over a dependency-safe order (claim_inverse).
// Every task a CTA may block in has all its predecessors claimed by CTAs that are already running:
// no co-residency assumption, no deadlock.  Task kinds:
//
//   POTRF (i, j), column-major:  C = K_ij - sum_{k<j} L_ik L_jk^T   (K_ij generated in-kernel:
//        the kernel-matrix build is fused into the factorisation), then
//          i == j : L_jj = chol(C) (8-column panels, warp-synchronous), streamed out panel by panel;
//                   X_jj = L_jj^-1 is the separate task XDIAG that follows those panels
//          i  > j : L_ij = C L_jj^-T: panelled forward substitution against the streamed panels of L_jj
//                   for the near tiles i <= j + NEAR_BAND, one tile product with X_jj for all others
//          i == nt (value-only evaluations): the residual row, see the kernel
//   TRTRI (r, J), J < r, listed one block column behind the Cholesky front:
//          X_rJ = -X_rr * sum_{K=J}^{r-1} L_rK X_KJ, computed and stored in TRANSPOSED form: the second matrix
//          holds Y = X^T = L^-T (upper triangle), Y_Jr = -(sum_K Y_JK L_rK^T) X_rr^T.  In this form every
//          contraction of the evaluation -- Cholesky, inverse, K^-1 = Y Y^T -- runs over the contiguous index of
//          BOTH operands (16-byte fragment loads, one TMA box shape), and the tile is stored as it is accumulated.
//
// The Cholesky has a 64-step dependency chain (diagonal tile -> tile below -> next diagonal
// tile) that leaves most SMs idle; the inverse's tile products fill them.  Producers publish a
// per-tile flag (CTA barrier, one cumulative device-scope fence, release store); consumers poll
// with relaxed loads and read tile data only through L2 (cp.async.cg / ld.global.cg), so no L1
// invalidation is ever needed.  A watchdog in the poll loop turns a lost dependency into an
// error code instead of a hang.
#pragma once
#include "kernels.cuh"
#include "tile_ops.cuh"

namespace eb {

using CfgP = TileCfg<64, 64, 32, 16>;  // 256 threads: 2 x 4 warps, 16 accumulator doubles / thread
constexpr int PSTAGES = 4;             // cp.async ring depth (two CTAs per SM share 227 KB)
constexpr int KS = KC + 8;             // stride of a staged KCONTIG slab (engine.cuh)
constexpr int PSTAGE_DOUBLES = 2 * 64 * KS;
constexpr int PS = TPS;                // stride of the 64x64 finalisation buffers (tile_ops.cuh)
constexpr int TS = 66;                 // stride of a parked accumulator used as a [k][n] operand (conflict-free with permuted k)
constexpr int TSK = 72;                // stride of a parked accumulator used as a [m][k] operand (LDS.128 fragments)
constexpr int NEAR_BAND = 2;           // tiles (j+1..j+NEAR_BAND, j) cannot wait for X_jj: they follow L_jj's panels

// TASK_TRTRI_PART (r, J, [k0, k)): helper that adds sum_{K=k0}^{k-1} L_rK X_KJ of a long inverse tile to
// the partial sum kept in the (unused) mirrored upper tile of X.  A tile's helpers form a chain of
// chunks, each claimable as soon as row k-1 of X exists, so the bulk of the last rows of the inverse
// -- the tail of the launch -- is done while the Cholesky chain leaves the SMs idle, and the tile's
// own task (k = first tile column it still has to add) is short.
// TASK_POTRF_PART (i, j, k): the same idea for the Cholesky: sum_{c<k} L_ic L_jc^T of tile (i, j) is
// pre-accumulated by a background task listed several block columns ahead and parked in the tile's
// own (not yet published) storage, so that the tile's task -- which near the diagonal sits on or
// next to the dependency chain -- has a bounded amount of work however late it is claimed.
// TASK_XDIAG (j, j): X_jj = L_jj^-1 by substitution on the identity, following the panels of L_jj as the
// diagonal tile's owner streams them out, so that the inverse is published a few microseconds after
// the factor itself: every tile of column j beyond the near band finishes with one product by X_jj^T.
enum FactorTaskType : int { TASK_POTRF = 0, TASK_TRTRI = 1, TASK_TRTRI_PART = 2, TASK_POTRF_PART = 3, TASK_XDIAG = 4 };
struct FactorTask {
  int type, i, j, k;
  int seq;  // position in the dependency-safe total order of the inverse's tasks (ready-first queues)
  int k0;   // helpers of the inverse: first tile column of this chunk (the chunk adds [k0, k))
};

struct FactorArgs {
  double* A;   // K -> L (lower triangle)
  double* W;   // Y = X^T = L^-T (upper triangle): the inverse is kept transposed
  double* Xd;  // compact [nt][64][64] buffer of the diagonal tiles X_jj = L_jj^-1 (xd_tile = 64 * 64, xd_ld = 64)
  long long xd_tile;
  int xd_ld;
  int ld, nt, n;
  int value_only;         // 1: Cholesky only, plus the residual carried as tile row nt (see TASK_POTRF); no inverse
  const double* r;        // [np] residual (value_only)
  const double* X;        // [np][DP] padded training inputs
  const double* noise;    // [np]
  const KernelParams* kp;
  const FactorTask* tasks;  // four queues back to back: chain tiles | other Cholesky tasks | inverse tiles | inverse helpers
  int ntasks;
  int qbase[5];             // queue q holds tasks [qbase[q], qbase[q+1])
  int* claimed;             // [ntasks] claim marks of the ready-first queues 2 and 3 (zeroed before the launch)
  const int* inv_order;     // [n_inv] the inverse's tasks in their dependency-safe order (counter[4] walks it)
  int n_inv;
  int n_sms;                // CTAs are numbered in two waves of n_sms
  int n_chain_sms;          // CTAs with (blockIdx % n_sms) < n_chain_sms serve the chain queue
  unsigned* counter;        // per queue: next task (queues 0, 1) / first task not known to be claimed (queues 2, 3)
  // flag arrays are [(nt+1)*nt], indexed tile_row * nt + tile_col (row nt: the residual row of value_only)
  int* ready;             // L tile flags (zeroed before the launch)
  int* xready;            // X tile flags (zeroed before the launch)
  int* pready;            // partial-sum flags of long inverse tiles (zeroed before the launch)
  int* cready;            // partial-sum flags of the Cholesky tiles (zeroed before the launch)
  int* lready;            // [nt*8] per-panel flags of the diagonal tiles (zeroed before the launch)
  int* sm_chain;          // [256] per-SM count of resident CTAs working on a chain tile (zeroed before the launch)
  double* logdet_part;    // [nt]
  double* rdiag;          // [nt][64] reciprocal pivots 1 / L_jj, published with the diagonal tile
  int* info;
  int* abort_flag;        // watchdog: set when a wait exceeded its budget
  unsigned long long* trace;  // optional [ntasks][TRACE_SLOTS] (diagnostics), or nullptr
};

// Several independent factorisations of the same shape can share one launch: a value-only evaluation,
// and any evaluation of a small problem, is bound by its dependency chain and leaves most SMs idle, so
// the chains of up to MAX_BATCH problems are interleaved (the ensemble sampler proposes half its walkers
// at once; lock-step optimisers evaluate many equally shaped GPs per round).  CTA c is at home in
// problem c % nb and helps the others.
struct FactorBatch {
  int nb;
  FactorArgs p[MAX_BATCH];
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr int TRACE_SLOTS = 10;  // 6 stamps, SM id, CTA id, ns waited for flags, ns yielded to the chain
#define EB_STAMP(slot)                                                                \
  do {                                                                                \
    if (a.trace != nullptr && tid == 0) {                                             \
      a.trace[(size_t)task * TRACE_SLOTS + (slot)] = gtime();                          \
      if ((slot) == 5) {                                                              \
        a.trace[(size_t)task * TRACE_SLOTS + 8] = s_wait[0];                           \
        a.trace[(size_t)task * TRACE_SLOTS + 9] = s_wait[1];                           \
      }                                                                               \
    }                                                                                 \
  } while (0)

// Flags are polled with a relaxed gpu-scope load: an acquire load would make ptxas emit
// CCTL.IVALL (L1 invalidate + drain of every outstanding load, including the cp.async ring) on
// every poll.  INVARIANT this relies on: tile data published under a flag is only ever read with instructions
// that bypass L1 -- cp.async.cg, ld.global.cg (__ldcg) -- so there is no stale L1 line to invalidate and the producer's
// __threadfence + release store make the data visible in L2 before the flag.  A plain ld.global of tile data
// anywhere in this kernel would break it (the parity suite at ragged sizes is what would notice).
__device__ __forceinline__ int ld_flag(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned smid() {
  unsigned v;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Lane 0 of every warp polls; returns false if the watchdog fired.
// The watchdog's budget is TIME (globaltimer), not a spin count: a CTA may legitimately wait long when several
// persistent launches share the device (refits from host threads, overlapping graphs, a profiler's replay).
constexpr unsigned long long WATCHDOG_NS = 20ull * 1000ull * 1000ull * 1000ull;  // 20 s without progress = a lost dependency
__device__ __forceinline__ bool watchdog_expired(unsigned long long& t0, const int* abort_flag) {
  if (ld_flag(abort_flag) != 0) return true;  // somebody else gave up: follow
  const unsigned long long now = gtime();
  if (t0 == 0ull) t0 = now;  // (first checkpoint, ~1024 polls in: the clock is only read by waits that are already long)
  return now - t0 > WATCHDOG_NS;
}

__device__ __forceinline__ bool wait_ready(const int* flag, int* abort_flag, int need = 1) {
  bool ok = true;
  if ((threadIdx.x & 31) == 0) {
    unsigned spins = 0;
    unsigned long long t0 = 0ull;
    while (ld_flag(flag) < need) {
      __nanosleep(20);
      if ((++spins & 0x3ff) == 0 && watchdog_expired(t0, abort_flag)) {
        atomicExch(abort_flag, 1);
        ok = false;
        break;
      }
    }
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  return ok;
}

// The same for two flags at once (f2 may be null): one L2 round trip per poll instead of two.
__device__ __forceinline__ bool wait_ready2(const int* f1, const int* f2, int* abort_flag) {
  bool ok = true;
  if ((threadIdx.x & 31) == 0) {
    unsigned spins = 0;
    unsigned long long t0 = 0ull;
    for (;;) {
      const int v1 = ld_flag(f1);
      const int v2 = (f2 != nullptr) ? ld_flag(f2) : 1;
      if (v1 != 0 && v2 != 0) break;
      __nanosleep(20);
      if ((++spins & 0x3ff) == 0 && watchdog_expired(t0, abort_flag)) {
        atomicExch(abort_flag, 1);
        ok = false;
        break;
      }
    }
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  return ok;
}

// Background work stands aside while a chain tile is being finalised on this SM.
__device__ __forceinline__ void yield_to_chain(const int* sm_flag) {
  if ((threadIdx.x & 31) == 0) {
    int spins = 0;
    while (ld_flag(sm_flag) != 0 && ++spins < 256) __nanosleep(200);  // a chain finalisation lasts ~10 us; never wait much longer
  }
  __syncwarp();
}

// Publish a tile: all of the CTA's stores happen-before the barrier, thread 0's cumulative
// device-scope fence orders them before the flag.
__device__ __forceinline__ int publish(int* flag, bool ok, int value = 1) {
  const int all_ok = __syncthreads_and(ok ? 1 : 0);
  if (threadIdx.x == 0) {
    __threadfence();
    st_release(flag, value);
  }
  return all_ok;
}

// acc += A(64 x nk*KC) * B(64 x nk*KC)^T, streaming the operands tile column by tile column as
// their flags appear.  A is KCONTIG; B is KCONTIG (Cholesky) or MCONTIG (inverse).  Chunk kt
// belongs to tile column kt / 4, guarded by fa[(kt/4) * sa] and (if fb) fb[(kt/4) * sb].
template <int LB>
__device__ __forceinline__ void accumulate_flagged(double (&acc)[CfgP::MI][CfgP::NI][2], const double* Ag, int lda,
                                                   const double* Bg, int ldb, int nk, const int* fa, int sa,
                                                   const int* fb, int sb, int* abort_flag, double* pipe, bool& ok,
                                                   const int* yield_to, unsigned long long* waited = nullptr) {
  // the k permutation pays when both operands load 16 bytes at a time (measured: 27.4 -> 30.6 TFLOP/s in
  // isolation); with an MCONTIG B it loses 3 %, so the inverse's loop keeps the plain order
  constexpr bool PERM = (LB == KCONTIG);
  using E = Engine<CfgP, KCONTIG, LB, PERM>;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp / CfgP::WARPS_N) * CfgP::WM;
  const int wn0 = (warp % CfgP::WARPS_N) * CfgP::WN;
  const int g = lane >> 2, t = lane & 3;
  const int ncol = nk >> 2;
  // `known`: leading tile columns whose flags have been seen set (one coalesced look per warp)
  int known = 0;
  auto rescan = [&](int from) {
    known = from;
    for (int base = from; base < ncol; base += 32) {
      const int kc = base + lane;
      int f = 1;
      if (kc < ncol) f = (ld_flag(fa + (size_t)kc * sa) != 0) && (fb == nullptr || ld_flag(fb + (size_t)kc * sb) != 0);
      const unsigned missing = __ballot_sync(0xffffffffu, f == 0);
      if (missing) {
        known = base + __ffs(missing) - 1;
        return;
      }
      known = min(base + 32, ncol);
    }
  };
  rescan(0);
  auto issue = [&](int kt) {
    int look_from = -1;
    if ((kt & 3) == 0) {
      const int kc = kt >> 2;
      if (kc >= known) {
        // at the dependency frontier every L2 round trip is on the critical path: both flags are
        // polled together, and the look at the following columns waits until the loads are out
        const unsigned long long w0 = waited ? gtime() : 0ull;
        ok = wait_ready2(fa + (size_t)kc * sa, fb != nullptr ? fb + (size_t)kc * sb : nullptr, abort_flag) && ok;
        if (waited) *waited += gtime() - w0;
        look_from = kc + 1;
      }
    }
    double* stage = pipe + (kt % PSTAGES) * PSTAGE_DOUBLES;
    const double* ap = Ag + (size_t)kt * KC;
    const double* bp = (LB == KCONTIG) ? Bg + (size_t)kt * KC : Bg + (size_t)kt * KC * ldb;
    E::template load_operand<64, KCONTIG>(stage, ap, lda, tid);
    E::template load_operand<64, LB>(stage + E::SA::DOUBLES, bp, ldb, tid);
    if (look_from >= 0) rescan(look_from);  // producers usually publish several tile columns while we wait
  };
  // Ring of PSTAGES = 4 stages, two chunks (one "pair") per barrier.  While pair kt is in the math,
  // pair kt+2 is in flight into the other two stages: it is issued right after the barrier whenever
  // its flags are already known (the common case), and only after the math -- where a blocking flag
  // wait cannot delay work on data that has landed -- at the dependency frontier.
  static_assert(PSTAGES == 4, "ring logic below assumes two pairs");
  if (nk > 0) {
    issue(0);
    issue(1);
  }
  cp_async_commit();
  for (int kt = 0; kt < nk; kt += 2) {
    // Priority for the Cholesky chain: while a co-resident CTA on this SM works on a chain tile
    // (diagonal / first sub-diagonal), background tasks stand aside so it has the FP64 pipe alone.
    // (the flag is fetched here and only looked at after the math, so its L2 round trip is hidden)
    int busy = 0;
    if (yield_to != nullptr && lane == 0) busy = ld_flag(yield_to);
    cp_async_wait<0>();
    __syncthreads();
    const bool more = kt + 2 < nk;
    const bool early = more && (((kt + 2) >> 2) < known);  // kt+2 and kt+3 share a tile column
    if (early) {
      issue(kt + 2);
      issue(kt + 3);
      cp_async_commit();
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const double* sA = pipe + ((kt + h) % PSTAGES) * PSTAGE_DOUBLES;
      const double* sB = sA + E::SA::DOUBLES;
      if (PERM) {
#pragma unroll
        for (int kk = 0; kk < KC; kk += 8) {
          double af[2][CfgP::MI], bf[2][CfgP::NI];
          load_fragments<KCONTIG, CfgP::MI>(af, sA, E::SA::STRIDE, wm0 + g, kk + 2 * t);
          load_fragments<LB, CfgP::NI>(bf, sB, E::SB::STRIDE, wn0 + g, kk + 2 * t);
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
            for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
              for (int ni = 0; ni < CfgP::NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[h2][mi], bf[h2][ni]);
        }
      } else {
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) {
          double af[CfgP::MI], bf[CfgP::NI];
#pragma unroll
          for (int mi = 0; mi < CfgP::MI; ++mi) af[mi] = sA[(wm0 + mi * 8 + g) * E::SA::STRIDE + kk + t];
#pragma unroll
          for (int ni = 0; ni < CfgP::NI; ++ni) bf[ni] = sB[(kk + t) * E::SB::STRIDE + wn0 + ni * 8 + g];
#pragma unroll
          for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
            for (int ni = 0; ni < CfgP::NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
      }
    }
    if (more && !early) {
      issue(kt + 2);
      issue(kt + 3);
      cp_async_commit();
    }
    if (busy != 0) {  // lane 0 only; the other lanes wait for it at the next barrier
      const unsigned long long w0 = waited ? gtime() : 0ull;
      int spins = 0;
      while (ld_flag(yield_to) != 0 && ++spins < 256) __nanosleep(200);
      if (waited) waited[1] += gtime() - w0;
    }
  }
  cp_async_wait<0>();
}


// ---- ready-first claiming (inverse queues) ---------------------------------------------------
// In-order claiming makes a CTA block inside a task whose last inputs are still being produced
// (measured: 16 % of all CTA time at N = 4096).  The inverse's tiles and helpers therefore sit in
// two queues with per-task claim marks: warp 0 looks at a window of 32 unclaimed candidates per
// queue and takes one whose LAST inputs are already published (earlier inputs then are, too).
// Only if neither queue shows a ready task does it fall back to the next unclaimed task of the
// original dependency-safe order (a ticket counter walks it) -- which may block, exactly as in-order
// claiming would, and keeps the no-deadlock argument: every blocking claim has all its predecessors
// claimed.
__device__ __forceinline__ bool inverse_task_ready(const FactorArgs& a, const FactorTask& tk) {
  const int nt = a.nt;
  if (tk.type == TASK_TRTRI) {  // X_rJ: last inputs X_{r-1,J}, L_{r,r-1}, X_rr and (if split) its helper's partial sum
    const int r = tk.i, J = tk.j;
    int f = ld_flag(a.xready + (size_t)(r - 1) * nt + J) & ld_flag(a.ready + (size_t)r * nt + (r - 1)) &
            ld_flag(a.xready + (size_t)r * nt + r);
    if (f != 0 && tk.k > J) f = ld_flag(a.pready + (size_t)r * nt + J) >= tk.k;
    return f != 0;
  }
  // helper (r, J, [k0, k)): last inputs X_{k-1,J} and L_{r,k-1}, and the previous chunk's partial sum
  int f = ld_flag(a.xready + (size_t)(tk.k - 1) * nt + tk.j) & ld_flag(a.ready + (size_t)tk.i * nt + (tk.k - 1));
  if (f != 0 && tk.k0 > tk.j) f = ld_flag(a.pready + (size_t)tk.i * nt + tk.j) >= tk.k0;
  return f != 0;
}

struct WindowLook {
  unsigned lo;        // window start (the queue's low-water mark when looked at)
  unsigned ready;     // lanes whose candidate is unclaimed and ready
  unsigned unclaimed; // lanes whose candidate is unclaimed
};

// Executed by all 32 lanes of warp 0.
__device__ __forceinline__ WindowLook look_window(const FactorArgs& a, int q, int lane) {
  WindowLook w;
  const unsigned qn = (unsigned)(a.qbase[q + 1] - a.qbase[q]);
  for (;;) {
    unsigned lo = 0;
    if (lane == 0) lo = (unsigned)ld_flag(reinterpret_cast<const int*>(a.counter + q));
    lo = __shfl_sync(0xffffffffu, lo, 0);
    w.lo = lo;
    w.ready = w.unclaimed = 0u;
    if (lo >= qn) return w;
    const unsigned idx = lo + lane;
    bool un = false, rdy = false;
    if (idx < qn) {
      const unsigned g = (unsigned)a.qbase[q] + idx;
      un = ld_flag(a.claimed + g) == 0;
      if (un) rdy = inverse_task_ready(a, a.tasks[g]);
    }
    w.unclaimed = __ballot_sync(0xffffffffu, un);
    w.ready = __ballot_sync(0xffffffffu, rdy);
    if (w.unclaimed != 0u) return w;
    // the whole window is claimed: move the low-water mark and look again
    if (lane == 0) atomicMax(a.counter + q, min(lo + 32u, qn));
    __syncwarp();
  }
}

// Try to take candidate `pick` (lane index) of window w in queue q; returns the global task index or NONE.
__device__ __forceinline__ unsigned take_candidate(const FactorArgs& a, int q, const WindowLook& w, int pick, int lane) {
  unsigned got = 0xffffffffu;
  if (lane == 0) {
    const unsigned g = (unsigned)a.qbase[q] + w.lo + (unsigned)pick;
    if (atomicCAS(a.claimed + g, 0, 1) == 0) {
      got = g;
      if (pick == __ffs(w.unclaimed) - 1) atomicMax(a.counter + q, w.lo + (unsigned)pick + 1u);  // prefix fully claimed
    }
  }
  return __shfl_sync(0xffffffffu, got, 0);
}

// Both inverse queues looked at in ONE pass (queue 2 holds tiles only, queue 3 helpers only): the low-water marks, the
// claim marks and the readiness flags of the two windows are loaded side by side, so a look costs three dependent L2
// round trips for both queues together instead of four to five per queue.
struct TwoLooks {
  WindowLook w[2];
};
__device__ __forceinline__ TwoLooks look_both(const FactorArgs& a, int lane) {
  TwoLooks o;
  const unsigned qn2 = (unsigned)(a.qbase[3] - a.qbase[2]), qn3 = (unsigned)(a.qbase[4] - a.qbase[3]);
  const int nt = a.nt;
  for (;;) {
    unsigned lo = 0;
    if (lane < 2) lo = (unsigned)ld_flag(reinterpret_cast<const int*>(a.counter + 2 + lane));
    const unsigned lo2 = __shfl_sync(0xffffffffu, lo, 0), lo3 = __shfl_sync(0xffffffffu, lo, 1);
    const bool v2 = lo2 + lane < qn2, v3 = lo3 + lane < qn3;
    const unsigned g2 = (unsigned)a.qbase[2] + lo2 + lane, g3 = (unsigned)a.qbase[3] + lo3 + lane;
    int c2 = 1, c3 = 1;
    if (v2) c2 = ld_flag(a.claimed + g2);
    if (v3) c3 = ld_flag(a.claimed + g3);
    // readiness (inverse_task_ready's conditions): both task records first, then every flag, none of the loads
    // waiting for the claim marks or for one another
    FactorTask t2 = {0, 1, 0, 0, 0, 0}, t3 = {0, 0, 0, 1, 0, 0};
    if (v2) t2 = a.tasks[g2];
    if (v3) t3 = a.tasks[g3];
    int f2a = 0, f2b = 0, f2c = 0, p2 = 0, need2 = 0, f3a = 0, f3b = 0, p3 = 0, need3 = 0;
    if (v2) {
      const int r = t2.i, J = t2.j;
      f2a = ld_flag(a.xready + (size_t)(r - 1) * nt + J);
      f2b = ld_flag(a.ready + (size_t)r * nt + (r - 1));
      f2c = ld_flag(a.xready + (size_t)r * nt + r);
      if (t2.k > J) {
        need2 = t2.k;
        p2 = ld_flag(a.pready + (size_t)r * nt + J);
      }
    }
    if (v3) {
      f3a = ld_flag(a.xready + (size_t)(t3.k - 1) * nt + t3.j);
      f3b = ld_flag(a.ready + (size_t)t3.i * nt + (t3.k - 1));
      if (t3.k0 > t3.j) {
        need3 = t3.k0;
        p3 = ld_flag(a.pready + (size_t)t3.i * nt + t3.j);
      }
    }
    const bool un2 = v2 && c2 == 0, un3 = v3 && c3 == 0;
    const bool rdy2 = un2 && (f2a & f2b & f2c) != 0 && p2 >= need2;
    const bool rdy3 = un3 && (f3a & f3b) != 0 && p3 >= need3;
    o.w[0].lo = lo2;
    o.w[1].lo = lo3;
    o.w[0].unclaimed = __ballot_sync(0xffffffffu, un2);
    o.w[0].ready = __ballot_sync(0xffffffffu, rdy2);
    o.w[1].unclaimed = __ballot_sync(0xffffffffu, un3);
    o.w[1].ready = __ballot_sync(0xffffffffu, rdy3);
    // a window that is claimed throughout: move that queue's low-water mark and look again
    const bool adv2 = lo2 < qn2 && o.w[0].unclaimed == 0u, adv3 = lo3 < qn3 && o.w[1].unclaimed == 0u;
    if (!adv2 && !adv3) return o;
    if (lane == 0) {
      if (adv2) atomicMax(a.counter + 2, min(lo2 + 32u, qn2));
      if (adv3) atomicMax(a.counter + 3, min(lo3 + 32u, qn3));
    }
    __syncwarp();
  }
}

#ifndef EB_CLAIM_BOTH
#define EB_CLAIM_BOTH 1  // (0: one queue at a time, a second look even when nothing was ready -- N = 4096 in groups of eight 2.566 vs 2.550 ms, N = 2048 0.419 vs 0.412)
#endif
// Claim one task from the inverse queues (2: tiles, 3: helpers).  Returns NONE when all are claimed -- or, with
// ready_only, when no task whose inputs are published is in sight.
__device__ __forceinline__ unsigned claim_inverse(const FactorArgs& a, int lane, bool ready_only = false) {
#if EB_CLAIM_BOTH
#pragma unroll 1
  for (int attempt = 0; attempt < 2; ++attempt) {
    const TwoLooks both = look_both(a, lane);
    bool lost = false;  // a second look only pays when a candidate was there and somebody else took it
#pragma unroll
    for (int qi = 0; qi < 2; ++qi) {
      const WindowLook w = both.w[qi];
      if (w.ready == 0u) continue;
      const int n = __popc(w.ready), want = (int)((blockIdx.x + attempt) % (unsigned)n);
      const unsigned g = take_candidate(a, 2 + qi, w, __fns(w.ready, 0, want + 1), lane);
      if (g != 0xffffffffu) return g;
      lost = true;
    }
    if (!lost) break;
  }
#else
#pragma unroll 1
  for (int attempt = 0; attempt < 2; ++attempt) {
#pragma unroll 1
    for (int q = 2; q <= 3; ++q) {
      const WindowLook w = look_window(a, q, lane);
      if (w.ready == 0u) continue;
      // spread concurrent claimers over the ready candidates
      const int n = __popc(w.ready), want = (int)((blockIdx.x + attempt) % (unsigned)n);
      const unsigned g = take_candidate(a, q, w, __fns(w.ready, 0, want + 1), lane);
      if (g != 0xffffffffu) return g;
    }
  }
#endif
  if (ready_only) return 0xffffffffu;
  // Nothing ready (or lost the races): the next unclaimed task of the safe order.  Tickets walk that
  // order exactly like in-order claiming, skipping what the ready-first path has taken.
  unsigned got = 0xffffffffu;
  if (lane == 0) {
    for (;;) {
      if (ld_flag(reinterpret_cast<const int*>(a.counter + 4)) >= a.n_inv) break;
      const unsigned s = atomicAdd(a.counter + 4, 1u);
      if (s >= (unsigned)a.n_inv) break;
      const unsigned g = (unsigned)a.inv_order[s];
      if (atomicCAS(a.claimed + g, 0, 1) == 0) {
        got = g;
        break;
      }
    }
  }
  return __shfl_sync(0xffffffffu, got, 0);
}

// Shared memory of one CTA (two CTAs per SM).  The finalisation buffers alias the cp.async
// ring: they are only touched after the accumulation loop has drained it.
template <int DP>
struct FactorSmem {
  static constexpr int PIPE = PSTAGES * PSTAGE_DOUBLES;
  static constexpr int C_OFF = 0;                   // [64][65]   C tile / substitution output
  static constexpr int L_OFF = C_OFF + 64 * PS;     // [64][65]   L_jj
  static constexpr int T_OFF = 0;                   // [64][68]   parked accumulator (inverse tasks)
  static constexpr int DINV_OFF = T_OFF + 64 * TSK;  // 4 x [64][KS] slabs of the diagonal inverse
  static constexpr int XI_OFF = 0;                  // [DP][64]  (K generation happens before the ring is used)
  static constexpr int XJ_OFF = XI_OFF + DP * 64;   // [DP][64]
  static constexpr int COL_OFF = PIPE;              // [128]
  static constexpr int RD_OFF = COL_OFF + 128;      // [64]
  static constexpr int DOUBLES = RD_OFF + 64 + 16;
  static constexpr int BYTES = DOUBLES * (int)sizeof(double);
  static_assert(2 * 64 * PS <= PIPE, "finalisation buffers must fit the ring");
  static_assert(2 * DP * 64 <= PIPE, "input staging must fit the ring");
  static_assert(DINV_OFF + 4 * 64 * KS <= PIPE, "inverse finalisation must fit the ring");
  static_assert(2 * BYTES + 2048 <= 227 * 1024, "two CTAs per SM");
};

#ifndef EB_BATCH_UNCHECKED_HELP
#define EB_BATCH_UNCHECKED_HELP 1  // shared launches: an idle inverse CTA takes the next Cholesky task whether or not its inputs are there yet
#endif
// Claim the next task of one problem for a CTA of the given role; NONE if all its queues are empty.
__device__ __forceinline__ unsigned claim_task(const FactorArgs& a, int role, int lane, bool unchecked_help = false) {
  unsigned tk0 = 0xffffffffu;
  // An inverse CTA that finds no inverse task with published inputs takes the next Cholesky task -- if THAT one's
  // last inputs are published -- instead of blocking inside an inverse task: the Cholesky is what the inverse waits
  // for.  (Measured, ms per evaluation, without -> with: groups of four at N = 4096 2.73 -> 2.58, groups of eight at
  // N = 2048 0.46 -> 0.42, stand-alone N = 8176 19.4 -> 19.1.)
  if (role == 2 && !a.value_only) {
    tk0 = claim_inverse(a, lane, true);
    if (tk0 != 0xffffffffu) return tk0;
    if (lane == 0) {
      const unsigned qn = (unsigned)(a.qbase[2] - a.qbase[1]);
      const unsigned head = (unsigned)ld_flag(reinterpret_cast<const int*>(a.counter + 1));
      if (head < qn) {
        const FactorTask& t = a.tasks[a.qbase[1] + head];
        // last tile column the task reads: j - 1 for a tile's own task, k - 1 for a pre-accumulation helper
        const int last = (t.type == TASK_POTRF_PART ? t.k : t.j) - 1;
        const int nt = a.nt;
        bool ready = true;
        if (last >= 0 && !unchecked_help) ready = ld_flag(a.ready + (size_t)t.i * nt + last) != 0 && (t.i == t.j || t.i >= nt || ld_flag(a.ready + (size_t)t.j * nt + last) != 0);
        if (ready) {
          const unsigned k = atomicAdd(a.counter + 1, 1u);
          if (k < qn) tk0 = (unsigned)a.qbase[1] + k;
        }
      }
    }
    tk0 = __shfl_sync(0xffffffffu, tk0, 0);
    if (tk0 != 0xffffffffu) return tk0;
  }
#pragma unroll 1
  for (int attempt = 0; attempt < 3 && tk0 == 0xffffffffu; ++attempt) {
    const int q = (role + attempt) % 3;
    if (q == 2) {
      tk0 = claim_inverse(a, lane);
    } else {
      if (lane == 0) {
        const unsigned qn = (unsigned)(a.qbase[q + 1] - a.qbase[q]);
        if (ld_flag(reinterpret_cast<const int*>(a.counter + q)) < (int)qn) {
          const unsigned k = atomicAdd(a.counter + q, 1u);
          if (k < qn) tk0 = (unsigned)a.qbase[q] + k;
        }
      }
      tk0 = __shfl_sync(0xffffffffu, tk0, 0);
    }
  }
  return tk0;
}

template <int DP, bool BATCH>
__global__ void __launch_bounds__(256, 2) factor_dataflow_kernel(const FactorBatch batch) {
  extern __shared__ __align__(16) double gsm[];
  using SM = FactorSmem<DP>;
  __shared__ unsigned s_task;
  __shared__ int s_role;
  __shared__ int s_prob;
  __shared__ unsigned long long s_wait[2];  // diagnostics: time thread 0 spent waiting in the accumulation loops
  __shared__ int s_pflags[16];  // per-panel flags of the two tile finalisers (epoch-valued, never reset)
  double* sC = gsm + SM::C_OFF;
  double* sL = gsm + SM::L_OFF;
  double* sT = gsm + SM::T_OFF;
  double* sD = gsm + SM::DINV_OFF;
  double* sXi = gsm + SM::XI_OFF;
  double* sXj = gsm + SM::XJ_OFF;
  double* colbuf = gsm + SM::COL_OFF;
  double* rdiag = gsm + SM::RD_OFF;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp / CfgP::WARPS_N) * CfgP::WM;
  const int wn0 = (warp % CfgP::WARPS_N) * CfgP::WN;
  const int g = lane >> 2, t = lane & 3;
  const int ld = batch.p[0].ld, nt = batch.p[0].nt;  // the problems of a batch share their shapes
  const int nb = BATCH ? batch.nb : 1;
  if (tid < 16) s_pflags[tid] = 0;
  if (tid == 0) s_role = -1;

  for (;;) {
    __syncthreads();  // previous task fully retired (smem reuse, s_task reuse)
    // Queues.  (0) the Cholesky's chain tiles (diagonal + first sub-diagonal + diagonal inverse), served
    // by a dozen CTAs in order; (1) all other Cholesky tasks in order, whose CTAs run ahead along the
    // dependency chain exactly as if the inverse were not there; (2, 3) the inverse's tiles and helpers,
    // claimed ready-first (claim_inverse) by CTAs co-resident with the queue-1 CTAs.  A CTA whose
    // queue is empty helps the others.
    if (warp == 0) {
      const FactorArgs& a0 = batch.p[0];
      const int local = (int)blockIdx.x / nb;  // CTA index within its home problem
      if (lane == 0 && s_role < 0) {
        // A dozen CTAs serve the chain queue, the rest of the first wave the Cholesky queue, the
        // second wave the inverse queues.  (Measured alternatives: reserving whole SMs for the chain by
        // %smid -- 6 or 12 SMs, band of 1 or 3 sub-diagonals -- made its tile factorisations 25 %
        // faster but lost more in look-ahead and bulk capacity: 307-314 vs 321 evals/s at N = 4096.)
        const int slot = BATCH ? local : (int)(blockIdx.x % (unsigned)a0.n_sms);
        // (batched full evaluations: the CTAs of a problem alternate between its Cholesky and its inverse queues)
        const int first_wave = BATCH ? ((local & 1) == 0) : (blockIdx.x < (unsigned)a0.n_sms);
        s_role = (slot < a0.n_chain_sms) ? 0 : ((first_wave || a0.value_only) ? 1 : 2);
      }
      __syncwarp();
      const int role = s_role;
      unsigned tk0 = 0xffffffffu;
      int prob = 0;
#pragma unroll 1
      for (int pp = 0; pp < nb && tk0 == 0xffffffffu; ++pp) {
        prob = BATCH ? ((int)(blockIdx.x % (unsigned)nb) + pp) % nb : 0;
        tk0 = claim_task(batch.p[prob], role, lane, BATCH && EB_BATCH_UNCHECKED_HELP);
      }
      if (lane == 0) {
        s_task = tk0;
        s_prob = prob;
      }
    }
    __syncthreads();
    const unsigned task = s_task;
    if (task == 0xffffffffu) break;
    const FactorArgs& a = batch.p[BATCH ? s_prob : 0];
    const FactorTask tk = a.tasks[task];
    const int ti = tk.i, tj = tk.j;
    const int i0 = ti * 64, j0 = tj * 64;
    bool ok = true;
    unsigned long long* wait_acc = (a.trace != nullptr && tid == 0) ? s_wait : nullptr;
    EB_STAMP(0);
    if (a.trace != nullptr && tid == 0) {
      a.trace[(size_t)task * TRACE_SLOTS + 6] = smid();
      a.trace[(size_t)task * TRACE_SLOTS + 7] = blockIdx.x;
      s_wait[0] = s_wait[1] = 0ull;
    }
    double acc[CfgP::MI][CfgP::NI][2];

    if (tk.type == TASK_XDIAG) {
      double v[2][8];
#pragma unroll
      for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int c = 0; c < 8; ++c) v[sl][c] = (lane + 32 * sl == 8 * warp + c) ? 1.0 : 0.0;
      int* my_sm = a.sm_chain + smid();
      ok = wait_ready(a.lready + (size_t)tj * 8, a.abort_flag) && ok;  // from here on the wait is bounded
      if (tid == 0) atomicAdd(my_sm, 1);
      EB_STAMP(3);
      {
        int* abort_flag = a.abort_flag;
        ok = tile_trsm_streamed(v, sL, rdiag, sC, a.A + (size_t)j0 * ld + j0, ld, a.rdiag + (size_t)tj * 64,
                                a.lready + (size_t)tj * 8,
                                [abort_flag](const int* f) { return wait_ready(f, abort_flag); }) && ok;
      }
      __syncthreads();
      // sC = I L_jj^-T = Y_jj (upper triangular).  X_jj = Y_jj^T goes to the compact [nt][64][64] buffer that the
      // tile finalisations read it from (k-contiguous rows), Y_jj to its place in the matrix of Y.
      double* Xt = a.Xd + (size_t)tj * a.xd_tile;
      const int xld = a.xd_ld;
      for (int e = tid; e < 64 * 64; e += 256) {
        const int r = e >> 6, c = e & 63;
        Xt[(size_t)r * xld + c] = (c <= r) ? sC[c * PS + r] : 0.0;
      }
      if (!a.value_only) {
        double* Yt = a.W + (size_t)j0 * ld + j0;
        for (int e = tid; e < 64 * 64; e += 256) {
          const int r = e >> 6, c = e & 63;
          Yt[(size_t)r * ld + c] = (c >= r) ? sC[r * PS + c] : 0.0;
        }
      }
      EB_STAMP(4);
      const int x_ok = publish(a.xready + (size_t)tj * nt + tj, ok);
      if (tid == 0) atomicSub(my_sm, 1);
      EB_STAMP(5);
      if (!x_ok) break;
      continue;
    }
    if (tk.type == TASK_POTRF_PART) {
      // ============  background pre-accumulation of a chain tile: sum_{c<k} L_ic L_jc^T  ============
#pragma unroll
      for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < CfgP::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
      accumulate_flagged<KCONTIG>(acc, a.A + (size_t)i0 * ld, ld, a.A + (size_t)j0 * ld, ld, tk.k * 4,
                                  a.ready + (size_t)ti * nt, 1, (ti != tj) ? a.ready + (size_t)tj * nt : nullptr, 1,
                                  a.abort_flag, gsm, ok, a.sm_chain + smid(), wait_acc);
      double* cs = a.A + (size_t)i0 * ld + j0;  // the tile's own storage (nobody reads it before L_ij is published)
#pragma unroll
      for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < CfgP::NI; ++ni) {
          double2 val;
          val.x = acc[mi][ni][0];
          val.y = acc[mi][ni][1];
          *reinterpret_cast<double2*>(cs + (size_t)(wm0 + mi * 8 + g) * ld + wn0 + ni * 8 + 2 * t) = val;
        }
      EB_STAMP(4);
      const int part_ok = publish(a.cready + (size_t)ti * nt + tj, ok);
      EB_STAMP(5);
      if (!part_ok) break;
      continue;
    }
    if (tk.type != TASK_POTRF) {
      // ==========  Y_Jr = -(sum_{K=J}^{r-1} Y_JK L_rK^T) X_rr^T   (X_rJ of the inverse, stored transposed)  ==========
      const bool part = tk.type == TASK_TRTRI_PART;
      const int kbeg = part ? tk.k0 : tk.k;  // first tile column this task adds
      const int kend = part ? tk.k : ti;     // one past the last
      double* scratch = a.W + (size_t)i0 * ld + j0;  // mirrored lower tile (r, J): never read as Y
      if (kbeg > tj) {
        // the helpers of a tile form a chain; pready holds the number of tile columns summed so far
        ok = wait_ready(a.pready + (size_t)ti * nt + tj, a.abort_flag, kbeg) && ok;
#pragma unroll
        for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < CfgP::NI; ++ni) {
            const double2 val =
                __ldcg(reinterpret_cast<const double2*>(scratch + (size_t)(wm0 + mi * 8 + g) * ld + wn0 + ni * 8 + 2 * t));
            acc[mi][ni][0] = val.x;
            acc[mi][ni][1] = val.y;
          }
      } else {
#pragma unroll
        for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < CfgP::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
      }
      // acc[m = J-local][n = r-local] += sum_k Y[J0+m][k] L[r0+n][k]: both operands k-contiguous
      accumulate_flagged<KCONTIG>(acc, a.W + (size_t)j0 * ld + (size_t)kbeg * 64, ld,
                                  a.A + (size_t)i0 * ld + (size_t)kbeg * 64, ld, (kend - kbeg) * 4,
                                  a.xready + (size_t)kbeg * nt + tj, nt, a.ready + (size_t)ti * nt + kbeg, 1,
                                  a.abort_flag, gsm, ok, a.sm_chain + smid(), wait_acc);
      if (part) {
#pragma unroll
        for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < CfgP::NI; ++ni) {
            double2 val;
            val.x = acc[mi][ni][0];
            val.y = acc[mi][ni][1];
            *reinterpret_cast<double2*>(scratch + (size_t)(wm0 + mi * 8 + g) * ld + wn0 + ni * 8 + 2 * t) = val;
          }
        EB_STAMP(4);
        const int part_ok = publish(a.pready + (size_t)ti * nt + tj, ok, kend);
        EB_STAMP(5);
        if (!part_ok) break;
        continue;
      }
      EB_STAMP(1);
      __syncthreads();  // ring drained by every warp before the finalisation buffers overwrite it
      // park the sum as an [m][k] operand (k = r-local)
#pragma unroll
      for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < CfgP::NI; ++ni) {
          double* pc = sT + (wm0 + mi * 8 + g) * TSK + wn0 + ni * 8 + 2 * t;
          pc[0] = acc[mi][ni][0];
          pc[1] = acc[mi][ni][1];
        }
      ok = wait_ready(a.xready + (size_t)ti * nt + ti, a.abort_flag) && ok;
      yield_to_chain(a.sm_chain + smid());
      EB_STAMP(2);
      using EK = Engine<CfgP, KCONTIG, KCONTIG>;
      const double* Dg = a.Xd + (size_t)ti * a.xd_tile;  // X_rr: rows n = r-local, k contiguous
      const int xld = a.xd_ld;
#pragma unroll
      for (int c = 0; c < 4; ++c) EK::template load_operand<64, KCONTIG>(sD + c * 64 * KS, Dg + c * KC, xld, tid);
      cp_async_commit();
#pragma unroll
      for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < CfgP::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
      cp_async_wait<0>();
      __syncthreads();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int kk = 0; kk < KC; kk += 8) {
          double af[2][CfgP::MI], bf[2][CfgP::NI];
          load_fragments<KCONTIG, CfgP::MI>(af, sT, TSK, wm0 + g, c * KC + kk + 2 * t);
          load_fragments<KCONTIG, CfgP::NI>(bf, sD + c * 64 * KS, KS, wn0 + g, kk + 2 * t);
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
            for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
              for (int ni = 0; ni < CfgP::NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[h2][mi], bf[h2][ni]);
        }
      }
      double* Wt = a.W + (size_t)j0 * ld + i0;  // tile (J, r) of Y
#pragma unroll
      for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < CfgP::NI; ++ni) {
          double2 val;
          val.x = -acc[mi][ni][0];
          val.y = -acc[mi][ni][1];
          *reinterpret_cast<double2*>(Wt + (size_t)(wm0 + mi * 8 + g) * ld + wn0 + ni * 8 + 2 * t) = val;
        }
      EB_STAMP(4);
      const int all_ok = publish(a.xready + (size_t)ti * nt + tj, ok);
      EB_STAMP(5);
      if (!all_ok) break;
      continue;
    }

    // ==================================  Cholesky tile (i, j)  ==================================
    // Chain tiles raise their SM's priority flag only for the latency-bound finalisation (never
    // while they wait for a producer: a background task on the same SM may be that producer).
    const bool chain_tile = (ti - tj) <= 1 && ti < nt;  // the band served by the chain queue (CHAIN_BAND in eb200.cu)
    int* my_sm = a.sm_chain + smid();
    if (!chain_tile) yield_to_chain(my_sm);  // the exp()-heavy K generation competes for the FP64 pipe
    if (ti == nt) {
      // Residual row (value-only evaluations): the residual rides along as row 0 of an extra tile
      // row of the matrix being factorised, [K r; r^T .] = [L 0; z^T .][L 0; z^T .]^T, so the same tile
      // pipeline that produces L_ij leaves z = L^-1 r in row nt*64 of A.  acc = -r_j in row 0.
#pragma unroll
      for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
        for (int c = 0; c < CfgP::NI * 2; ++c) {
          const int lr = wm0 + mi * 8 + g, gj = j0 + wn0 + (c >> 1) * 8 + 2 * t + (c & 1);
          acc[mi][c >> 1][c & 1] = (lr == 0) ? -a.r[gj] : 0.0;
        }
    } else {
    // stage the training inputs of the tile's rows / columns for the K_ij generation
    for (int e = tid; e < 64 * DP; e += 256) {
      int r = e / DP, m = e % DP;
      sXi[m * 64 + r] = a.X[(size_t)(i0 + r) * DP + m];
      sXj[m * 64 + r] = a.X[(size_t)(j0 + r) * DP + m];
    }
    __syncthreads();
    // acc = -K_ij, generated straight into the accumulator layout (before any panel wait, so the
    // exp() work overlaps the wait for the producers)
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
    }
    __syncthreads();  // the staged inputs alias the cp.async ring
    EB_STAMP(1);
    const int kbeg = tk.k;  // > 0: tile columns [0, kbeg) were pre-accumulated by a TASK_POTRF_PART
    if (kbeg > 0) {
      ok = wait_ready(a.cready + (size_t)ti * nt + tj, a.abort_flag) && ok;
      const double* cs = a.A + (size_t)i0 * ld + j0;
#pragma unroll
      for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < CfgP::NI; ++ni) {
          const double2 val =
              __ldcg(reinterpret_cast<const double2*>(cs + (size_t)(wm0 + mi * 8 + g) * ld + wn0 + ni * 8 + 2 * t));
          acc[mi][ni][0] += val.x;
          acc[mi][ni][1] += val.y;
        }
    }
    // left-looking accumulation over the finished panels: acc += sum_k L_ik L_jk^T
    accumulate_flagged<KCONTIG>(acc, a.A + (size_t)i0 * ld + (size_t)kbeg * 64, ld, a.A + (size_t)j0 * ld + (size_t)kbeg * 64, ld,
                                (tj - kbeg) * 4, a.ready + (size_t)ti * nt + kbeg, 1,
                                (ti != tj) ? a.ready + (size_t)tj * nt + kbeg : nullptr, 1, a.abort_flag, gsm, ok,
                                chain_tile ? nullptr : my_sm, wait_acc);
    __syncthreads();  // ring drained by every warp before the finalisation buffers overwrite it
    double* At = a.A + (size_t)i0 * ld + j0;
    const bool near_tile = (ti - tj) <= NEAR_BAND && ti < nt;  // finishes by substitution against the streamed panels
    if (!near_tile) {
      // Off-chain tiles: L_ij = C X_jj^T with the explicit inverse of the diagonal tile (published by
      // its owner a few microseconds after L_jj itself) -- one 64^3 tensor-core product instead of a
      // latency-bound substitution.  Tile (i, j) is an input of tile (i, j+1), so every tile row is a
      // dependency chain of its own, and what a tile does after its last input arrives bounds the
      // whole factorisation as much as the diagonal chain does.
#pragma unroll
      for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < CfgP::NI; ++ni) {
          double* pc = sT + (wm0 + mi * 8 + g) * TSK + wn0 + ni * 8 + 2 * t;  // C = -acc as a [m][k] operand
          pc[0] = -acc[mi][ni][0];
          pc[1] = -acc[mi][ni][1];
        }
      EB_STAMP(2);
      ok = wait_ready(a.xready + (size_t)tj * nt + tj, a.abort_flag) && ok;
      yield_to_chain(my_sm);
      EB_STAMP(3);
      using EK = Engine<CfgP, KCONTIG, KCONTIG>;
      const double* Dg = a.Xd + (size_t)tj * a.xd_tile;  // X_jj
      const int xld = a.xd_ld;
#pragma unroll
      for (int c = 0; c < 4; ++c) EK::template load_operand<64, KCONTIG>(sD + c * 64 * KS, Dg + c * KC, xld, tid);
      cp_async_commit();
#pragma unroll
      for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < CfgP::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
      cp_async_wait<0>();
      __syncthreads();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int kk = 0; kk < KC; kk += 8) {
          double af[2][CfgP::MI], bf[2][CfgP::NI];
          load_fragments<KCONTIG, CfgP::MI>(af, sT, TSK, wm0 + g, c * KC + kk + 2 * t);
          load_fragments<KCONTIG, CfgP::NI>(bf, sD + c * 64 * KS, KS, wn0 + g, kk + 2 * t);
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
            for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
              for (int ni = 0; ni < CfgP::NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[h2][mi], bf[h2][ni]);
        }
      }
#pragma unroll
      for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < CfgP::NI; ++ni) {
          double2 val;
          val.x = acc[mi][ni][0];
          val.y = acc[mi][ni][1];
          *reinterpret_cast<double2*>(At + (size_t)(wm0 + mi * 8 + g) * ld + wn0 + ni * 8 + 2 * t) = val;
        }
      EB_STAMP(4);
      const int all_ok = publish(a.ready + (size_t)ti * nt + tj, ok);
      EB_STAMP(5);
      if (!all_ok) break;
      continue;
    }
    // C = -acc parked in shared memory
#pragma unroll
    for (int mi = 0; mi < CfgP::MI; ++mi) {
      const int lr = wm0 + mi * 8 + g;
#pragma unroll
      for (int c = 0; c < CfgP::NI * 2; ++c) sC[lr * PS + wn0 + (c >> 1) * 8 + 2 * t + (c & 1)] = -acc[mi][c >> 1][c & 1];
    }
    __syncthreads();
    EB_STAMP(2);
    double v[2][8];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
      for (int c = 0; c < 8; ++c) v[sl][c] = sC[(lane + 32 * sl) * PS + 8 * warp + c];

    if (ti == tj) {
      if (tid == 0) atomicAdd(my_sm, 1);
      // (the panels of L_jj and the reciprocal pivots are streamed to global memory as they are produced)
      // (epoch: unique per call within this CTA -- task numbers repeat across the problems of a batch)
      tile_cholesky(v, sL, rdiag, a.info, j0, s_pflags, (int)task + 1 + (BATCH ? s_prob * a.ntasks : 0), At, ld, a.rdiag + (size_t)tj * 64,
                    a.lready + (size_t)tj * 8);
      EB_STAMP(4);
      int all_ok = publish(a.ready + (size_t)ti * nt + tj, ok);
      if (tid == 0) atomicSub(my_sm, 1);
      EB_STAMP(5);
      // ---- off the critical path: the log-determinant share ----
      if (tid < 64) colbuf[tid] = log(rdiag[tid]);
      __syncthreads();
      if (tid == 0) {
        double sum = 0.0;
        for (int j = 0; j < 64; ++j) sum += colbuf[j];
        a.logdet_part[tj] = -sum;  // sum_j log L_jj = -sum_j log(1 / L_jj)
      }
      if (!all_ok) break;
    } else {
      // near tiles (j+1 .. j+NEAR_BAND, j): substitution against L_jj, panel by panel as the diagonal
      // tile's owner streams them out (no priority flag here: this phase waits for panels of another CTA, and a
      // flag held across a wait can starve the very producer it waits for)
      // The chain tile (j+1, j) takes the FP64 pipe of its SM for itself once panel 0 is out: from
      // then on the diagonal tile's owner is inside its (bounded) tile factorisation.
      if (chain_tile) {
        ok = wait_ready(a.lready + (size_t)tj * 8, a.abort_flag) && ok;
        if (tid == 0) atomicAdd(my_sm, 1);
      }
      EB_STAMP(3);
      {
        int* abort_flag = a.abort_flag;
        ok = tile_trsm_streamed(v, sL, rdiag, sC, a.A + (size_t)j0 * ld + j0, ld, a.rdiag + (size_t)tj * 64,
                                a.lready + (size_t)tj * 8,
                                [abort_flag](const int* f) { return wait_ready(f, abort_flag); }) && ok;
      }
      __syncthreads();
      for (int e = tid; e < 64 * 64; e += 256) {
        const int rr = e >> 6, c = e & 63;
        At[(size_t)rr * ld + c] = sC[rr * PS + c];
      }
      EB_STAMP(4);
      const int all_ok = publish(a.ready + (size_t)ti * nt + tj, ok);
      if (chain_tile && tid == 0) atomicSub(my_sm, 1);
      EB_STAMP(5);
      if (!all_ok) break;
    }
  }
}


// Micro-benchmark of the dataflow main loop in isolation (all flags preset): C (64x64 tiles) =
// A * B^T with B KCONTIG, or A * B with B MCONTIG; used to tune the loop without the chain.
template <int LB>
__global__ void __launch_bounds__(256, 2) accumulate_bench_kernel(const double* A, int lda, const double* B, int ldb,
                                                                  double* C, int ldc, int nk, int tiles_n, int ntiles,
                                                                  const int* ones, int* abort_flag) {
  extern __shared__ __align__(16) double gsm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp / CfgP::WARPS_N) * CfgP::WM;
  const int wn0 = (warp % CfgP::WARPS_N) * CfgP::WN;
  const int g = lane >> 2, t = lane & 3;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int i0 = (tile / tiles_n) * 64, j0 = (tile % tiles_n) * 64;
    double acc[CfgP::MI][CfgP::NI][2];
#pragma unroll
    for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < CfgP::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    bool ok = true;
    __syncthreads();
    accumulate_flagged<LB>(acc, A + (size_t)i0 * lda, lda, (LB == KCONTIG) ? B + (size_t)j0 * ldb : B + j0, ldb, nk, ones, 0,
                           nullptr, 0, abort_flag, gsm, ok, nullptr);
#pragma unroll
    for (int mi = 0; mi < CfgP::MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < CfgP::NI; ++ni) {
        double2 val;
        val.x = acc[mi][ni][0];
        val.y = acc[mi][ni][1];
        *reinterpret_cast<double2*>(C + (size_t)(i0 + wm0 + mi * 8 + g) * ldc + j0 + wn0 + ni * 8 + 2 * t) = val;
      }
  }
}

}  // namespace eb
