// C ABI implementation: GP handle, launch program (tile-task lists), CUDA graphs.
// See include/emulator_b200.h for the contract and DESIGN.md for the algorithm.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/emulator_b200.h"
#include "kernels.cuh"
#include "factor_dataflow.cuh"
#include "tma_gemm.cuh"

using namespace eb;

namespace {

thread_local std::string g_err;

int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}

#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      return fail(std::string(#call) + " failed: " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                  std::to_string(__LINE__) + ")");                                                    \
    }                                                                                                 \
  } while (0)

// Tunables (development: read once from the environment).
int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

int pad_dim(int d) {
  static const int buckets[] = {2, 3, 4, 5, 6, 8, 9, 10, 12, 16};
  for (int b : buckets)
    if (d <= b) return b;
  return -1;
}

// Dispatch a DP-templated launch.
#define DP_SWITCH(dp, ...)                        \
  switch (dp) {                                   \
    case 2: { constexpr int DP = 2; __VA_ARGS__; } break;   \
    case 3: { constexpr int DP = 3; __VA_ARGS__; } break;   \
    case 4: { constexpr int DP = 4; __VA_ARGS__; } break;   \
    case 5: { constexpr int DP = 5; __VA_ARGS__; } break;   \
    case 6: { constexpr int DP = 6; __VA_ARGS__; } break;   \
    case 8: { constexpr int DP = 8; __VA_ARGS__; } break;   \
    case 9: { constexpr int DP = 9; __VA_ARGS__; } break;   \
    case 10: { constexpr int DP = 10; __VA_ARGS__; } break; \
    case 12: { constexpr int DP = 12; __VA_ARGS__; } break; \
    case 16: { constexpr int DP = 16; __VA_ARGS__; } break; \
    default: break;                               \
  }

enum OpKind {
  OP_PREP,
  OP_KBUILD,
  OP_FACTOR_DF,    // persistent dataflow Cholesky + triangular inverse (K generated in-kernel)
  OP_TRMV,
  OP_TRMV_T,
  OP_LAUUM,
  OP_FINALIZE
};

struct Op {
  OpKind kind;
  int task_off = 0, task_cnt = 0;
  int stage = 0;      // 1 params (+ diagnostic kbuild), 2 factorisation, 3 alpha, 4 K^-1 + traces, 5 final
  bool grad_only = false;
  bool debug_only = false;
};

template <class Cfg, int LA, int LB>
cudaError_t set_gemm_attr() {
  return cudaFuncSetAttribute(tile_gemm_kernel<Cfg, LA, LB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              Engine<Cfg, LA, LB>::SMEM_BYTES);
}

}  // namespace

struct eb200_gp {
  int device = 0, n = 0, d = 0, dp = 0, np = 0, nt = 0, nT = 0;
  cudaStream_t stream = nullptr;
  double *A = nullptr, *W = nullptr, *X = nullptr, *noise = nullptr, *r = nullptr, *z = nullptr, *alpha = nullptr;
  double *logdet_part = nullptr, *tpartial = nullptr, *trace_partial = nullptr;
  double *p_dev = nullptr, *out_dev = nullptr;
  KernelParams* kp = nullptr;
  int* info = nullptr;
  double *h_p = nullptr, *h_out = nullptr;  // pinned
  TileTask* tasks = nullptr;
  int2* ktiles = nullptr;
  FactorTask* ftasks = nullptr;  // dataflow task list: Cholesky tasks, then inverse tasks
  int n_ftasks = 0, qbase[5] = {0, 0, 0, 0, 0};
  int* inv_order = nullptr;  // the inverse's tasks (indices into ftasks) in their dependency-safe order
  int n_inv = 0;
  // State of batched value-only evaluations (allocated on first use).  Problem 0 factorises in the handle's
  // own matrix, problems 1.. in matrices of their own; the small per-problem arrays are contiguous so that
  // one memset, one copy and one launch serve the whole batch.
  struct BatchState {
    double* A[MAX_BATCH] = {nullptr};
    double *Xd = nullptr, *logdet = nullptr, *rdiag = nullptr, *out = nullptr, *p_dev = nullptr;  // [MAX_BATCH][..]
    int* zeroed = nullptr;  // [MAX_BATCH] x (flags | counters (8) | info, abort (2)), cleared per batch
    KernelParams* kp = nullptr;
    double* h_stage = nullptr;  // pinned: MAX_BATCH parameter vectors + results
    int slots = 0;             // matrices allocated so far
  } batch;
  // State of grouped multi-handle evaluations led by this handle (allocated on first use): parameter vectors
  // and results of the group's problems, contiguous so that one copy each way serves the whole group.
  struct GroupState {
    double *p_dev = nullptr, *out_dev = nullptr;  // [MAX_BATCH][MAXD + 1], [MAX_BATCH][MAXD + 8]
    double* h_stage = nullptr;                    // pinned: both of the above
  } group;
  size_t state_ints = 0;  // ints behind `ready` zeroed before every factorisation: flags, claim marks, counters, info, abort flag
  FactorTask* ftasks_v = nullptr;  // value-only list: Cholesky tiles + the residual row
  int n_ftasks_v = 0;
  int qbase_v[5] = {0, 0, 0, 0, 0};
  unsigned* counter = nullptr;
  int* ready = nullptr;
  int* xready = nullptr;
  int* abort_flag = nullptr;
  double* rdiag = nullptr;
  unsigned long long* trace = nullptr;  // diagnostics only
  int n_sms = 148;
  int n_ktiles = 0, n_lauum = 0, lauum_full_off = 0, lauum_full_cnt = 0;
  std::vector<Op> ops;
  cudaGraphExec_t g_grad = nullptr, g_lml = nullptr, g_val = nullptr;  // kernels only (device-pointer entry points)
  // the same with the copy of the parameters from h_p and of the result into h_out as graph nodes: one API
  // call per evaluation from host buffers
  cudaGraphExec_t gh_grad = nullptr, gh_lml = nullptr, gh_val = nullptr;
  int launches_grad = 0, launches_lml = 0, launches_val = 0;
  bool factor_valid = false;
  // prediction scratch
  double *mu_dev = nullptr, *var_dev = nullptr, *T_in = nullptr;  // staging of the host-pointer predict (np rows)
  double *h_pred = nullptr;  // pinned staging for predict (np * (d + 2) doubles)
  char *arena = nullptr, *pinned = nullptr;  // one device / one pinned allocation behind the small arrays above
  // grow-only device buffers of the grid prediction (parameter rows, independent values, results)
  double *grid_theta = nullptr, *grid_ind = nullptr, *grid_mu = nullptr, *grid_var = nullptr, *grid_e1 = nullptr;
  size_t grid_theta_cap = 0, grid_ind_cap = 0, grid_out_cap = 0, grid_e1_cap = 0;
  // Grow-only scratch of predictions with more query rows than training rows (sweeps): K* rows, padded query
  // rows, per-tile row sums and the variance GEMM's task list for chunks of up to `rows` query rows.  Smaller
  // predictions use the handle's resident buffers (K* goes to matrix A, dead once X = L^-1 exists).
  struct PredictScratch {
    double *ks = nullptr, *T = nullptr, *rowss = nullptr, *e2 = nullptr;
    TileTask* tasks = nullptr;
    int rows = 0, tasks_mt = -1, tasks_cnt = 0;
    size_t e2_rows = 0;  // parameter rows the E2 table of fused grid predictions holds
    CUtensorMap map_ks;  // TMA view of ks (valid when ks != nullptr)
  } big;
  // TMA views (tma_gemm.cuh, 128-row boxes): of Y = L^-T in matrix W (K^-1 tiles) and of X = L^-1 in matrix A
  // (predictive variance; X is transposed into A, where L is dead by then, before the first variance prediction
  // that follows an evaluation: x_in_a)
  CUtensorMap map_y, map_x;
  bool x_in_a = false;
  double* xdiag = nullptr;  // [nt][64][64] diagonal tiles X_jj = L_jj^-1 (what the tile finalisations multiply with)
};

namespace {

void sort_tasks(std::vector<TileTask>& v, size_t off) {
  std::stable_sort(v.begin() + off, v.end(), [](const TileTask& a, const TileTask& b) { return a.nk > b.nk; });
}

// Build the static launch program for one (np): parameters, dataflow factorisation
// (Cholesky + triangular inverse), alpha, X^T X tiles + gradient traces, final reduction.
void build_program(eb200_gp* h, std::vector<TileTask>& tasks) {
  const int nT = h->nT, np = h->np;
  auto& ops = h->ops;
  {
    Op op;
    op.kind = OP_PREP;
    op.stage = 1;
    ops.push_back(op);
    op.kind = OP_KBUILD;  // diagnostics only: the product path generates K inside the factorisation
    op.debug_only = true;
    ops.push_back(op);
  }
  // ---- Cholesky + triangular inverse: one persistent dataflow kernel ----
  {
    Op op;
    op.stage = 2;
    op.kind = OP_FACTOR_DF;
    ops.push_back(op);
  }
  // ---- z = X r = Y^T r, alpha = X^T z = Y z  (the inverse is kept transposed: Y = L^-T) ----
  {
    Op op;
    op.stage = 3;
    op.kind = OP_TRMV_T;
    ops.push_back(op);
    op.kind = OP_TRMV;
    ops.push_back(op);
  }
  // ---- K^-1 tiles + gradient traces ----
  // Tile (I, J) = sum_k Y[i][k] Y[j][k] contracts over k >= I*128.  A contraction can be cut into K-slices (the trace is
  // linear in K^-1; exactly one slice per tile carries the alpha alpha^T term).  The unsplit list also
  // serves the diagnostics path that stores K^-1.
  {
    // Small and medium problems have few tiles per SM, each with a long serial contraction (a 128x128 tile
    // with 1024 rows keeps one SM busy for 134 us), so the longest tile bounds the kernel.
    // Rule: no task longer than the average load of one SM.  Measured on B200 (whole evaluation, ms,
    // unsliced -> sliced): N = 256 0.197 -> 0.170, 1000 0.538 -> 0.427, 2048 1.059 -> 0.912,
    // 2500 1.327 -> 1.199, 3000 1.635 -> 1.571, 4096 and 8176 unchanged (shorter slices lose to the
    // repeated epilogues there: 3.07 ms at a quarter of the average load).
    int LAUUM_SLICE;  // contraction rows / KC per slice
    {
      long long units = 0;
      for (int J = 0; J < nT; ++J)
        for (int I = J; I < nT; ++I) units += (np - I * 128) / KC;
      const long long per = units / h->n_sms;
      LAUUM_SLICE = (int)std::max(4ll, (per + 1) / 2 * 2);  // even (the main loop takes chunks in pairs)
    }
    size_t off = tasks.size();
    for (int J = 0; J < nT; ++J)
      for (int I = J; I < nT; ++I) {
        const int total = (np - I * 128) / KC;
        const int nslice = (total + LAUUM_SLICE - 1) / LAUUM_SLICE;
        for (int sidx = 0; sidx < nslice; ++sidx) {
          // slice boundaries in pairs of chunks (total is a multiple of 8)
          const int c0 = 2 * (int)((long long)(total / 2) * sidx / nslice), c1 = 2 * (int)((long long)(total / 2) * (sidx + 1) / nslice);
          const int k0 = I * 128 + c0 * KC;
          tasks.push_back({I * 128, J * 128, I * 128, k0, J * 128, k0, c1 - c0, (I == J ? 1 : 0) | (sidx == 0 ? 2 : 0)});
        }
      }
    sort_tasks(tasks, off);
    Op op;
    op.kind = OP_LAUUM;
    op.task_off = (int)off;
    op.task_cnt = (int)(tasks.size() - off);
    op.stage = 4;
    op.grad_only = true;
    ops.push_back(op);
    h->n_lauum = op.task_cnt;
    // unsplit list (diagnostics)
    off = tasks.size();
    for (int J = 0; J < nT; ++J)
      for (int I = J; I < nT; ++I)
        tasks.push_back({I * 128, J * 128, I * 128, I * 128, J * 128, I * 128, (np - I * 128) / KC, (I == J ? 1 : 0) | 2});
    h->lauum_full_off = (int)off;
    h->lauum_full_cnt = (int)(tasks.size() - off);
  }
  {
    Op op;
    op.kind = OP_FINALIZE;
    op.stage = 5;
    ops.push_back(op);
  }
}

// The dataflow kernel's arguments for the handle's own state (slot 0).
void fill_factor_args(eb200_gp* h, bool vo, FactorArgs& fa) {
  const int np = h->np;
  const size_t nf = (size_t)(h->nt + 1) * h->nt;
  fa.A = h->A; fa.W = h->W; fa.ld = np; fa.nt = h->nt; fa.n = h->n; fa.X = h->X; fa.noise = h->noise; fa.kp = h->kp;
  fa.Xd = h->xdiag; fa.xd_tile = (long long)TB * TB; fa.xd_ld = TB;
  fa.value_only = vo ? 1 : 0; fa.r = h->r;
  fa.tasks = vo ? h->ftasks_v : h->ftasks; fa.ntasks = vo ? h->n_ftasks_v : h->n_ftasks;
  for (int q = 0; q < 5; ++q) fa.qbase[q] = vo ? h->qbase_v[q] : h->qbase[q];
  fa.counter = h->counter; fa.ready = h->ready; fa.xready = fa.ready + nf; fa.pready = fa.xready + nf;
  fa.cready = fa.pready + nf; fa.lready = fa.cready + nf; fa.sm_chain = fa.lready + 8 * h->nt; fa.claimed = fa.sm_chain + 512;
  fa.inv_order = h->inv_order; fa.n_inv = vo ? 0 : h->n_inv;
  fa.logdet_part = h->logdet_part; fa.rdiag = h->rdiag; fa.info = h->info; fa.abort_flag = h->abort_flag;
  fa.trace = h->trace;
  fa.n_sms = h->n_sms;
  fa.n_chain_sms = 1;
}

// The pointers of the stages after the factorisation for the handle's own state.
void fill_post_args(eb200_gp* h, int mode, PostArgs& pa) {
  pa.W = h->W; pa.r = h->r; pa.z = h->z; pa.tpartial = h->tpartial; pa.alpha = h->alpha; pa.Xtr = h->X; pa.kp = h->kp;
  pa.trace_partial = h->trace_partial; pa.logdet_part = h->logdet_part;
  // value-only: z = L^-1 r is the residual row of the factorised matrix
  pa.zfin = mode == EB200_EVAL_VALUE_ONLY ? h->A + (size_t)h->np * h->np : h->z;
  pa.info = h->info; pa.abort_flag = h->abort_flag; pa.out = h->out_dev;
}

// Flags, claim marks, queue counters, info and the watchdog's abort flag of one handle: one memset.
cudaError_t reset_state(eb200_gp* h, cudaStream_t st) { return cudaMemsetAsync(h->ready, 0, h->state_ints * sizeof(int), st); }

// Launch the kernels of one post-factorisation op for the problems of `pb` (all shaped like h).
// (`group`: the handles of the problems of pb, or nullptr for the single problem h)
int launch_post(eb200_gp* h, const Op& op, const PostBatch& pb, cudaStream_t st, int mode, double* kinv_out,
                eb200_gp* const* group = nullptr) {
  const int want_grad = mode == EB200_EVAL_GRAD;
  const int np = h->np, nb = pb.nb;
  switch (op.kind) {
    case OP_TRMV_T:  // z = Y^T r
      trmv_t_partial_kernel<<<dim3(h->nT, h->nT, nb), 128, 0, st>>>(pb, np, np);
      trmv_t_reduce_kernel<<<dim3(h->nT, nb), 128, 0, st>>>(pb, np);
      return 2;
    case OP_TRMV:  // alpha = Y z
      trmv_upper_kernel<<<dim3(np / 8, nb), 256, 0, st>>>(pb, np, np);
      return 1;
    case OP_LAUUM: {
      // the diagnostics path (K^-1 stored) needs whole tiles; it has no more tasks than the sliced list
      const int toff = kinv_out ? h->lauum_full_off : op.task_off;
      const int tcnt = kinv_out ? h->lauum_full_cnt : op.task_cnt;
      if (kinv_out) cudaMemsetAsync(h->trace_partial, 0, (size_t)h->n_lauum * (h->dp + 1) * sizeof(double), st);
      LauumMaps maps;
      for (int b = 0; b < nb; ++b) maps.y[b] = group ? group[b]->map_y : h->map_y;
      DP_SWITCH(h->dp, (lauum_trace_kernel<DP><<<tcnt * nb, TMA_THREADS, LauumSmem<DP>::BYTES, st>>>(h->tasks + toff, pb, maps, np, h->n, kinv_out)));
      return 1;
    }
    case OP_FINALIZE:
      finalize_kernel<<<nb, 256, 0, st>>>(pb, h->nt, h->n, h->n_lauum, h->dp, h->d, want_grad);
      return 1;
    default:
      return 0;
  }
}

// Enqueue one op.  Returns number of kernel launches issued.
int launch_op(eb200_gp* h, const Op& op, cudaStream_t st, int mode, double* kinv_out) {
  const int np = h->np;
  switch (op.kind) {
    case OP_PREP:
      prep_params_kernel<<<1, 32, 0, st>>>(h->p_dev, h->d, h->kp);
      return 1;
    case OP_KBUILD:
      DP_SWITCH(h->dp, (kbuild_kernel<DP><<<h->n_ktiles, 256, 0, st>>>(h->ktiles, h->X, h->noise, h->kp, h->n, np, h->A)));
      return 1;
    case OP_FACTOR_DF: {
      const bool vo = mode == EB200_EVAL_VALUE_ONLY;
      reset_state(h, st);
      FactorBatch fb;
      fb.nb = 1;
      FactorArgs& fa = fb.p[0];
      fill_factor_args(h, vo, fa);
      // two CTAs per SM: the first wave (one per SM) serves the Cholesky queue, the second the inverse
      const int grid = std::min(2 * h->n_sms, fa.ntasks);
      // at least one CTA always serves the chain queue, never so many that the other queues starve
      fa.n_chain_sms = std::max(1, std::min({6, h->nt / 8, grid / 4}));
      DP_SWITCH(h->dp, (factor_dataflow_kernel<DP, false><<<grid, 256, FactorSmem<DP>::BYTES, st>>>(fb)));
      return 1;
    }
    default: {
      PostBatch pb;
      pb.nb = 1;
      fill_post_args(h, mode, pb.p[0]);
      return launch_post(h, op, pb, st, mode, kinv_out);
    }
  }
}

int enqueue_eval(eb200_gp* h, cudaStream_t st, int mode, int max_stage, double* kinv_out, int* launches) {
  int cnt = 0;
  for (const Op& op : h->ops) {
    if (op.stage > max_stage) continue;
    if (op.grad_only && mode != EB200_EVAL_GRAD) continue;
    if (op.debug_only && max_stage != 1) continue;
    if (mode == EB200_EVAL_VALUE_ONLY && op.stage == 3) continue;  // no inverse, no alpha
    cnt += launch_op(h, op, st, mode, kinv_out);
  }
  CK(cudaGetLastError());
  if (launches) *launches = cnt;
  return 0;
}

int capture_graph(eb200_gp* h, int mode, cudaGraphExec_t* exec, int* launches, bool host_copies = false) {
  CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  if (host_copies) cudaMemcpyAsync(h->p_dev, h->h_p, (h->d + 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  int rc = enqueue_eval(h, h->stream, mode, 99, nullptr, launches);
  if (host_copies)
    cudaMemcpyAsync(h->h_out, h->out_dev, EB200_RESULT_LEN(h->d) * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
  if (rc || e != cudaSuccess) {
    if (graph) cudaGraphDestroy(graph);
    if (rc) return rc;
    CK(e);
  }
  e = cudaGraphInstantiate(exec, graph, 0);
  cudaGraphDestroy(graph);
  CK(e);
  return 0;
}

int set_attrs_impl();
// Function attributes are per device and set once per device, whichever host thread gets there first.
int set_attrs() {
  static std::mutex mu;
  static bool done_dev[64] = {false};
  int dev = 0;
  CK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (done_dev[dev & 63]) return 0;
  if (set_attrs_impl()) return 1;
  done_dev[dev & 63] = true;
  return 0;
}
int set_attrs_impl() {
  CK((set_gemm_attr<Cfg128, KCONTIG, KCONTIG>()));
  CK(cudaFuncSetAttribute(var_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          Engine<Cfg128, KCONTIG, KCONTIG>::SMEM_BYTES));
  CK(cudaFuncSetAttribute(var_gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_BYTES));
  CK(cudaFuncSetAttribute(tile_gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_BYTES));
  CK(cudaFuncSetAttribute(var_grid_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_BYTES));
#define SET_POTRF(DPV)                                                                                       \
  CK(cudaFuncSetAttribute(factor_dataflow_kernel<DPV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                          FactorSmem<DPV>::BYTES));                                                        \
  CK(cudaFuncSetAttribute(factor_dataflow_kernel<DPV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                          FactorSmem<DPV>::BYTES));
  SET_POTRF(2) SET_POTRF(3) SET_POTRF(4) SET_POTRF(5) SET_POTRF(6) SET_POTRF(8) SET_POTRF(9) SET_POTRF(10)
  SET_POTRF(12) SET_POTRF(16)
#undef SET_POTRF
#define SET_LAUUM(DPV)                                                                                       \
  CK(cudaFuncSetAttribute(lauum_trace_kernel<DPV>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                          LauumSmem<DPV>::BYTES));
  SET_LAUUM(2) SET_LAUUM(3) SET_LAUUM(4) SET_LAUUM(5) SET_LAUUM(6) SET_LAUUM(8) SET_LAUUM(9) SET_LAUUM(10)
  SET_LAUUM(12) SET_LAUUM(16)
#undef SET_LAUUM
  return 0;
}

__global__ void pad_rows_kernel(const double* __restrict__ src, int m, int d, int dp, double* __restrict__ dst) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)m * dp) return;
  int c = (int)(e % dp);
  size_t r = e / dp;
  dst[e] = (c < d) ? src[r * d + c] : 0.0;
}

// Query rows of a prediction grid, built on the device: row q = (independent[q % n_ind], theta[q / n_ind][:]),
// optionally rounded to float32 first (the reference's emulators cast their query points to float32,
// gaussian_process.py:277-279), written zero-padded to dp columns.
__global__ void grid_rows_kernel(const double* __restrict__ theta, const double* __restrict__ ind, size_t q0, int mc,
                                 int n_ind, int d, int dp, int round_f32, double* __restrict__ dst) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)mc * dp) return;
  const int c = (int)(e % dp);
  const size_t q = q0 + e / dp;
  double v = 0.0;
  if (c == 0)
    v = ind[q % n_ind];
  else if (c < d)
    v = theta[(q / n_ind) * (size_t)(d - 1) + (c - 1)];
  if (round_f32) v = (double)(float)v;
  dst[e] = v;
}

int grow(double** buf, size_t* cap, size_t need) {
  if (*cap >= need) return 0;
  if (*buf) CK(cudaFree(*buf));
  *buf = nullptr;
  *cap = 0;
  CK(cudaMalloc(buf, need * sizeof(double)));
  *cap = need;
  return 0;
}

// TMA descriptor of a row-major fp64 matrix [rows][cols] with leading dimension ld: boxes of 128 rows x 16
// columns (one operand slab of tma_gemm.cuh), 128-byte swizzle.  The driver entry point is looked up once.
int make_tensor_map(CUtensorMap* out, const double* base, size_t rows, size_t cols, size_t ld, unsigned box_rows = 128) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!encode) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult status;
      CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &status));
      if (status != cudaDriverEntryPointSuccess || !fn) return fail("cuTensorMapEncodeTiled is not available in this driver");
      encode = reinterpret_cast<EncodeFn>(fn);
    }
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)box_rows};
  const cuuint32_t elem[2] = {1u, 1u};
  const CUresult rc = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, elem,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with code " + std::to_string((int)rc));
  return 0;
}

// Task list of the variance GEMM for mt 128-row blocks of query rows: tile (M, I) contracts over k < (I+1)*128.
// Blocks of query rows are taken in groups of VAR_GROUP (their K* rows, 4 MB per block at N = 4096, stay in L2
// while the group's tiles read them), longest contraction first within a group.
int build_var_tasks(eb200_gp* h, int mt, TileTask* dst, int* cnt, cudaStream_t st) {
  static const int group = std::max(1, env_int("EB200_VAR_GROUP", 1 << 20));
  std::vector<TileTask> v;
  for (int M0 = 0; M0 < mt; M0 += group) {
    const size_t off = v.size();
    for (int M = M0; M < std::min(mt, M0 + group); ++M)
      for (int I = 0; I < h->nT; ++I) v.push_back({M * 128, I * 128, M * 128, 0, I * 128, 0, (I + 1) * 128 / KC, 0});
    sort_tasks(v, off);
  }
  CK(cudaMemcpyAsync(dst, v.data(), v.size() * sizeof(TileTask), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));  // v goes out of scope
  *cnt = (int)v.size();
  return 0;
}

// Largest number of query rows one pass of the prediction kernels takes: everything up to EB200_PREDICT_CHUNK rows
// (default 16384: the variance GEMM's launch tail is amortised), in multiples of 128.
int predict_chunk_rows(const eb200_gp* h, size_t m) {
  static const int max_rows = std::max(128, env_int("EB200_PREDICT_CHUNK", 16384) / 128 * 128);
  return (int)std::min<size_t>((std::max<size_t>(m, 1) + 127) / 128 * 128, (size_t)max_rows);
}

// Grow-only scratch for chunks of `rows` query rows: padded query rows, and for the variance the K* rows, per-tile
// row sums and the GEMM's task list.
int ensure_predict_scratch(eb200_gp* h, int rows, int want_var, bool need_ks = true) {
  eb200_gp::PredictScratch& ps = h->big;
  if (ps.rows < rows || (want_var && !ps.rowss)) {
    rows = std::max(rows, ps.rows);
    for (void* q : {(void*)ps.ks, (void*)ps.T, (void*)ps.rowss, (void*)ps.tasks})
      if (q) CK(cudaFree(q));
    double* keep_e2 = ps.e2;
    const size_t keep_e2_rows = ps.e2_rows;
    ps = eb200_gp::PredictScratch();
    ps.e2 = keep_e2;
    ps.e2_rows = keep_e2_rows;
    CK(cudaMalloc(&ps.T, (size_t)rows * h->dp * sizeof(double)));
    if (want_var) {
      CK(cudaMalloc(&ps.rowss, (size_t)h->nT * rows * sizeof(double)));
      CK(cudaMalloc(&ps.tasks, (size_t)(rows / 128) * h->nT * sizeof(TileTask)));
    }
    ps.rows = rows;
  }
  if (want_var && need_ks && !ps.ks) {  // K* rows: only predictions that are not fused grids need them
    CK(cudaMalloc(&ps.ks, (size_t)ps.rows * h->np * sizeof(double)));
    if (make_tensor_map(&ps.map_ks, ps.ks, ps.rows, h->np, h->np)) return 1;
  }
  return 0;
}

// (grid: the chunk is rows [grid->q0, grid->q0 + mc) of a prediction grid, whose K* rows come from the factorised
//  form of the kernel -- predict_grid_mean_kernel -- instead of from explicit query rows)
struct GridChunk {
  const double* theta;
  const double* e1;
  size_t q0;
  int n_ind, round_f32;
  bool fused;  // the variance kernel rebuilds K* = E1 * E2 in shared memory (no K* rows in global memory)
};

// Predict one chunk (mc <= chunk_rows query rows) whose padded coordinates already sit in the scratch's row buffer (big.T).
int predict_chunk(eb200_gp* h, int chunk_rows, int mc, int want_var, double* mu_dev, double* var_dev, cudaStream_t st,
                  const GridChunk* grid = nullptr) {
  const int np = h->np;
  const int blocks = (mc + 31) / 32;
  const bool fused = grid && grid->fused && want_var;
  double* ks = (want_var && !fused) ? h->big.ks : nullptr;
  if (grid) {
    const size_t r_first = grid->q0 / grid->n_ind, r_last = (grid->q0 + mc - 1) / grid->n_ind;
    const unsigned gblocks = (unsigned)(r_last - r_first + 1);  // one CTA per parameter row
    if (fused && h->big.e2_rows < gblocks) {
      if (h->big.e2) CK(cudaFree(h->big.e2));
      h->big.e2 = nullptr;
      h->big.e2_rows = 0;
      CK(cudaMalloc(&h->big.e2, (size_t)gblocks * np * sizeof(double)));
      h->big.e2_rows = gblocks;
    }
    DP_SWITCH(h->dp, (predict_grid_mean_kernel<DP><<<gblocks, 256, 0, st>>>(grid->theta, grid->q0, mc, grid->n_ind, h->d, grid->round_f32,
                                                                            grid->e1, h->X, h->alpha, h->kp, h->n, np, mu_dev, ks, np,
                                                                            fused ? h->big.e2 : nullptr)));
  } else {
    DP_SWITCH(h->dp, (predict_mean_kernel<DP><<<blocks, 256, 0, st>>>(h->big.T, mc, h->X, h->alpha, h->kp, h->n, np, mu_dev, ks, np)));
  }
  if (want_var) {
    if (!h->x_in_a) {
      // the variance contracts K* with the ROWS of X = L^-1; the evaluation leaves Y = X^T (k-contiguous for its own
      // contractions), so X is written out once per factorisation -- into matrix A, where L is no longer needed
      transpose_upper_kernel<<<dim3(np / 32, np / 32), 256, 0, st>>>(h->W, np, h->A);
      h->x_in_a = true;
    }
    const int mt = (mc + 127) / 128;
    if (h->big.tasks_mt != mt) {
      if (build_var_tasks(h, mt, h->big.tasks, &h->big.tasks_cnt, st)) return 1;
      h->big.tasks_mt = mt;
    }
    const int cnt = h->big.tasks_cnt, ld_rowss = h->big.rows;
    static const int use_tma = env_int("EB200_VAR_TMA", 1);
    if (fused)
      var_grid_fused_kernel<<<cnt, FUSED_THREADS, TMA_SMEM_BYTES, st>>>(h->big.tasks, h->map_x, grid->e1, h->big.e2, np, grid->q0, mc,
                                                                        grid->n_ind, h->big.rowss, ld_rowss);
    else if (use_tma)
      var_gemm_tma_kernel<<<cnt, TMA_THREADS, TMA_SMEM_BYTES, st>>>(h->big.tasks, h->big.map_ks, h->map_x, h->big.rowss, ld_rowss);
    else
      var_gemm_kernel<<<cnt, 256, Engine<Cfg128, KCONTIG, KCONTIG>::SMEM_BYTES, st>>>(h->big.tasks, ks, np, h->A, np, h->big.rowss, ld_rowss);
    var_finish_kernel<<<(mc + 255) / 256, 256, 0, st>>>(h->big.rowss, ld_rowss, h->nT, h->kp, mc, var_dev);
  }
  CK(cudaGetLastError());
  return 0;
}

// Grid prediction with everything in device memory (see eb200_gp_predict_grid_host).
int predict_grid_impl(eb200_gp* h, size_t n_theta, const double* theta_dev, int n_ind, const double* ind_dev, int round_f32,
                      int want_var, double* mu_dev, double* var_dev, cudaStream_t st) {
  const size_t m = n_theta * (size_t)n_ind;
  const int chunk = predict_chunk_rows(h, m);
  static const int factorised_env = env_int("EB200_GRID_FACTORISED", 1), fused_env = env_int("EB200_GRID_FUSED", 1);
  const bool fused = factorised_env && fused_env && want_var;
  if (ensure_predict_scratch(h, chunk, want_var, !fused)) return 1;
  // The kernel factorises over its axes: exp(-(ind_b - x_j0)^2 / 2 M_0) is a table of n_ind x N entries shared by all
  // parameter rows, so the grid costs one exponential per (parameter row, training point) instead of one per entry.
  static const int factorised = env_int("EB200_GRID_FACTORISED", 1);
  if (factorised) {
    if (grow(&h->grid_e1, &h->grid_e1_cap, (size_t)n_ind * h->np)) return 1;
    const size_t tot = (size_t)n_ind * h->np;
    grid_axis_table_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ind_dev, n_ind, h->X, h->dp, h->n, h->np, h->kp, round_f32, h->grid_e1);
    for (size_t q0 = 0; q0 < m; q0 += chunk) {
      const int mc = (int)std::min<size_t>(chunk, m - q0);
      const GridChunk gc{theta_dev, h->grid_e1, q0, n_ind, round_f32, fused};
      if (predict_chunk(h, chunk, mc, want_var, mu_dev + q0, want_var ? var_dev + q0 : nullptr, st, &gc)) return 1;
    }
    return 0;
  }
  double* rows = h->big.T;
  for (size_t q0 = 0; q0 < m; q0 += chunk) {
    const int mc = (int)std::min<size_t>(chunk, m - q0);
    const size_t tot = (size_t)mc * h->dp;
    grid_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(theta_dev, ind_dev, q0, mc, n_ind, h->d, h->dp, round_f32, rows);
    if (predict_chunk(h, chunk, mc, want_var, mu_dev + q0, want_var ? var_dev + q0 : nullptr, st)) return 1;
  }
  return 0;
}

}  // namespace

extern "C" {

int eb200_version(void) { return 100; }
const char* eb200_last_error(void) { return g_err.c_str(); }
int eb200_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
  return c;
}

int eb200_device_mem_info(int device, unsigned long long* free_bytes, unsigned long long* total_bytes) {
  if (!free_bytes || !total_bytes) return fail("device_mem_info: null argument");
  CK(cudaSetDevice(device));
  size_t f = 0, t = 0;
  CK(cudaMemGetInfo(&f, &t));
  *free_bytes = f;
  *total_bytes = t;
  return 0;
}

static int create_impl(eb200_gp* h, int device, int n, int d, const double* x_host, const double* noise_host);

int eb200_gp_create(eb200_gp** out, int device, int n, int d, const double* x_host, const double* noise_host) {
  if (!out || !x_host || !noise_host) return fail("eb200_gp_create: null argument");
  *out = nullptr;
  if (n < 1) return fail("eb200_gp_create: n must be >= 1");
  if (d < 1 || d > EB200_MAX_DIM) return fail("eb200_gp_create: kernel dimension must be in [1, 16]");
  CK(cudaSetDevice(device));
  if (set_attrs()) return 1;
  eb200_gp* h = new eb200_gp();
  if (create_impl(h, device, n, d, x_host, noise_host)) {
    // (typically cudaMalloc out of memory with many live handles: give back whatever this one had got, and
    //  clear the sticky last-error so that the next call on this thread does not report it again)
    const std::string why = g_err;
    eb200_gp_destroy(h);
    cudaGetLastError();
    return fail(why);
  }
  *out = h;
  return 0;
}

static int create_impl(eb200_gp* h, int device, int n, int d, const double* x_host, const double* noise_host) {
  h->device = device;
  h->n = n;
  h->d = d;
  h->dp = pad_dim(d);
  h->np = ((n + 127) / 128) * 128;
  h->nt = h->np / TB;
  h->nT = h->np / 128;
  const int np = h->np, dp = h->dp;
  const size_t mat = (size_t)np * np * sizeof(double);
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  const size_t mat_a = mat + (size_t)TB * np * sizeof(double);  // + the residual tile row of value-only evaluations
  CK(cudaMalloc(&h->A, mat_a));
  CK(cudaMalloc(&h->W, mat));
  // Everything small lives in one device arena and one pinned block: a handle used to cost ~30 cudaMalloc
  // and as many (synchronising) cudaFree calls, which is what creating and destroying the 64 per-bin GPs
  // of a binned emulator, or re-creating a released handle, mostly paid for.
  {
    struct Want { void** slot; size_t bytes; };
    std::vector<Want> dev_wants = {
        {(void**)&h->X, (size_t)np * dp * sizeof(double)},
        {(void**)&h->noise, np * sizeof(double)},
        {(void**)&h->r, np * sizeof(double)},
        {(void**)&h->z, np * sizeof(double)},
        {(void**)&h->alpha, np * sizeof(double)},
        {(void**)&h->logdet_part, h->nt * sizeof(double)},
        {(void**)&h->tpartial, (size_t)h->nT * np * sizeof(double)},
        {(void**)&h->p_dev, (MAXD + 1) * sizeof(double)},
        {(void**)&h->out_dev, (MAXD + 8) * sizeof(double)},
        {(void**)&h->kp, sizeof(KernelParams)},
        {(void**)&h->T_in, (size_t)np * d * sizeof(double)},
        {(void**)&h->mu_dev, np * sizeof(double)},
        {(void**)&h->var_dev, np * sizeof(double)},
        {(void**)&h->rdiag, (size_t)np * sizeof(double)},
    };
    auto align = [](size_t b) { return (b + 255) & ~(size_t)255; };
    size_t total = 0;
    for (const Want& w : dev_wants) total += align(w.bytes);
    CK(cudaMalloc(&h->arena, total));
    CK(cudaMemset(h->arena, 0, total));  // (r, z, alpha and the flags start at zero)
    size_t off = 0;
    for (const Want& w : dev_wants) {
      *w.slot = h->arena + off;
      off += align(w.bytes);
    }
    const size_t pinned = align((MAXD + 1) * sizeof(double)) + align((MAXD + 8) * sizeof(double)) +
                          (size_t)np * (d + 2) * sizeof(double);
    CK(cudaMallocHost(&h->pinned, pinned));
    h->h_p = reinterpret_cast<double*>(h->pinned);
    h->h_out = reinterpret_cast<double*>(h->pinned + align((MAXD + 1) * sizeof(double)));
    h->h_pred = reinterpret_cast<double*>(h->pinned + align((MAXD + 1) * sizeof(double)) + align((MAXD + 8) * sizeof(double)));
  }
  CK(cudaMemset(h->W, 0, mat));    // (Y is upper triangular: the strictly-lower 64-blocks of its diagonal 128-tiles must stay zero,
  CK(cudaMemset(h->A, 0, mat_a));  //  as the strictly-upper ones of L)
  CK(cudaMalloc(&h->xdiag, (size_t)h->nt * TB * TB * sizeof(double)));
  if (make_tensor_map(&h->map_y, h->W, np, np, np) || make_tensor_map(&h->map_x, h->A, np, np, np)) return 1;
  CK(cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, device));
  {
    // Queue 0: the Cholesky's chain tasks (diagonal tile, first sub-diagonal tile, inverse of the diagonal
    // tile) in order.  Queue 1: the other Cholesky tiles column by column plus the background
    // pre-accumulation tasks.  Queues 2 and 3: the inverse's tiles row by row and their pre-accumulation
    // tasks, claimed ready-first.
    //
    // A Cholesky tile (i, c) with more than POTRF_LA + POTRF_MINP tile products is split: a background
    // task listed right after column c - POTRF_LA - 1 (when its inputs exist) pre-accumulates tile
    // columns [0, c - POTRF_LA); the tile's own task adds the last POTRF_LA and finalises.  Near the
    // diagonal those own tasks sit on (or a few steps off) the dependency chain, so bounding their
    // work keeps the chain from stalling on a late-claimed tile.
    std::vector<FactorTask> ft;
    const int nt = h->nt;
    const int POTRF_LA = 8, POTRF_MINP = 4;  // (a sweep of LA 6..12, MINP 2..8 on B200 moves the evaluation by < 1 %)
    auto split = [&](int c) { return c > POTRF_LA + POTRF_MINP; };
    // The chain queue holds the band i - c <= CHAIN_BAND: tile (c + d, c) is needed d steps after the
    // diagonal tile of its column, so the tiles next to the diagonal have almost no slack either.
    const int CHAIN_BAND = 1;
    for (int c = 0; c < nt; ++c) {
      for (int i = c; i < std::min(nt, c + CHAIN_BAND + 1); ++i) ft.push_back({TASK_POTRF, i, c, split(c) ? c - POTRF_LA : 0});
      ft.push_back({TASK_XDIAG, c, c, 0});
    }
    h->qbase[1] = (int)ft.size();
    for (int c = 0; c < nt; ++c) {
      for (int i = c + CHAIN_BAND + 1; i < nt; ++i) ft.push_back({TASK_POTRF, i, c, split(c) ? c - POTRF_LA : 0});
      const int cc = c + POTRF_LA + 1;  // column whose tiles can now be pre-accumulated up to column c
      if (cc < nt && split(cc))
        for (int i = cc; i < nt; ++i) ft.push_back({TASK_POTRF_PART, i, cc, cc - POTRF_LA});
    }
    h->qbase[2] = (int)ft.size();
    // Inverse tiles row by row.  A tile with more than MAIN_K + MIN_PART tile products is split: a
    // helper task (listed as soon as its inputs exist, i.e. after row k-1) pre-accumulates tile
    // columns [J, k), the tile's own task adds the last MAIN_K and finalises.  This moves most of
    // the work of the last rows -- the tail of the launch -- into the time the Cholesky chain
    // leaves SMs idle.
    // (Measured: cutting a tile's helper into chained chunks of 8 tile columns, each claimable as soon as
    //  its inputs exist, fills the starved first 300 us but makes the chunks wait on one another later:
    //  3.11 ms per evaluation against 3.00 ms with one helper per tile.)
    const int MAIN_K = 12, MIN_PART = 8, INV_CHUNK = 1 << 20;  // (MAIN_K 8..16, MIN_PART 4..12: < 1.5 %)
    // The dependency-safe total order is row by row, a row's tiles followed by the helpers whose inputs
    // exist once that row is done; `seq` records it.  Tiles go to queue 2 and helpers to queue 3, both
    // claimed ready-first (factor_dataflow.cuh), falling back to this order.
    std::vector<std::vector<FactorTask>> after_row(nt);
    for (int r = 1; r < nt; ++r)
      for (int J = 0; J < r; ++J)
        if (r - J > MAIN_K + MIN_PART) {
          // helper chunks [c0, c1) of INV_CHUNK tile columns, each listed once row c1 - 1 of X exists
          const int end = r - MAIN_K;
          for (int c0 = J; c0 < end;) {
            int c1 = std::min(c0 + INV_CHUNK, end);
            if (end - c1 < INV_CHUNK / 2) c1 = end;  // no tiny last chunk
            after_row[c1 - 1].push_back({TASK_TRTRI_PART, r, J, c1, 0, c0});
            c0 = c1;
          }
        }
    std::vector<FactorTask> mains, helpers;
    int seq = 0;
    for (int r = 1; r < nt; ++r) {
      for (int J = 0; J < r; ++J) mains.push_back({TASK_TRTRI, r, J, (r - J > MAIN_K + MIN_PART) ? r - MAIN_K : J, seq++, 0});
      for (auto& tk : after_row[r]) {
        tk.seq = seq++;
        helpers.push_back(tk);
      }
    }
    std::vector<int> order(seq);
    for (size_t k = 0; k < mains.size(); ++k) order[mains[k].seq] = (int)(ft.size() + k);
    ft.insert(ft.end(), mains.begin(), mains.end());
    h->qbase[3] = (int)ft.size();
    for (size_t k = 0; k < helpers.size(); ++k) order[helpers[k].seq] = (int)(ft.size() + k);
    ft.insert(ft.end(), helpers.begin(), helpers.end());
    h->n_inv = seq;
    CK(cudaMalloc(&h->inv_order, std::max<size_t>(1, order.size()) * sizeof(int)));
    CK(cudaMemcpy(h->inv_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice));
    h->n_ftasks = (int)ft.size();
    h->qbase[4] = h->n_ftasks;
    {
      // flags + claim marks, then the queue counters (8), info and the watchdog's abort flag: everything one
      // evaluation starts from zero sits in one block cleared by one memset (reset_state)
      const size_t flag_ints = (size_t)4 * (h->nt + 1) * h->nt + 8 * h->nt + 512 + ft.size();
      h->state_ints = flag_ints + 8 + 2;
      CK(cudaMalloc(&h->ready, h->state_ints * sizeof(int)));
      CK(cudaMemset(h->ready, 0, h->state_ints * sizeof(int)));
      h->counter = reinterpret_cast<unsigned*>(h->ready + flag_ints);
      h->info = h->ready + flag_ints + 8;
      h->abort_flag = h->info + 1;
    }
    CK(cudaMalloc(&h->ftasks, ft.size() * sizeof(FactorTask)));
    CK(cudaMemcpy(h->ftasks, ft.data(), ft.size() * sizeof(FactorTask), cudaMemcpyHostToDevice));
    // Value-only list (likelihood without gradient, e.g. ensemble sampling): the Cholesky tasks of
    // queues 0 and 1 with one more tile row, nt, that carries the residual (factor_dataflow.cuh), and
    // no inverse.  (Measured: carrying the residual row in the full evaluation as well costs the
    // factorisation 44 us and saves the 15 us z = X r kernel, so the full path keeps that kernel.)
    std::vector<FactorTask> fv(ft.begin(), ft.begin() + h->qbase[1]);
    h->qbase_v[1] = (int)fv.size();
    for (int c = 0; c < nt; ++c) {
      // (the residual row's tile is never a chain-queue tile, also where it falls inside the band)
      for (int i = std::min(c + CHAIN_BAND + 1, nt); i <= nt; ++i) fv.push_back({TASK_POTRF, i, c, split(c) ? c - POTRF_LA : 0});
      const int cc = c + POTRF_LA + 1;
      if (cc < nt && split(cc))
        for (int i = cc; i <= nt; ++i) fv.push_back({TASK_POTRF_PART, i, cc, cc - POTRF_LA});
    }
    h->n_ftasks_v = (int)fv.size();
    h->qbase_v[2] = h->qbase_v[3] = h->qbase_v[4] = h->n_ftasks_v;
    CK(cudaMalloc(&h->ftasks_v, fv.size() * sizeof(FactorTask)));
    CK(cudaMemcpy(h->ftasks_v, fv.data(), fv.size() * sizeof(FactorTask), cudaMemcpyHostToDevice));
  }

  // inputs, padded
  {
    std::vector<double> xp((size_t)np * dp, 0.0), nz(np, 0.0);
    for (int i = 0; i < n; ++i) {
      for (int m = 0; m < d; ++m) xp[(size_t)i * dp + m] = x_host[(size_t)i * d + m];
      nz[i] = noise_host[i];
    }
    CK(cudaMemcpy(h->X, xp.data(), xp.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->noise, nz.data(), nz.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  // kbuild tile list: lower 64-tiles + zero-fill of strictly-upper halves of diagonal 128-tiles
  {
    std::vector<int2> kt;
    for (int i = 0; i < h->nt; ++i)
      for (int j = 0; j <= i; ++j) kt.push_back(make_int2(i, j));
    for (int b = 0; b < h->nT; ++b) kt.push_back(make_int2(2 * b, 2 * b + 1));
    h->n_ktiles = (int)kt.size();
    CK(cudaMalloc(&h->ktiles, kt.size() * sizeof(int2)));
    CK(cudaMemcpy(h->ktiles, kt.data(), kt.size() * sizeof(int2), cudaMemcpyHostToDevice));
  }
  {
    std::vector<TileTask> tasks;
    build_program(h, tasks);
    CK(cudaMalloc(&h->tasks, tasks.size() * sizeof(TileTask)));
    CK(cudaMemcpy(h->tasks, tasks.data(), tasks.size() * sizeof(TileTask), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->trace_partial, (size_t)std::max(1, h->n_lauum) * (dp + 1) * sizeof(double)));
  }
  // (the six evaluation graphs -- three modes, with and without the host copies -- are captured on first
  //  use: a handle typically needs one or two of them, and capture + instantiation is most of what
  //  creating a small handle costs)
  return 0;
}

int eb200_gp_destroy(eb200_gp* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->g_grad) cudaGraphExecDestroy(h->g_grad);
  if (h->g_lml) cudaGraphExecDestroy(h->g_lml);
  if (h->g_val) cudaGraphExecDestroy(h->g_val);
  for (cudaGraphExec_t g : {h->gh_grad, h->gh_lml, h->gh_val})
    if (g) cudaGraphExecDestroy(g);
  void* dev[] = {h->A, h->W, h->xdiag, h->arena, h->trace_partial, h->tasks, h->ktiles, h->grid_theta, h->grid_ind, h->grid_mu, h->grid_var, h->grid_e1,
                 h->ftasks, h->ftasks_v, h->inv_order, h->ready};
  for (void* p : dev)
    if (p) cudaFree(p);
  {
    eb200_gp::BatchState& bs = h->batch;
    for (int b = 1; b < bs.slots; ++b) cudaFree(bs.A[b]);
    void* own[] = {bs.Xd, bs.logdet, bs.rdiag, bs.out, bs.p_dev, bs.zeroed, bs.kp};
    for (void* q : own)
      if (q) cudaFree(q);
    if (bs.h_stage) cudaFreeHost(bs.h_stage);
  }
  for (void* q : {(void*)h->big.ks, (void*)h->big.T, (void*)h->big.rowss, (void*)h->big.tasks, (void*)h->big.e2})
    if (q) cudaFree(q);
  if (h->group.p_dev) cudaFree(h->group.p_dev);
  if (h->group.out_dev) cudaFree(h->group.out_dev);
  if (h->group.h_stage) cudaFreeHost(h->group.h_stage);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int eb200_gp_n(const eb200_gp* h) { return h ? h->n : -1; }
int eb200_gp_dim(const eb200_gp* h) { return h ? h->d : -1; }
int eb200_gp_padded_n(const eb200_gp* h) { return h ? h->np : -1; }
int eb200_gp_eval_launches(const eb200_gp* h, int mode) {
  if (!h) return -1;
  int cnt = 0;  // kernel launches of one evaluation = kernel nodes of its graph
  for (const Op& op : h->ops) {
    if (op.debug_only || (op.grad_only && mode != EB200_EVAL_GRAD)) continue;
    if (mode == EB200_EVAL_VALUE_ONLY && op.stage == 3) continue;
    cnt += (op.kind == OP_TRMV_T) ? 2 : 1;
  }
  return cnt;
}

int eb200_gp_set_residual_host(eb200_gp* h, const double* r_host) {
  if (!h || !r_host) return fail("set_residual: null argument");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  memcpy(h->h_pred, r_host, h->n * sizeof(double));
  CK(cudaMemcpyAsync(h->r, h->h_pred, h->n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->factor_valid = false;
  h->x_in_a = false;  // (matrix A holds L again)
  return 0;
}

int eb200_gp_set_residual_device(eb200_gp* h, const double* r_dev, void* stream) {
  if (!h || !r_dev) return fail("set_residual: null argument");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  CK(cudaMemcpyAsync(h->r, r_dev, h->n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  h->factor_valid = false;
  h->x_in_a = false;  // (matrix A holds L again)
  return 0;
}

// The evaluation graph of a mode, captured on first use (nullptr + error message on failure).
static cudaGraphExec_t eval_graph(eb200_gp* h, int mode, bool host_copies = false) {
  cudaGraphExec_t* slot = host_copies ? (mode == EB200_EVAL_GRAD ? &h->gh_grad : (mode == EB200_EVAL_VALUE_ONLY ? &h->gh_val : &h->gh_lml))
                                      : (mode == EB200_EVAL_GRAD ? &h->g_grad : (mode == EB200_EVAL_VALUE_ONLY ? &h->g_val : &h->g_lml));
  if (*slot == nullptr) {
    int* launches = mode == EB200_EVAL_GRAD ? &h->launches_grad : (mode == EB200_EVAL_VALUE_ONLY ? &h->launches_val : &h->launches_lml);
    if (capture_graph(h, mode, slot, launches, host_copies)) *slot = nullptr;
  }
  return *slot;
}
static cudaGraphExec_t eval_graph_host(eb200_gp* h, int mode) { return eval_graph(h, mode, true); }

int eb200_gp_eval_device(eb200_gp* h, const double* p_dev, int mode, double* result_dev, void* stream) {
  if (!h || !p_dev) return fail("eval: null argument");
  if (mode < 0 || mode > 2) return fail("eval: mode must be EB200_EVAL_LML, EB200_EVAL_GRAD or EB200_EVAL_VALUE_ONLY");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  CK(cudaMemcpyAsync(h->p_dev, p_dev, (h->d + 1) * sizeof(double), cudaMemcpyDeviceToDevice, st));
  cudaGraphExec_t g = eval_graph(h, mode);
  if (!g) return 1;
  CK(cudaGraphLaunch(g, st));
  if (result_dev)
    CK(cudaMemcpyAsync(result_dev, h->out_dev, EB200_RESULT_LEN(h->d) * sizeof(double), cudaMemcpyDeviceToDevice, st));
  h->factor_valid = mode != EB200_EVAL_VALUE_ONLY;  // a value-only evaluation leaves no L^-1 / alpha behind
  h->x_in_a = false;  // (matrix A holds L again)
  return 0;
}

int eb200_gp_eval_host(eb200_gp* h, const double* p_host, int mode, double* result_host) {
  if (!h || !p_host || !result_host) return fail("eval: null argument");
  if (mode < 0 || mode > 2) return fail("eval: mode must be EB200_EVAL_LML, EB200_EVAL_GRAD or EB200_EVAL_VALUE_ONLY");
  CK(cudaSetDevice(h->device));
  const int len = EB200_RESULT_LEN(h->d);
  memcpy(h->h_p, p_host, (h->d + 1) * sizeof(double));
  cudaGraphExec_t g = eval_graph_host(h, mode);
  if (!g) return 1;
  CK(cudaGraphLaunch(g, h->stream));  // parameters in, kernels, result out: one launch
  CK(cudaStreamSynchronize(h->stream));
  memcpy(result_host, h->h_out, len * sizeof(double));
  if (mode != EB200_EVAL_GRAD)
    for (int i = 0; i <= h->d; ++i) result_host[1 + i] = 0.0;
  // prediction state (L^-1, alpha) is only there after a successful factorisation that kept it
  h->factor_valid = mode != EB200_EVAL_VALUE_ONLY && result_host[h->d + 2] == 0.0;
  h->x_in_a = false;  // (matrix A holds L again)
  return 0;
}

static int ensure_group_state(eb200_gp* h) {
  eb200_gp::GroupState& gs = h->group;
  if (gs.p_dev) return 0;
  const size_t PS_ = MAXD + 1, OS_ = MAXD + 8;
  CK(cudaMalloc(&gs.p_dev, MAX_BATCH * PS_ * sizeof(double)));
  CK(cudaMalloc(&gs.out_dev, MAX_BATCH * OS_ * sizeof(double)));
  CK(cudaMallocHost(&gs.h_stage, MAX_BATCH * (PS_ + OS_) * sizeof(double)));
  return 0;
}

// One evaluation on each of m equally shaped handles (m <= MAX_BATCH) as ONE sequence of launches on stream
// `st`: the factorisations share a launch with their dependency chains interleaved (factor_dataflow.cuh: one
// Cholesky chain cannot feed 148 SMs until it is a few hundred columns long, m of them can), and every later
// stage is one launch over the problems of the group.  Problem b reads p_dev + b * p_stride and writes its
// result vector to out_dev + b * out_stride (device pointers).  Results equal stand-alone evaluations bit for
// bit: every tile and every reduction is computed by the same code in the same order whoever claims it.
static int enqueue_group(eb200_gp** g, int m, int mode, const double* p_dev, int p_stride, double* out_dev, int out_stride,
                         cudaStream_t st, cudaEvent_t* stage_events = nullptr, unsigned long long* const* traces = nullptr) {
  eb200_gp* h0 = g[0];
  const int nt = h0->nt;
  const bool vo = mode == EB200_EVAL_VALUE_ONLY;
  FactorBatch fb;
  fb.nb = m;
  PostBatch pb;
  pb.nb = m;
  PrepBatch prep;
  const int ntasks = vo ? h0->n_ftasks_v : h0->n_ftasks;
  const int grid = (int)std::min<long long>(2ll * h0->n_sms, (long long)m * ntasks);
  // chain CTAs per problem: a CTA of the chain queue claims in order and waits inside its task, so it is a
  // CTA slot that mostly idles -- as few as keep the chain fed
  // (measured at N = 4096: see groups_pay)
  static const int chain_env = env_int("EB200_GROUP_CHAIN", 0);
  int n_chain = std::max(1, std::min({6, nt / 4, (grid / m) / 4}));
  if (nt > 32) n_chain = std::min(n_chain, 2);
  if (chain_env > 0) n_chain = chain_env;
  for (int b = 0; b < m; ++b) {
    eb200_gp* h = g[b];
    CK(reset_state(h, st));
    prep.kp[b] = h->kp;
    FactorArgs& fa = fb.p[b];
    fill_factor_args(h, vo, fa);
    fa.sm_chain = fb.p[0].sm_chain;  // one priority flag per SM for the whole launch
    fa.trace = traces ? traces[b] : nullptr;
    fa.n_chain_sms = n_chain;
    fill_post_args(h, mode, pb.p[b]);
    pb.p[b].out = out_dev + (size_t)b * out_stride;
  }
  // (stage_events, diagnostics: [0] before the parameters, [s] after the last kernel of stage s = 1..5)
  if (stage_events) CK(cudaEventRecord(stage_events[0], st));
  prep_params_group_kernel<<<m, 32, 0, st>>>(p_dev, p_stride, h0->d, prep);
  if (stage_events) CK(cudaEventRecord(stage_events[1], st));
  DP_SWITCH(h0->dp, (factor_dataflow_kernel<DP, true><<<grid, 256, FactorSmem<DP>::BYTES, st>>>(fb)));
  if (stage_events) CK(cudaEventRecord(stage_events[2], st));
  for (const Op& op : h0->ops) {
    if (op.stage < 3 || op.debug_only) continue;
    if (op.grad_only && mode != EB200_EVAL_GRAD) continue;
    if (vo && op.stage == 3) continue;
    launch_post(h0, op, pb, st, mode, nullptr, g);
    if (stage_events) CK(cudaEventRecord(stage_events[op.stage], st));  // (re-recorded by a stage's later kernels: the last stands)
  }
  CK(cudaGetLastError());
  return 0;
}

static bool same_shape(const eb200_gp* a, const eb200_gp* b) {
  return a->np == b->np && a->dp == b->dp && a->d == b->d && a->n == b->n && a->device == b->device;
}

// Whether handles shaped like h share factorisation launches when several are evaluated together.
// Measured on B200, ms per LML+grad evaluation, stand-alone -> grouped (d = 9 unless noted):
//   N = 256 (d = 8) 0.012 vs 0.038 and N = 640 0.065 vs 0.077 with the handles' own graphs overlapping on their streams:
//   below ~14 tile rows that is the better way;  N = 1000 (d = 3) 0.405 -> 0.109;  2048: 0.88 -> 0.42;  3072: 1.51 -> 1.16;
//   4096: 2.92 -> 2.56 in groups of eight (2.58 in groups of four; 2 chain CTAs per problem: 1 -> 2.57 / 2.83, 3 -> 2.58);
//   6144: 8.62 -> 8.06 and 8176: 19.26 -> 18.50 in groups of four (18.60 in pairs).
static bool groups_pay(const eb200_gp* h, int mode) {
  static const int lo = env_int("EB200_GROUP_MIN_NT", 14), hi = env_int("EB200_GROUP_MAX_NT", 160);
  return mode != EB200_EVAL_VALUE_ONLY && h->nt >= lo && h->nt <= hi;
}
static int group_size(const eb200_gp* h) {
  static const int env = env_int("EB200_GROUP_SIZE", 0);
  const int m = env > 0 ? env : (h->nt <= 64 ? MAX_BATCH : 4);
  return std::max(1, std::min((int)MAX_BATCH, m));
}

static int check_distinct(eb200_gp** gps, int nb, const char* who) {
  std::vector<eb200_gp*> v(gps, gps + nb);
  std::sort(v.begin(), v.end());
  for (int b = 0; b < nb; ++b) {
    if (!v[b]) return fail(std::string(who) + ": null handle");
    if (b > 0 && v[b] == v[b - 1]) return fail(std::string(who) + ": the same handle appears twice (its matrices would be factorised twice at once)");
  }
  return 0;
}

// One evaluation on each of nb handles in ONE native call.  Equally shaped problems of at least 14 tile rows
// share launches, MAX_BATCH at a time (enqueue_group); everything else is enqueued on the handles' own streams.
// Either way the evaluations overlap on the device and are collected afterwards.
int eb200_gp_eval_multi_host(eb200_gp** gps, int nb, const double* p_host, int p_stride, int mode, double* result_host,
                             int result_stride) {
  if (!gps || !p_host || !result_host || nb < 0) return fail("eval_multi: bad argument");
  if (mode < 0 || mode > 2) return fail("eval: mode must be EB200_EVAL_LML, EB200_EVAL_GRAD or EB200_EVAL_VALUE_ONLY");
  if (check_distinct(gps, nb, "eval_multi")) return 1;
  for (int b = 0; b < nb; ++b)
    if (p_stride < gps[b]->d + 1 || result_stride < EB200_RESULT_LEN(gps[b]->d)) return fail("eval_multi: stride too small");
  const int PS_ = MAXD + 1, OS_ = MAXD + 8;
  std::vector<int> lead(nb, -1);  // for grouped problems: index of the group's first handle
  for (int b = 0; b < nb;) {
    eb200_gp* h = gps[b];
    CK(cudaSetDevice(h->device));
    int m = 1;
    if (groups_pay(h, mode))
      while (b + m < nb && m < group_size(h) && same_shape(gps[b + m], h)) ++m;
    if (m >= 2) {
      if (ensure_group_state(h)) return 1;
      double* stage_p = h->group.h_stage;
      for (int k = 0; k < m; ++k) {
        memcpy(stage_p + (size_t)k * PS_, p_host + (size_t)(b + k) * p_stride, (h->d + 1) * sizeof(double));
        lead[b + k] = b;
      }
      CK(cudaMemcpyAsync(h->group.p_dev, stage_p, (size_t)m * PS_ * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      if (enqueue_group(gps + b, m, mode, h->group.p_dev, PS_, h->group.out_dev, OS_, h->stream)) return 1;
      CK(cudaMemcpyAsync(stage_p + (size_t)MAX_BATCH * PS_, h->group.out_dev, (size_t)m * OS_ * sizeof(double), cudaMemcpyDeviceToHost,
                         h->stream));
    } else {
      memcpy(h->h_p, p_host + (size_t)b * p_stride, (h->d + 1) * sizeof(double));
      cudaGraphExec_t g = eval_graph_host(h, mode);
      if (!g) return 1;
      CK(cudaGraphLaunch(g, h->stream));
    }
    b += m;
  }
  for (int b = 0; b < nb; ++b) {
    eb200_gp* h = gps[b];
    const int len = EB200_RESULT_LEN(h->d);
    double* dst = result_host + (size_t)b * result_stride;
    CK(cudaSetDevice(h->device));
    if (lead[b] >= 0) {
      eb200_gp* h0 = gps[lead[b]];
      if (lead[b] == b) CK(cudaStreamSynchronize(h0->stream));
      memcpy(dst, h0->group.h_stage + (size_t)MAX_BATCH * PS_ + (size_t)(b - lead[b]) * OS_, len * sizeof(double));
    } else {
      CK(cudaStreamSynchronize(h->stream));
      memcpy(dst, h->h_out, len * sizeof(double));
    }
    if (mode != EB200_EVAL_GRAD)
      for (int i = 0; i <= h->d; ++i) dst[1 + i] = 0.0;
    // prediction state (L^-1, alpha) is only there after a successful factorisation that kept it
    h->factor_valid = mode != EB200_EVAL_VALUE_ONLY && dst[h->d + 2] == 0.0;
    h->x_in_a = false;  // (matrix A holds L again)
  }
  return 0;
}

// The same with device pointers: parameter vectors p_dev[nb][p_stride] and result vectors
// result_dev[nb][result_stride] live in HBM, everything is enqueued on `stream` (NULL: the first handle's own)
// and nothing synchronises.  All handles must be on the current device of `stream`.
int eb200_gp_eval_multi_device(eb200_gp** gps, int nb, const double* p_dev, int p_stride, int mode, double* result_dev,
                               int result_stride, void* stream) {
  if (!gps || !p_dev || !result_dev || nb < 0) return fail("eval_multi_device: bad argument");
  if (mode < 0 || mode > 2) return fail("eval: mode must be EB200_EVAL_LML, EB200_EVAL_GRAD or EB200_EVAL_VALUE_ONLY");
  if (nb == 0) return 0;
  if (check_distinct(gps, nb, "eval_multi_device")) return 1;
  for (int b = 0; b < nb; ++b) {
    if (p_stride < gps[b]->d + 1 || result_stride < EB200_RESULT_LEN(gps[b]->d)) return fail("eval_multi_device: stride too small");
    if (gps[b]->device != gps[0]->device) return fail("eval_multi_device: handles on different devices");
  }
  CK(cudaSetDevice(gps[0]->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : gps[0]->stream;
  for (int b = 0; b < nb;) {
    eb200_gp* h = gps[b];
    int m = 1;
    if (groups_pay(h, mode))
      while (b + m < nb && m < group_size(h) && same_shape(gps[b + m], h)) ++m;
    if (m >= 2) {
      if (enqueue_group(gps + b, m, mode, p_dev + (size_t)b * p_stride, p_stride, result_dev + (size_t)b * result_stride, result_stride, st))
        return 1;
      for (int k = 0; k < m; ++k) gps[b + k]->factor_valid = mode != EB200_EVAL_VALUE_ONLY;
      for (int k = 0; k < m; ++k) gps[b + k]->x_in_a = false;  // (matrix A holds L again)
    } else {
      if (eb200_gp_eval_device(h, p_dev + (size_t)b * p_stride, mode, result_dev + (size_t)b * result_stride, st)) return 1;
    }
    b += m;
  }
  return 0;
}

// nb value-only evaluations whose factorisations share launches, MAX_BATCH at a time.
int eb200_gp_eval_value_batch_host(eb200_gp* h, int nb, const double* p_host, double* result_host) {
  if (!h || !p_host || !result_host || nb < 1) return fail("eval_value_batch: bad argument");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int np = h->np, nt = h->nt, P = h->d + 1, len = EB200_RESULT_LEN(h->d);
  const size_t nf = (size_t)(nt + 1) * nt;
  const size_t flag_ints = 4 * nf + 8 * nt + 512 + (size_t)h->n_ftasks;
  const size_t zero_ints = flag_ints + 8 + 2;  // per problem: flags, 8 counters, info + abort flag
  const int PS_ = MAXD + 1, OS_ = MAXD + 8;    // strides of the parameter / result arrays
  eb200_gp::BatchState& bs = h->batch;
  if (!bs.zeroed) {
    CK(cudaMalloc(&bs.zeroed, (size_t)MAX_BATCH * zero_ints * sizeof(int)));
    CK(cudaMalloc(&bs.Xd, (size_t)MAX_BATCH * nt * TB * TB * sizeof(double)));
    CK(cudaMalloc(&bs.logdet, (size_t)MAX_BATCH * nt * sizeof(double)));
    CK(cudaMalloc(&bs.rdiag, (size_t)MAX_BATCH * np * sizeof(double)));
    CK(cudaMalloc(&bs.out, (size_t)MAX_BATCH * OS_ * sizeof(double)));
    CK(cudaMalloc(&bs.p_dev, (size_t)MAX_BATCH * PS_ * sizeof(double)));
    CK(cudaMalloc(&bs.kp, (size_t)MAX_BATCH * sizeof(KernelParams)));
    CK(cudaMallocHost(&bs.h_stage, (size_t)MAX_BATCH * (PS_ + OS_) * sizeof(double)));
    bs.A[0] = h->A;
    bs.slots = 1;
  }
  double* stage_p = bs.h_stage;
  double* stage_out = bs.h_stage + (size_t)MAX_BATCH * PS_;
  for (int g0 = 0; g0 < nb; g0 += MAX_BATCH) {
    const int m = std::min(MAX_BATCH, nb - g0);
    for (; bs.slots < m; ++bs.slots) {
      CK(cudaMalloc(&bs.A[bs.slots], ((size_t)np + TB) * np * sizeof(double)));
      CK(cudaMemset(bs.A[bs.slots], 0, ((size_t)np + TB) * np * sizeof(double)));
    }
    FactorBatch fb;
    fb.nb = m;
    ValueBatchPtrs zs;
    const int grid = std::min(2 * h->n_sms, m * h->n_ftasks_v);
    for (int b = 0; b < m; ++b) {
      memcpy(stage_p + (size_t)b * PS_, p_host + (size_t)(g0 + b) * P, P * sizeof(double));
      int* zb = bs.zeroed + (size_t)b * zero_ints;
      FactorArgs& fa = fb.p[b];
      fill_factor_args(h, true, fa);
      fa.A = bs.A[b]; fa.kp = bs.kp + b;
      fa.ready = zb; fa.xready = fa.ready + nf; fa.pready = fa.xready + nf; fa.cready = fa.pready + nf;
      fa.lready = fa.cready + nf;
      fa.sm_chain = bs.zeroed + 4 * nf + 8 * nt;  // one priority flag per SM, shared by the batch (problem 0's)
      fa.claimed = fa.lready + 8 * nt + 512;
      fa.counter = reinterpret_cast<unsigned*>(zb + flag_ints);
      fa.info = zb + flag_ints + 8; fa.abort_flag = fa.info + 1;
      fa.Xd = bs.Xd + (size_t)b * nt * TB * TB; fa.xd_tile = (long long)TB * TB; fa.xd_ld = TB;
      fa.logdet_part = bs.logdet + (size_t)b * nt; fa.rdiag = bs.rdiag + (size_t)b * np;
      fa.trace = nullptr;
      fa.n_chain_sms = std::max(1, std::min({8, nt / 4, (grid / m) / 4}));
      zs.z[b] = bs.A[b] + (size_t)np * np;
    }
    CK(cudaMemcpyAsync(bs.p_dev, stage_p, (size_t)m * PS_ * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(bs.zeroed, 0, (size_t)m * zero_ints * sizeof(int), st));
    prep_params_kernel<<<m, 32, 0, st>>>(bs.p_dev, h->d, bs.kp, PS_);
    DP_SWITCH(h->dp, (factor_dataflow_kernel<DP, true><<<grid, 256, FactorSmem<DP>::BYTES, st>>>(fb)));
    // info lives at a fixed offset inside each problem's zeroed block: hand the reduction a strided view
    finalize_value_batch_kernel<<<m, 256, 0, st>>>(bs.logdet, nt, zs, h->n, h->d,
                                                   bs.zeroed + flag_ints + 8, bs.out, OS_, (int)zero_ints);
    CK(cudaMemcpyAsync(stage_out, bs.out, (size_t)m * OS_ * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    for (int b = 0; b < m; ++b) memcpy(result_host + (size_t)(g0 + b) * len, stage_out + (size_t)b * OS_, len * sizeof(double));
  }
  h->factor_valid = false;  // value-only evaluations leave no L^-1 / alpha behind
  h->x_in_a = false;  // (matrix A holds L again)
  return 0;
}

int eb200_gp_predict_device(eb200_gp* h, int m, const double* t_dev, int want_var, double* mu_dev, double* var_dev,
                            void* stream) {
  if (!h || !t_dev || !mu_dev || (want_var && !var_dev)) return fail("predict: null argument");
  if (!h->factor_valid) return fail("predict: no factorisation resident (evaluate with EB200_EVAL_LML or EB200_EVAL_GRAD first)");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  const int chunk = predict_chunk_rows(h, (size_t)std::max(m, 0));
  if (ensure_predict_scratch(h, chunk, want_var)) return 1;
  double* rows = h->big.T;
  for (int q0 = 0; q0 < m; q0 += chunk) {
    const int mc = std::min(chunk, m - q0);
    const size_t tot = (size_t)mc * h->dp;
    pad_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(t_dev + (size_t)q0 * h->d, mc, h->d, h->dp, rows);
    if (predict_chunk(h, chunk, mc, want_var, mu_dev + q0, want_var ? var_dev + q0 : nullptr, st)) return 1;
  }
  return 0;
}

int eb200_gp_predict_host(eb200_gp* h, int m, const double* t_host, int want_var, double* mu_host, double* var_host) {
  if (!h || !t_host || !mu_host || (want_var && !var_host)) return fail("predict: null argument");
  if (!h->factor_valid) return fail("predict: no factorisation resident (evaluate with EB200_EVAL_LML or EB200_EVAL_GRAD first)");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int d = h->d;
  const int chunk = std::min(h->np, predict_chunk_rows(h, (size_t)std::max(m, 0)));  // (the pinned staging block holds np rows)
  if (ensure_predict_scratch(h, chunk, want_var)) return 1;
  for (int q0 = 0; q0 < m; q0 += chunk) {
    const int mc = std::min(chunk, m - q0);
    double* stage_t = h->h_pred;
    double* stage_mu = h->h_pred + (size_t)h->np * d;
    double* stage_var = stage_mu + h->np;
    memcpy(stage_t, t_host + (size_t)q0 * d, (size_t)mc * d * sizeof(double));
    CK(cudaMemcpyAsync(h->T_in, stage_t, (size_t)mc * d * sizeof(double), cudaMemcpyHostToDevice, st));
    const size_t tot = (size_t)mc * h->dp;
    pad_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(h->T_in, mc, d, h->dp, h->big.T);
    if (predict_chunk(h, chunk, mc, want_var, h->mu_dev, h->var_dev, st)) return 1;
    CK(cudaMemcpyAsync(stage_mu, h->mu_dev, mc * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (want_var) CK(cudaMemcpyAsync(stage_var, h->var_dev, mc * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(mu_host + q0, stage_mu, mc * sizeof(double));
    if (want_var) memcpy(var_host + q0, stage_var, mc * sizeof(double));
  }
  return 0;
}

int eb200_gp_predict_grid_device(eb200_gp* h, long long n_theta, const double* theta_dev, int n_ind, const double* ind_dev,
                                 int round_f32, int want_var, double* mu_dev, double* var_dev, void* stream) {
  if (!h || !theta_dev || !ind_dev || !mu_dev || (want_var && !var_dev)) return fail("predict_grid: null argument");
  if (n_theta < 0 || n_ind < 1 || h->d < 2) return fail("predict_grid: needs n_ind >= 1 and a kernel dimension >= 2");
  if (!h->factor_valid) return fail("predict: no factorisation resident (evaluate with EB200_EVAL_LML or EB200_EVAL_GRAD first)");
  if (n_theta == 0) return 0;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  return predict_grid_impl(h, (size_t)n_theta, theta_dev, n_ind, ind_dev, round_f32, want_var, mu_dev, var_dev, st);
}

int eb200_gp_predict_grid_host(eb200_gp* h, int n_theta, const double* theta_host, int n_ind, const double* ind_host,
                               int round_f32, int want_var, double* mu_host, double* var_host) {
  if (!h || !theta_host || !ind_host || !mu_host || (want_var && !var_host)) return fail("predict_grid: null argument");
  if (n_theta < 0 || n_ind < 1 || h->d < 2) return fail("predict_grid: needs n_ind >= 1 and a kernel dimension >= 2");
  if (!h->factor_valid) return fail("predict: no factorisation resident (evaluate with EB200_EVAL_LML or EB200_EVAL_GRAD first)");
  if (n_theta == 0) return 0;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const size_t m = (size_t)n_theta * n_ind;
  if (grow(&h->grid_theta, &h->grid_theta_cap, (size_t)n_theta * (h->d - 1))) return 1;
  if (grow(&h->grid_ind, &h->grid_ind_cap, (size_t)n_ind)) return 1;
  if (m > h->grid_out_cap) {
    if (h->grid_mu) CK(cudaFree(h->grid_mu));
    if (h->grid_var) CK(cudaFree(h->grid_var));
    h->grid_mu = h->grid_var = nullptr;
    h->grid_out_cap = 0;
    CK(cudaMalloc(&h->grid_mu, m * sizeof(double)));
    CK(cudaMalloc(&h->grid_var, m * sizeof(double)));
    h->grid_out_cap = m;
  }
  CK(cudaMemcpyAsync(h->grid_theta, theta_host, (size_t)n_theta * (h->d - 1) * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(h->grid_ind, ind_host, (size_t)n_ind * sizeof(double), cudaMemcpyHostToDevice, st));
  if (predict_grid_impl(h, (size_t)n_theta, h->grid_theta, n_ind, h->grid_ind, round_f32, want_var, h->grid_mu, h->grid_var, st)) return 1;
  CK(cudaMemcpyAsync(mu_host, h->grid_mu, m * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (want_var) CK(cudaMemcpyAsync(var_host, h->grid_var, m * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int eb200_gp_synchronize(eb200_gp* h) {
  if (!h) return fail("synchronize: null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int eb200_gp_debug_stage(eb200_gp* h, const double* p_host, int stage) {
  if (!h || !p_host) return fail("debug_stage: null argument");
  CK(cudaSetDevice(h->device));
  memcpy(h->h_p, p_host, (h->d + 1) * sizeof(double));
  CK(cudaMemcpyAsync(h->p_dev, h->h_p, (h->d + 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (enqueue_eval(h, h->stream, stage >= 4, stage, stage >= 4 ? h->A : nullptr, nullptr)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  h->factor_valid = false;
  h->x_in_a = false;  // (matrix A holds L again)
  return 0;
}

int eb200_gp_debug_matrix(eb200_gp* h, int which, double* out_host) {
  if (!h || !out_host) return fail("debug_matrix: null argument");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(out_host, which ? h->W : h->A, (size_t)h->np * h->np * sizeof(double), cudaMemcpyDeviceToHost));
  if (which) {  // the device keeps Y = X^T; the caller asked for X = L^-1
    const size_t np = (size_t)h->np;
    for (size_t i = 0; i < np; ++i)
      for (size_t j = i + 1; j < np; ++j) std::swap(out_host[i * np + j], out_host[j * np + i]);
  }
  return 0;
}

int eb200_gp_debug_alpha(eb200_gp* h, double* alpha_host) {
  if (!h || !alpha_host) return fail("debug_alpha: null argument");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(alpha_host, h->alpha, h->n * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

static int factor_trace_impl(eb200_gp* h, const double* p_host, int mode, unsigned long long* out_host, int* tasks_out) {
  if (!h || !p_host || !out_host || !tasks_out) return fail("factor_trace: null argument");
  CK(cudaSetDevice(h->device));
  const bool vo = mode == EB200_EVAL_VALUE_ONLY;
  const int ntasks = vo ? h->n_ftasks_v : h->n_ftasks;
  CK(cudaMalloc(&h->trace, (size_t)ntasks * TRACE_SLOTS * sizeof(unsigned long long)));
  CK(cudaMemset(h->trace, 0, (size_t)ntasks * TRACE_SLOTS * sizeof(unsigned long long)));
  memcpy(h->h_p, p_host, (h->d + 1) * sizeof(double));
  CK(cudaMemcpyAsync(h->p_dev, h->h_p, (h->d + 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  int rc = enqueue_eval(h, h->stream, mode, 2, nullptr, nullptr);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (!rc && e == cudaSuccess) {
    e = cudaMemcpy(out_host, h->trace, (size_t)ntasks * TRACE_SLOTS * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(tasks_out, vo ? h->ftasks_v : h->ftasks, (size_t)ntasks * sizeof(FactorTask), cudaMemcpyDeviceToHost);
  }
  cudaFree(h->trace);
  h->trace = nullptr;
  h->factor_valid = false;
  h->x_in_a = false;  // (matrix A holds L again)
  if (rc) return rc;
  CK(e);
  return 0;
}
int eb200_gp_debug_factor_trace(eb200_gp* h, const double* p_host, unsigned long long* out_host, int* tasks_out) {
  return factor_trace_impl(h, p_host, EB200_EVAL_LML, out_host, tasks_out);
}
int eb200_gp_debug_value_trace(eb200_gp* h, const double* p_host, unsigned long long* out_host, int* tasks_out) {
  return factor_trace_impl(h, p_host, EB200_EVAL_VALUE_ONLY, out_host, tasks_out);
}
// The same stamps for a GROUP of equally shaped handles sharing one factorisation launch: out_host[nb][tasks][10].
int eb200_gp_debug_group_trace(eb200_gp** gps, int nb, const double* p_host, int p_stride, unsigned long long* out_host, int* tasks_out) {
  if (!gps || !p_host || !out_host || !tasks_out || nb < 1 || nb > MAX_BATCH) return fail("group_trace: bad argument");
  if (check_distinct(gps, nb, "group_trace")) return 1;
  eb200_gp* h = gps[0];
  for (int b = 0; b < nb; ++b)
    if (!same_shape(gps[b], h)) return fail("group_trace: the handles must share one shape");
  CK(cudaSetDevice(h->device));
  if (ensure_group_state(h)) return 1;
  const int PS_ = MAXD + 1, OS_ = MAXD + 8;
  const size_t per = (size_t)h->n_ftasks * TRACE_SLOTS;
  unsigned long long* dev = nullptr;
  CK(cudaMalloc(&dev, nb * per * sizeof(unsigned long long)));
  CK(cudaMemset(dev, 0, nb * per * sizeof(unsigned long long)));
  unsigned long long* ptrs[MAX_BATCH];
  for (int b = 0; b < nb; ++b) {
    ptrs[b] = dev + b * per;
    memcpy(h->group.h_stage + (size_t)b * PS_, p_host + (size_t)b * p_stride, (h->d + 1) * sizeof(double));
  }
  CK(cudaMemcpyAsync(h->group.p_dev, h->group.h_stage, (size_t)nb * PS_ * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  int rc = 0;
  for (int rep = 0; rep < 2 && !rc; ++rep)  // (the second run is the one kept: warm caches)
    rc = enqueue_group(gps, nb, EB200_EVAL_LML, h->group.p_dev, PS_, h->group.out_dev, OS_, h->stream, nullptr, ptrs);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (!rc && e == cudaSuccess) {
    e = cudaMemcpy(out_host, dev, nb * per * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(tasks_out, h->ftasks, (size_t)h->n_ftasks * sizeof(FactorTask), cudaMemcpyDeviceToHost);
  }
  cudaFree(dev);
  for (int b = 0; b < nb; ++b) { gps[b]->factor_valid = false; gps[b]->x_in_a = false; }
  if (rc) return rc;
  CK(e);
  return 0;
}
int eb200_gp_factor_tasks(const eb200_gp* h) { return h ? h->n_ftasks : -1; }
int eb200_gp_value_tasks(const eb200_gp* h) { return h ? h->n_ftasks_v : -1; }

int eb200_gp_profile_stages(eb200_gp* h, const double* p_host, int reps, float* ms_out) {
  if (!h || !p_host || !ms_out || reps < 1) return fail("profile_stages: bad argument");
  CK(cudaSetDevice(h->device));
  memcpy(h->h_p, p_host, (h->d + 1) * sizeof(double));
  CK(cudaMemcpyAsync(h->p_dev, h->h_p, (h->d + 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  cudaEvent_t ev[6];
  for (auto& e : ev) CK(cudaEventCreate(&e));
  for (int s = 0; s < 5; ++s) ms_out[s] = 0.f;
  for (int rep = 0; rep <= reps; ++rep) {  // rep 0 is a warm-up
    CK(cudaEventRecord(ev[0], h->stream));
    for (int stage = 1; stage <= 5; ++stage) {
      for (const Op& op : h->ops)
        if (op.stage == stage && !op.debug_only) launch_op(h, op, h->stream, 1, nullptr);
      CK(cudaEventRecord(ev[stage], h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    if (rep == 0) continue;
    for (int s = 0; s < 5; ++s) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, ev[s], ev[s + 1]));
      ms_out[s] += ms / reps;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  h->factor_valid = true;
  h->x_in_a = false;  // (matrix A holds L again)
  return 0;
}

int eb200_gp_profile_group_stages(eb200_gp** gps, int nb, const double* p_host, int p_stride, int reps, float* ms_out) {
  if (!gps || !p_host || !ms_out || reps < 1 || nb < 1 || nb > MAX_BATCH) return fail("profile_group_stages: bad argument");
  if (check_distinct(gps, nb, "profile_group_stages")) return 1;
  eb200_gp* h = gps[0];
  for (int b = 0; b < nb; ++b)
    if (!same_shape(gps[b], h) || p_stride < h->d + 1) return fail("profile_group_stages: the handles must share one shape");
  CK(cudaSetDevice(h->device));
  if (ensure_group_state(h)) return 1;
  const int PS_ = MAXD + 1, OS_ = MAXD + 8;
  for (int b = 0; b < nb; ++b) memcpy(h->group.h_stage + (size_t)b * PS_, p_host + (size_t)b * p_stride, (h->d + 1) * sizeof(double));
  CK(cudaMemcpyAsync(h->group.p_dev, h->group.h_stage, (size_t)nb * PS_ * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  cudaEvent_t ev[6];
  for (auto& e : ev) CK(cudaEventCreate(&e));
  for (int s = 0; s < 5; ++s) ms_out[s] = 0.f;
  for (int rep = 0; rep <= reps; ++rep) {  // rep 0 is a warm-up
    if (enqueue_group(gps, nb, EB200_EVAL_GRAD, h->group.p_dev, PS_, h->group.out_dev, OS_, h->stream, ev)) return 1;
    CK(cudaStreamSynchronize(h->stream));
    if (rep == 0) continue;
    for (int s = 0; s < 5; ++s) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, ev[s], ev[s + 1]));
      ms_out[s] += ms / reps;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  for (int b = 0; b < nb; ++b) gps[b]->factor_valid = true;
  for (int b = 0; b < nb; ++b) gps[b]->x_in_a = false;  // (matrix A holds L again)
  return 0;
}

int eb200_bench_accumulate(int m, int n, int k, int b_mcontig, int grid, const double* a_dev, const double* b_dev,
                           double* c_dev, void* stream) {
  if (m % 64 || n % 64 || k % 64) return fail("bench_accumulate: m, n, k must be multiples of 64");
  static std::mutex mu;  // (the cached device state below is shared by all callers)
  std::lock_guard<std::mutex> lock(mu);
  static int* ones = nullptr;
  if (!ones) {
    CK(cudaMalloc(&ones, 2 * sizeof(int)));
    int h[2] = {1, 0};
    CK(cudaMemcpy(ones, h, sizeof(h), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(accumulate_bench_kernel<KCONTIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, FactorSmem<2>::BYTES));
    CK(cudaFuncSetAttribute(accumulate_bench_kernel<MCONTIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, FactorSmem<2>::BYTES));
  }
  const int ntiles = (m / 64) * (n / 64);
  if (b_mcontig)
    accumulate_bench_kernel<MCONTIG><<<grid, 256, FactorSmem<2>::BYTES, (cudaStream_t)stream>>>(a_dev, k, b_dev, n, c_dev, n, k / KC, n / 64,
                                                                                                ntiles, ones, ones + 1);
  else
    accumulate_bench_kernel<KCONTIG><<<grid, 256, FactorSmem<2>::BYTES, (cudaStream_t)stream>>>(a_dev, k, b_dev, k, c_dev, n, k / KC, n / 64,
                                                                                                ntiles, ones, ones + 1);
  CK(cudaGetLastError());
  return 0;
}

int eb200_bench_gemm_nt(int m, int n, int k, const double* a_dev, const double* b_dev, double* c_dev, void* stream) {
  if (m % 128 || n % 128 || k % (2 * KC)) return fail("bench_gemm_nt: m, n must be multiples of 128 and k of 32");
  if (set_attrs()) return 1;
  static std::mutex mu;  // (the cached task list below is shared by all callers)
  std::lock_guard<std::mutex> lock(mu);
  static TileTask* d_tasks = nullptr;
  static int cached_m = -1, cached_n = -1, cached_k = -1;
  if (cached_m != m || cached_n != n || cached_k != k) {
    std::vector<TileTask> v;
    for (int I = 0; I < m / 128; ++I)
      for (int J = 0; J < n / 128; ++J) v.push_back({I * 128, J * 128, I * 128, 0, J * 128, 0, k / KC, 0});
    if (d_tasks) cudaFree(d_tasks);
    CK(cudaMalloc(&d_tasks, v.size() * sizeof(TileTask)));
    CK(cudaMemcpy(d_tasks, v.data(), v.size() * sizeof(TileTask), cudaMemcpyHostToDevice));
    cached_m = m;
    cached_n = n;
    cached_k = k;
  }
  static const int use_tma = env_int("EB200_GEMM_TMA", 1);
  if (use_tma) {
    CUtensorMap map_a, map_b;
    if (make_tensor_map(&map_a, a_dev, m, k, k) || make_tensor_map(&map_b, b_dev, n, k, k)) return 1;
    tile_gemm_tma_kernel<<<(m / 128) * (n / 128), TMA_THREADS, TMA_SMEM_BYTES, (cudaStream_t)stream>>>(d_tasks, map_a, map_b, c_dev, n);
  } else {
    tile_gemm_kernel<Cfg128, KCONTIG, KCONTIG>
        <<<(m / 128) * (n / 128), Cfg128::THREADS, Engine<Cfg128, KCONTIG, KCONTIG>::SMEM_BYTES, (cudaStream_t)stream>>>(
            d_tasks, c_dev, n, a_dev, k, b_dev, k, 1.0, 0.0);
  }
  CK(cudaGetLastError());
  return 0;
}

}  // extern "C"
