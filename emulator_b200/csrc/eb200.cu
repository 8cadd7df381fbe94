// C ABI implementation: GP handle, launch program (tile-task lists), CUDA graphs.
// See include/emulator_b200.h for the contract and DESIGN.md for the algorithm.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/emulator_b200.h"
#include "kernels.cuh"
#include "potrf_dataflow.cuh"

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
  OP_POTRF_DF,     // persistent dataflow Cholesky (K generated in-kernel)
  OP_DIAG_INV,     // batched 64x64 triangular inverses of the diagonal tiles
  OP_GEMM64_KK,    // Cfg64, A KCONTIG, B KCONTIG
  OP_GEMM128_KK,   // Cfg128, K/K
  OP_GEMM64_KM,    // Cfg64, A KCONTIG, B MCONTIG
  OP_GEMM128_KM,   // Cfg128, K/M
  OP_PLACE_DINV,
  OP_TRMV,
  OP_TRMV_T,
  OP_LAUUM,
  OP_FINALIZE
};

enum Buf { BUF_A = 0, BUF_W = 1, BUF_DINV = 2 };

struct Op {
  OpKind kind;
  int task_off = 0, task_cnt = 0;
  int cbuf = BUF_A, abuf = BUF_A, bbuf = BUF_A;
  int dinv_tile = 0;  // for BUF_DINV operands / OP_DIAG
  double alpha = 1.0, beta = 0.0;
  int stage = 0;      // debug stage this op belongs to (1 kbuild, 2 potrf, 3 trtri, 4 lauum/trace)
  bool grad_only = false;
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
  double *dinv = nullptr, *logdet_part = nullptr, *tpartial = nullptr, *trace_partial = nullptr;
  double *p_dev = nullptr, *out_dev = nullptr;
  KernelParams* kp = nullptr;
  int* info = nullptr;
  double *h_p = nullptr, *h_out = nullptr;  // pinned
  TileTask* tasks = nullptr;
  int2* ktiles = nullptr;
  int2* ptiles = nullptr;     // potrf task list
  int n_ptiles = 0;
  unsigned* counter = nullptr;
  int* ready = nullptr;
  int* abort_flag = nullptr;
  double* rdiag = nullptr;
  unsigned long long* trace = nullptr;  // diagnostics only
  int n_sms = 148;
  int n_ktiles = 0, n_lauum = 0;
  std::vector<Op> ops;
  cudaGraphExec_t g_grad = nullptr, g_lml = nullptr;
  int launches_grad = 0, launches_lml = 0;
  bool factor_valid = false;
  // prediction scratch
  double *T_dev = nullptr, *mu_dev = nullptr, *var_dev = nullptr, *rowss = nullptr, *T_in = nullptr;
  double *h_pred = nullptr;  // pinned staging for predict (np * (d + 2) doubles)
  TileTask* var_tasks = nullptr;
  int var_tasks_mt = -1, var_tasks_cnt = 0;
};

namespace {

double* buf_ptr(eb200_gp* h, int b, int dinv_tile) {
  if (b == BUF_A) return h->A;
  if (b == BUF_W) return h->W;
  return h->dinv + (size_t)dinv_tile * TB * TB;
}
int buf_ld(eb200_gp* h, int b) { return b == BUF_DINV ? TB : h->np; }

void sort_tasks(std::vector<TileTask>& v, size_t off) {
  std::stable_sort(v.begin() + off, v.end(), [](const TileTask& a, const TileTask& b) { return a.nk > b.nk; });
}

// Build the static launch program for one (np): Cholesky (right-looking, 128-wide outer panels
// made of two 64-wide sub-panels), recursive blocked triangular inverse, X^T X + traces.
void build_program(eb200_gp* h, std::vector<TileTask>& tasks) {
  const int nt = h->nt, nT = h->nT, np = h->np;
  auto& ops = h->ops;
  auto push_gemm = [&](OpKind kind, size_t off, int cb, int ab, int bb, int dt, double alpha, double beta,
                       int stage) {
    if (tasks.size() == off) return;
    sort_tasks(tasks, off);
    Op op;
    op.kind = kind;
    op.task_off = (int)off;
    op.task_cnt = (int)(tasks.size() - off);
    op.cbuf = cb;
    op.abuf = ab;
    op.bbuf = bb;
    op.dinv_tile = dt;
    op.alpha = alpha;
    op.beta = beta;
    op.stage = stage;
    ops.push_back(op);
  };
  {
    Op op;
    op.kind = OP_PREP;
    op.stage = 1;
    ops.push_back(op);
    op.kind = OP_KBUILD;
    ops.push_back(op);
  }
  // ---- Cholesky: one persistent dataflow kernel + batched diagonal inverses ----
  {
    Op op;
    op.stage = 2;
    op.kind = OP_POTRF_DF;
    ops.push_back(op);
    op.kind = OP_DIAG_INV;
    ops.push_back(op);
  }
  // ---- triangular inverse: X = L^-1 ----
  {
    Op op;
    op.kind = OP_PLACE_DINV;
    op.stage = 3;
    ops.push_back(op);
  }
  {
    // level 0: inside every 128 block, X21 = -X22 (L21 X11) with 64x64 tiles
    size_t off = tasks.size();
    for (int b = 0; b < nT; ++b) {
      int r0 = b * 128;
      tasks.push_back({r0 + 64, r0, r0 + 64, r0, r0, r0, TB / KC, 0});
    }
    push_gemm(OP_GEMM64_KM, off, BUF_W, BUF_A, BUF_A, 0, 1.0, 0.0, 3);
    off = tasks.size();
    for (int b = 0; b < nT; ++b) {
      int r0 = b * 128;
      tasks.push_back({r0 + 64, r0, r0 + 64, r0 + 64, r0 + 64, r0, TB / KC, 0});
    }
    push_gemm(OP_GEMM64_KM, off, BUF_A, BUF_A, BUF_W, 0, -1.0, 0.0, 3);
  }
  {
    struct Job { int r0, h1, h2; };
    std::vector<std::vector<Job>> levels;
    // returns level index of the node covering [b0, b0+nb) 128-blocks (0 = leaf 128 block)
    struct Rec {
      std::vector<std::vector<Job>>& levels;
      int go(int b0, int nb) {
        if (nb == 1) return 0;
        int n1 = (nb + 1) / 2, n2 = nb - n1;
        int l1 = go(b0, n1), l2 = go(b0 + n1, n2);
        int lv = std::max(l1, l2) + 1;
        if ((int)levels.size() < lv) levels.resize(lv);
        levels[lv - 1].push_back({b0 * 128, n1 * 128, n2 * 128});
        return lv;
      }
    } rec{levels};
    rec.go(0, nT);
    for (auto& jobs : levels) {
      size_t off = tasks.size();
      for (auto& j : jobs)
        for (int I = 0; I < j.h2 / 128; ++I)
          for (int J = 0; J < j.h1 / 128; ++J) {
            int kb = J * 128;
            tasks.push_back({j.r0 + j.h1 + I * 128, j.r0 + J * 128, j.r0 + j.h1 + I * 128, j.r0 + kb, j.r0 + kb,
                             j.r0 + J * 128, (j.h1 - kb) / KC, 0});
          }
      push_gemm(OP_GEMM128_KM, off, BUF_W, BUF_A, BUF_A, 0, 1.0, 0.0, 3);
      off = tasks.size();
      for (auto& j : jobs)
        for (int I = 0; I < j.h2 / 128; ++I)
          for (int J = 0; J < j.h1 / 128; ++J)
            tasks.push_back({j.r0 + j.h1 + I * 128, j.r0 + J * 128, j.r0 + j.h1 + I * 128, j.r0 + j.h1, j.r0 + j.h1,
                             j.r0 + J * 128, (I + 1) * 128 / KC, 0});
      push_gemm(OP_GEMM128_KM, off, BUF_A, BUF_A, BUF_W, 0, -1.0, 0.0, 3);
    }
  }
  // ---- z = X r, alpha = X^T z ----
  {
    Op op;
    op.stage = 3;
    op.kind = OP_TRMV;
    ops.push_back(op);
    op.kind = OP_TRMV_T;
    ops.push_back(op);
  }
  // ---- K^-1 tiles + gradient traces ----
  {
    size_t off = tasks.size();
    for (int J = 0; J < nT; ++J)
      for (int I = J; I < nT; ++I)
        tasks.push_back({I * 128, J * 128, I * 128, I * 128, I * 128, J * 128, (np - I * 128) / KC, I == J ? 1 : 0});
    sort_tasks(tasks, off);
    Op op;
    op.kind = OP_LAUUM;
    op.task_off = (int)off;
    op.task_cnt = (int)(tasks.size() - off);
    op.stage = 4;
    op.grad_only = true;
    ops.push_back(op);
    h->n_lauum = op.task_cnt;
  }
  {
    Op op;
    op.kind = OP_FINALIZE;
    op.stage = 5;
    ops.push_back(op);
  }
}

// Enqueue one op.  Returns number of kernel launches issued.
int launch_op(eb200_gp* h, const Op& op, cudaStream_t st, int want_grad, double* kinv_out) {
  const int np = h->np;
  switch (op.kind) {
    case OP_PREP:
      prep_params_kernel<<<1, 32, 0, st>>>(h->p_dev, h->d, h->kp);
      return 1;
    case OP_KBUILD:
      DP_SWITCH(h->dp, (kbuild_kernel<DP><<<h->n_ktiles, 256, 0, st>>>(h->ktiles, h->X, h->noise, h->kp, h->n, np, h->A)));
      return 1;
    case OP_POTRF_DF: {
      cudaMemsetAsync(h->counter, 0, sizeof(unsigned), st);
      cudaMemsetAsync(h->ready, 0, (size_t)h->nt * h->nt * sizeof(int), st);
      PotrfArgs pa;
      pa.A = h->A; pa.ld = np; pa.nt = h->nt; pa.n = h->n; pa.X = h->X; pa.noise = h->noise; pa.kp = h->kp;
      pa.tiles = h->ptiles; pa.ntasks = h->n_ptiles; pa.counter = h->counter; pa.ready = h->ready;
      pa.logdet_part = h->logdet_part; pa.info = h->info; pa.abort_flag = h->abort_flag; pa.trace = h->trace; pa.rdiag = h->rdiag;
      const int grid = std::min(h->n_sms, h->n_ptiles);
      DP_SWITCH(h->dp, (potrf_dataflow_kernel<DP><<<grid, 256, PotrfSmem<DP>::BYTES, st>>>(pa)));
      return 1;
    }
    case OP_DIAG_INV:
      diag_inverse_kernel<<<h->nt, 256, DIAG_INV_SMEM_BYTES, st>>>(h->A, np, h->dinv);
      return 1;
    case OP_PLACE_DINV:
      place_dinv_kernel<<<h->nt, 256, 0, st>>>(h->A, np, h->dinv);
      return 1;
    case OP_GEMM64_KK:
      tile_gemm_kernel<Cfg64, KCONTIG, KCONTIG><<<op.task_cnt, Cfg64::THREADS, Engine<Cfg64, KCONTIG, KCONTIG>::SMEM_BYTES, st>>>(
          h->tasks + op.task_off, buf_ptr(h, op.cbuf, 0), buf_ld(h, op.cbuf), buf_ptr(h, op.abuf, 0), buf_ld(h, op.abuf),
          buf_ptr(h, op.bbuf, op.dinv_tile), buf_ld(h, op.bbuf), op.alpha, op.beta);
      return 1;
    case OP_GEMM128_KK:
      tile_gemm_kernel<Cfg128, KCONTIG, KCONTIG><<<op.task_cnt, Cfg128::THREADS, Engine<Cfg128, KCONTIG, KCONTIG>::SMEM_BYTES, st>>>(
          h->tasks + op.task_off, buf_ptr(h, op.cbuf, 0), buf_ld(h, op.cbuf), buf_ptr(h, op.abuf, 0), buf_ld(h, op.abuf),
          buf_ptr(h, op.bbuf, op.dinv_tile), buf_ld(h, op.bbuf), op.alpha, op.beta);
      return 1;
    case OP_GEMM64_KM:
      tile_gemm_kernel<Cfg64, KCONTIG, MCONTIG><<<op.task_cnt, Cfg64::THREADS, Engine<Cfg64, KCONTIG, MCONTIG>::SMEM_BYTES, st>>>(
          h->tasks + op.task_off, buf_ptr(h, op.cbuf, 0), buf_ld(h, op.cbuf), buf_ptr(h, op.abuf, 0), buf_ld(h, op.abuf),
          buf_ptr(h, op.bbuf, op.dinv_tile), buf_ld(h, op.bbuf), op.alpha, op.beta);
      return 1;
    case OP_GEMM128_KM:
      tile_gemm_kernel<Cfg128, KCONTIG, MCONTIG><<<op.task_cnt, Cfg128::THREADS, Engine<Cfg128, KCONTIG, MCONTIG>::SMEM_BYTES, st>>>(
          h->tasks + op.task_off, buf_ptr(h, op.cbuf, 0), buf_ld(h, op.cbuf), buf_ptr(h, op.abuf, 0), buf_ld(h, op.abuf),
          buf_ptr(h, op.bbuf, op.dinv_tile), buf_ld(h, op.bbuf), op.alpha, op.beta);
      return 1;
    case OP_TRMV:
      trmv_lower_kernel<<<np / 8, 256, 0, st>>>(h->A, np, np, h->r, h->z);
      return 1;
    case OP_TRMV_T:
      trmv_t_partial_kernel<<<dim3(h->nT, h->nT), 128, 0, st>>>(h->A, np, np, h->z, h->tpartial);
      trmv_t_reduce_kernel<<<h->nT, 128, 0, st>>>(h->tpartial, np, h->alpha);
      return 2;
    case OP_LAUUM:
      DP_SWITCH(h->dp, (lauum_trace_kernel<DP><<<op.task_cnt, 256, LauumSmem<DP>::BYTES, st>>>(
                           h->tasks + op.task_off, h->A, np, h->X, h->alpha, h->kp, h->n, h->trace_partial, kinv_out)));
      return 1;
    case OP_FINALIZE:
      finalize_kernel<<<1, 256, 0, st>>>(h->logdet_part, h->nt, h->z, h->n, h->trace_partial, h->n_lauum, h->dp, h->d,
                                         want_grad, h->info, h->abort_flag, h->out_dev);
      return 1;
  }
  return 0;
}

int enqueue_eval(eb200_gp* h, cudaStream_t st, int want_grad, int max_stage, double* kinv_out, int* launches) {
  int cnt = 0;
  CK(cudaMemsetAsync(h->info, 0, sizeof(int), st));
  CK(cudaMemsetAsync(h->abort_flag, 0, sizeof(int), st));
  for (const Op& op : h->ops) {
    if (op.stage > max_stage) continue;
    if (op.grad_only && !want_grad) continue;
    cnt += launch_op(h, op, st, want_grad, kinv_out);
  }
  CK(cudaGetLastError());
  if (launches) *launches = cnt;
  return 0;
}

int capture_graph(eb200_gp* h, int want_grad, cudaGraphExec_t* exec, int* launches) {
  cudaGraph_t graph;
  CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  int rc = enqueue_eval(h, h->stream, want_grad, 99, nullptr, launches);
  cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
  if (rc) return rc;
  CK(e);
  CK(cudaGraphInstantiate(exec, graph, 0));
  CK(cudaGraphDestroy(graph));
  return 0;
}

int set_attrs() {
  static bool done_dev[64] = {false};
  int dev = 0;
  CK(cudaGetDevice(&dev));
  bool& done = done_dev[dev & 63];
  if (done) return 0;
  CK((set_gemm_attr<Cfg64, KCONTIG, KCONTIG>()));
  CK((set_gemm_attr<Cfg128, KCONTIG, KCONTIG>()));
  CK((set_gemm_attr<Cfg64, KCONTIG, MCONTIG>()));
  CK((set_gemm_attr<Cfg128, KCONTIG, MCONTIG>()));
  CK(cudaFuncSetAttribute(var_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          Engine<Cfg128, KCONTIG, KCONTIG>::SMEM_BYTES));
  CK(cudaFuncSetAttribute(diag_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_INV_SMEM_BYTES));
#define SET_POTRF(DPV)                                                                                       \
  CK(cudaFuncSetAttribute(potrf_dataflow_kernel<DPV>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                          PotrfSmem<DPV>::BYTES));
  SET_POTRF(2) SET_POTRF(3) SET_POTRF(4) SET_POTRF(5) SET_POTRF(6) SET_POTRF(8) SET_POTRF(9) SET_POTRF(10)
  SET_POTRF(12) SET_POTRF(16)
#undef SET_POTRF
#define SET_LAUUM(DPV)                                                                                       \
  CK(cudaFuncSetAttribute(lauum_trace_kernel<DPV>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                          LauumSmem<DPV>::BYTES));
  SET_LAUUM(2) SET_LAUUM(3) SET_LAUUM(4) SET_LAUUM(5) SET_LAUUM(6) SET_LAUUM(8) SET_LAUUM(9) SET_LAUUM(10)
  SET_LAUUM(12) SET_LAUUM(16)
#undef SET_LAUUM
  done = true;
  return 0;
}

__global__ void pad_rows_kernel(const double* __restrict__ src, int m, int d, int dp, double* __restrict__ dst) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)m * dp) return;
  int c = (int)(e % dp);
  size_t r = e / dp;
  dst[e] = (c < d) ? src[r * d + c] : 0.0;
}

int ensure_var_tasks(eb200_gp* h, int mt) {
  if (h->var_tasks_mt == mt) return 0;
  std::vector<TileTask> v;
  for (int M = 0; M < mt; ++M)
    for (int I = 0; I < h->nT; ++I) v.push_back({M * 128, I * 128, M * 128, 0, I * 128, 0, (I + 1) * 128 / KC, 0});
  sort_tasks(v, 0);
  CK(cudaMemcpyAsync(h->var_tasks, v.data(), v.size() * sizeof(TileTask), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));  // v goes out of scope
  h->var_tasks_mt = mt;
  h->var_tasks_cnt = (int)v.size();
  return 0;
}

// Predict one chunk (mc <= np query rows) whose padded coordinates already sit in h->T_dev.
int predict_chunk(eb200_gp* h, int mc, int want_var, double* mu_dev, double* var_dev, cudaStream_t st) {
  const int np = h->np;
  const int blocks = (mc + 31) / 32;
  double* ks = want_var ? h->W : nullptr;
  DP_SWITCH(h->dp, (predict_mean_kernel<DP><<<blocks, 256, 0, st>>>(h->T_dev, mc, h->X, h->alpha, h->kp, h->n, np, mu_dev, ks, np)));
  if (want_var) {
    const int mt = (mc + 127) / 128;
    if (ensure_var_tasks(h, mt)) return 1;
    var_gemm_kernel<<<h->var_tasks_cnt, 256, Engine<Cfg128, KCONTIG, KCONTIG>::SMEM_BYTES, st>>>(h->var_tasks, h->W, np, h->A, np,
                                                                                                 h->rowss, np);
    var_finish_kernel<<<(mc + 255) / 256, 256, 0, st>>>(h->rowss, np, h->nT, h->kp, mc, var_dev);
  }
  CK(cudaGetLastError());
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

int eb200_gp_create(eb200_gp** out, int device, int n, int d, const double* x_host, const double* noise_host) {
  if (!out || !x_host || !noise_host) return fail("eb200_gp_create: null argument");
  if (n < 1) return fail("eb200_gp_create: n must be >= 1");
  if (d < 1 || d > EB200_MAX_DIM) return fail("eb200_gp_create: kernel dimension must be in [1, 16]");
  CK(cudaSetDevice(device));
  if (set_attrs()) return 1;
  eb200_gp* h = new eb200_gp();
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
  CK(cudaMalloc(&h->A, mat));
  CK(cudaMalloc(&h->W, mat));
  CK(cudaMalloc(&h->X, (size_t)np * dp * sizeof(double)));
  CK(cudaMalloc(&h->noise, np * sizeof(double)));
  CK(cudaMalloc(&h->r, np * sizeof(double)));
  CK(cudaMalloc(&h->z, np * sizeof(double)));
  CK(cudaMalloc(&h->alpha, np * sizeof(double)));
  CK(cudaMalloc(&h->dinv, (size_t)h->nt * TB * TB * sizeof(double)));
  CK(cudaMalloc(&h->logdet_part, h->nt * sizeof(double)));
  CK(cudaMalloc(&h->tpartial, (size_t)h->nT * np * sizeof(double)));
  CK(cudaMalloc(&h->p_dev, (MAXD + 1) * sizeof(double)));
  CK(cudaMalloc(&h->out_dev, (MAXD + 8) * sizeof(double)));
  CK(cudaMalloc(&h->kp, sizeof(KernelParams)));
  CK(cudaMalloc(&h->info, sizeof(int)));
  CK(cudaMallocHost(&h->h_p, (MAXD + 1) * sizeof(double)));
  CK(cudaMallocHost(&h->h_out, (MAXD + 8) * sizeof(double)));
  CK(cudaMalloc(&h->T_dev, (size_t)np * dp * sizeof(double)));
  CK(cudaMalloc(&h->T_in, (size_t)np * d * sizeof(double)));
  CK(cudaMalloc(&h->mu_dev, np * sizeof(double)));
  CK(cudaMalloc(&h->var_dev, np * sizeof(double)));
  CK(cudaMalloc(&h->rowss, (size_t)h->nT * np * sizeof(double)));
  CK(cudaMalloc(&h->var_tasks, (size_t)h->nT * h->nT * sizeof(TileTask)));
  CK(cudaMallocHost(&h->h_pred, (size_t)np * (d + 2) * sizeof(double)));
  CK(cudaMemset(h->r, 0, np * sizeof(double)));
  CK(cudaMemset(h->z, 0, np * sizeof(double)));
  CK(cudaMemset(h->alpha, 0, np * sizeof(double)));
  CK(cudaMemset(h->W, 0, mat));
  CK(cudaMemset(h->A, 0, mat));  // strictly-upper 64-blocks of diagonal 128-tiles must stay zero
  CK(cudaMalloc(&h->counter, sizeof(unsigned)));
  CK(cudaMalloc(&h->ready, (size_t)h->nt * h->nt * sizeof(int)));
  CK(cudaMalloc(&h->abort_flag, sizeof(int)));
  CK(cudaMalloc(&h->rdiag, (size_t)np * sizeof(double)));
  CK(cudaMemset(h->abort_flag, 0, sizeof(int)));
  CK(cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, device));
  {
    std::vector<int2> pt;
    for (int j = 0; j < h->nt; ++j)
      for (int i = j; i < h->nt; ++i) pt.push_back(make_int2(i, j));
    h->n_ptiles = (int)pt.size();
    CK(cudaMalloc(&h->ptiles, pt.size() * sizeof(int2)));
    CK(cudaMemcpy(h->ptiles, pt.data(), pt.size() * sizeof(int2), cudaMemcpyHostToDevice));
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
  if (capture_graph(h, 1, &h->g_grad, &h->launches_grad)) return 1;
  if (capture_graph(h, 0, &h->g_lml, &h->launches_lml)) return 1;
  *out = h;
  return 0;
}

int eb200_gp_destroy(eb200_gp* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->g_grad) cudaGraphExecDestroy(h->g_grad);
  if (h->g_lml) cudaGraphExecDestroy(h->g_lml);
  void* dev[] = {h->A, h->W, h->X, h->noise, h->r, h->z, h->alpha, h->dinv, h->logdet_part, h->tpartial,
                 h->trace_partial, h->p_dev, h->out_dev, h->kp, h->info, h->tasks, h->ktiles, h->T_dev, h->T_in,
                 h->mu_dev, h->var_dev, h->rowss, h->var_tasks, h->ptiles, h->counter, h->ready, h->abort_flag, h->rdiag};
  for (void* p : dev)
    if (p) cudaFree(p);
  if (h->h_p) cudaFreeHost(h->h_p);
  if (h->h_out) cudaFreeHost(h->h_out);
  if (h->h_pred) cudaFreeHost(h->h_pred);
  cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int eb200_gp_n(const eb200_gp* h) { return h ? h->n : -1; }
int eb200_gp_dim(const eb200_gp* h) { return h ? h->d : -1; }
int eb200_gp_padded_n(const eb200_gp* h) { return h ? h->np : -1; }
int eb200_gp_eval_launches(const eb200_gp* h, int want_grad) {
  return h ? (want_grad ? h->launches_grad : h->launches_lml) : -1;
}

int eb200_gp_set_residual_host(eb200_gp* h, const double* r_host) {
  if (!h || !r_host) return fail("set_residual: null argument");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  memcpy(h->h_pred, r_host, h->n * sizeof(double));
  CK(cudaMemcpyAsync(h->r, h->h_pred, h->n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->factor_valid = false;
  return 0;
}

int eb200_gp_set_residual_device(eb200_gp* h, const double* r_dev, void* stream) {
  if (!h || !r_dev) return fail("set_residual: null argument");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  CK(cudaMemcpyAsync(h->r, r_dev, h->n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  h->factor_valid = false;
  return 0;
}

int eb200_gp_eval_device(eb200_gp* h, const double* p_dev, int want_grad, double* result_dev, void* stream) {
  if (!h || !p_dev) return fail("eval: null argument");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  CK(cudaMemcpyAsync(h->p_dev, p_dev, (h->d + 1) * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CK(cudaGraphLaunch(want_grad ? h->g_grad : h->g_lml, st));
  if (result_dev)
    CK(cudaMemcpyAsync(result_dev, h->out_dev, EB200_RESULT_LEN(h->d) * sizeof(double), cudaMemcpyDeviceToDevice, st));
  h->factor_valid = true;
  return 0;
}

int eb200_gp_eval_host(eb200_gp* h, const double* p_host, int want_grad, double* result_host) {
  if (!h || !p_host || !result_host) return fail("eval: null argument");
  CK(cudaSetDevice(h->device));
  const int len = EB200_RESULT_LEN(h->d);
  memcpy(h->h_p, p_host, (h->d + 1) * sizeof(double));
  CK(cudaMemcpyAsync(h->p_dev, h->h_p, (h->d + 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaGraphLaunch(want_grad ? h->g_grad : h->g_lml, h->stream));
  CK(cudaMemcpyAsync(h->h_out, h->out_dev, len * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  memcpy(result_host, h->h_out, len * sizeof(double));
  if (!want_grad)
    for (int i = 0; i <= h->d; ++i) result_host[1 + i] = 0.0;
  h->factor_valid = true;
  return 0;
}

int eb200_gp_predict_device(eb200_gp* h, int m, const double* t_dev, int want_var, double* mu_dev, double* var_dev,
                            void* stream) {
  if (!h || !t_dev || !mu_dev || (want_var && !var_dev)) return fail("predict: null argument");
  if (!h->factor_valid) return fail("predict: no factorisation resident (call eval first)");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  for (int q0 = 0; q0 < m; q0 += h->np) {
    const int mc = std::min(h->np, m - q0);
    const size_t tot = (size_t)mc * h->dp;
    pad_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(t_dev + (size_t)q0 * h->d, mc, h->d, h->dp, h->T_dev);
    if (predict_chunk(h, mc, want_var, mu_dev + q0, want_var ? var_dev + q0 : nullptr, st)) return 1;
  }
  return 0;
}

int eb200_gp_predict_host(eb200_gp* h, int m, const double* t_host, int want_var, double* mu_host, double* var_host) {
  if (!h || !t_host || !mu_host || (want_var && !var_host)) return fail("predict: null argument");
  if (!h->factor_valid) return fail("predict: no factorisation resident (call eval first)");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int d = h->d;
  for (int q0 = 0; q0 < m; q0 += h->np) {
    const int mc = std::min(h->np, m - q0);
    double* stage_t = h->h_pred;
    double* stage_mu = h->h_pred + (size_t)h->np * d;
    double* stage_var = stage_mu + h->np;
    memcpy(stage_t, t_host + (size_t)q0 * d, (size_t)mc * d * sizeof(double));
    CK(cudaMemcpyAsync(h->T_in, stage_t, (size_t)mc * d * sizeof(double), cudaMemcpyHostToDevice, st));
    const size_t tot = (size_t)mc * h->dp;
    pad_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(h->T_in, mc, d, h->dp, h->T_dev);
    if (predict_chunk(h, mc, want_var, h->mu_dev, h->var_dev, st)) return 1;
    CK(cudaMemcpyAsync(stage_mu, h->mu_dev, mc * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (want_var) CK(cudaMemcpyAsync(stage_var, h->var_dev, mc * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(mu_host + q0, stage_mu, mc * sizeof(double));
    if (want_var) memcpy(var_host + q0, stage_var, mc * sizeof(double));
  }
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
  if (enqueue_eval(h, h->stream, stage >= 4, stage, stage >= 4 ? h->W : nullptr, nullptr)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  h->factor_valid = false;
  return 0;
}

int eb200_gp_debug_matrix(eb200_gp* h, int which, double* out_host) {
  if (!h || !out_host) return fail("debug_matrix: null argument");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(out_host, which ? h->W : h->A, (size_t)h->np * h->np * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int eb200_gp_debug_alpha(eb200_gp* h, double* alpha_host) {
  if (!h || !alpha_host) return fail("debug_alpha: null argument");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(alpha_host, h->alpha, h->n * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int eb200_gp_debug_potrf_trace(eb200_gp* h, const double* p_host, unsigned long long* out_host, int* tiles_out) {
  if (!h || !p_host || !out_host || !tiles_out) return fail("potrf_trace: null argument");
  CK(cudaSetDevice(h->device));
  CK(cudaMalloc(&h->trace, (size_t)h->n_ptiles * 6 * sizeof(unsigned long long)));
  CK(cudaMemset(h->trace, 0, (size_t)h->n_ptiles * 6 * sizeof(unsigned long long)));
  memcpy(h->h_p, p_host, (h->d + 1) * sizeof(double));
  CK(cudaMemcpyAsync(h->p_dev, h->h_p, (h->d + 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  int rc = enqueue_eval(h, h->stream, 0, 2, nullptr, nullptr);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (!rc && e == cudaSuccess) {
    e = cudaMemcpy(out_host, h->trace, (size_t)h->n_ptiles * 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(tiles_out, h->ptiles, (size_t)h->n_ptiles * sizeof(int2), cudaMemcpyDeviceToHost);
  }
  cudaFree(h->trace);
  h->trace = nullptr;
  h->factor_valid = false;
  if (rc) return rc;
  CK(e);
  return 0;
}

int eb200_gp_profile_stages(eb200_gp* h, const double* p_host, int reps, float* ms_out) {
  if (!h || !p_host || !ms_out || reps < 1) return fail("profile_stages: bad argument");
  CK(cudaSetDevice(h->device));
  memcpy(h->h_p, p_host, (h->d + 1) * sizeof(double));
  CK(cudaMemcpyAsync(h->p_dev, h->h_p, (h->d + 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  cudaEvent_t ev[6];
  for (auto& e : ev) CK(cudaEventCreate(&e));
  for (int s = 0; s < 5; ++s) ms_out[s] = 0.f;
  for (int rep = 0; rep <= reps; ++rep) {  // rep 0 is a warm-up
    CK(cudaMemsetAsync(h->info, 0, sizeof(int), h->stream));
    CK(cudaEventRecord(ev[0], h->stream));
    for (int stage = 1; stage <= 5; ++stage) {
      for (const Op& op : h->ops)
        if (op.stage == stage) launch_op(h, op, h->stream, 1, nullptr);
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
  return 0;
}

int eb200_bench_gemm_nt(int m, int n, int k, const double* a_dev, const double* b_dev, double* c_dev, void* stream) {
  if (m % 128 || n % 128 || k % KC) return fail("bench_gemm_nt: m, n must be multiples of 128 and k of 16");
  if (set_attrs()) return 1;
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
  tile_gemm_kernel<Cfg128, KCONTIG, KCONTIG>
      <<<(m / 128) * (n / 128), Cfg128::THREADS, Engine<Cfg128, KCONTIG, KCONTIG>::SMEM_BYTES, (cudaStream_t)stream>>>(
          d_tasks, c_dev, n, a_dev, k, b_dev, k, 1.0, 0.0);
  CK(cudaGetLastError());
  return 0;
}

}  // extern "C"
