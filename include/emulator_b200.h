/*
 * emulator_b200 -- C ABI of the B200-native Gaussian-process hot path behind swiftemulator.
 *
 * This is the drop-in boundary: a plain C interface (pointers and sizes, no torch or C++
 * types) that a `george.GP`-shaped host object binds to.  Each entry point names the
 * reference call it replaces; paths are relative to the SWIFTSIM/emulator source tree and
 * `george` refers to the third-party package the reference delegates to (george 0.4.x,
 * un-vendored; see DESIGN.md "Oracle").
 *
 * Conventions
 *   - all floating point data is IEEE fp64, row-major, densely packed
 *   - hyperparameter vector p has P = d + 1 entries: [log c, log M_0 .. log M_{d-1}]
 *     (george `ConstantKernel(log_constant) * ExpSquaredKernel(metric=M, ndim=d)`,
 *     built at swiftemulator/emulators/gaussian_process.py:181-183)
 *   - every function returns 0 on success, non-zero on failure; eb200_last_error() gives text
 *   - `info` follows LAPACK dpotrf: 0 = factorised, k > 0 = leading minor k not positive
 *     definite (the host raises numpy.linalg.LinAlgError, as george/scipy do)
 *   - *_host entry points take host buffers and include the host<->device copies;
 *     *_device entry points take device pointers, enqueue on `stream` (a cudaStream_t cast to
 *     void*, NULL = the handle's own stream) and do not synchronise
 */
#ifndef EMULATOR_B200_H
#define EMULATOR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct eb200_gp eb200_gp;

#define EB200_MAX_DIM 16
/* length of the result vector written by the eval entry points for a handle of dimension d:
 * [lml, grad_0 .. grad_d, info, logdet, quad]  ->  d + 5 doubles */
#define EB200_RESULT_LEN(d) ((d) + 5)

int eb200_version(void);
const char* eb200_last_error(void);
int eb200_device_count(void);
/* free / total device memory in bytes (the host side keeps the number of resident GP handles within it) */
int eb200_device_mem_info(int device, unsigned long long* free_bytes, unsigned long long* total_bytes);

/* Replaces george.GP.compute(x, yerr)  (swiftemulator/emulators/gaussian_process.py:203-206,
 * gaussian_process_bins.py:239-242, gaussian_process_mcmc.py:235-238,
 * gaussian_process_one_dim.py:202-205).  Uploads the design matrix x[n][d] and the diagonal
 * term noise[n] (= yerr^2 + white-noise variance, formed by the host exactly as george does)
 * once and allocates all device state (K/L/L^-1 matrix, scratch, task lists, CUDA graphs). */
int eb200_gp_create(eb200_gp** out, int device, int n, int d, const double* x_host, const double* noise_host);
int eb200_gp_destroy(eb200_gp* gp);
int eb200_gp_n(const eb200_gp* gp);
int eb200_gp_dim(const eb200_gp* gp);

/* Residual r = y - mean(x), the vector george forms in GP.log_likelihood / _compute_alpha. */
int eb200_gp_set_residual_host(eb200_gp* gp, const double* r_host);
int eb200_gp_set_residual_device(eb200_gp* gp, const double* r_dev, void* stream);

/* Replaces the pair gp.set_parameter_vector(p); gp.log_likelihood(y) /
 * gp.grad_log_likelihood(y)  (gaussian_process.py:208-214 and the three sibling emulators):
 * rebuilds K at p, factorises, and returns the log marginal likelihood and (EB200_EVAL_GRAD)
 * its gradient with respect to p in one pass.  result has EB200_RESULT_LEN(d) entries.
 *   EB200_EVAL_LML         value; L^-1 and alpha stay resident for eb200_gp_predict_*
 *   EB200_EVAL_GRAD        value + gradient; L^-1 and alpha stay resident
 *   EB200_EVAL_VALUE_ONLY  value only through the Cholesky factor alone (a third of the flops):
 *                          the evaluation the ensemble sampler makes per walker
 *                          (gaussian_process_mcmc.py:240-265); leaves NO prediction state */
#define EB200_EVAL_LML 0
#define EB200_EVAL_GRAD 1
#define EB200_EVAL_VALUE_ONLY 2
int eb200_gp_eval_host(eb200_gp* gp, const double* p_host, int mode, double* result_host);
int eb200_gp_eval_device(eb200_gp* gp, const double* p_dev, int mode, double* result_dev, void* stream);

/* One evaluation (any mode) on each of nb DISTINCT handles -- the per-bin GPs of GaussianProcessEmulatorBins
 * (gaussian_process_bins.py:206-262), the leave-one-out refits of CrossCheck (sensitivity/cross_check.py:98-114),
 * optimiser restarts: problems the reference works through one after the other.  Equally shaped problems of
 * 14 or more 64-row tile rows share launches, up to 8 at a time (4 above 4096 rows): their Cholesky dependency chains are
 * interleaved in one factorisation launch (one chain cannot feed 148 SMs, several can) and every later stage is
 * one launch over the group; smaller problems are enqueued on the handles' own streams.  Results are
 * bit-identical to eb200_gp_eval_host on each handle.  Problem b reads p_host + b * p_stride and writes
 * result_host + b * result_stride.  A handle listed twice is an error. */
int eb200_gp_eval_multi_host(eb200_gp** gps, int nb, const double* p_host, int p_stride, int mode, double* result_host,
                             int result_stride);
/* The same with the parameter vectors and the result vectors in device memory; everything is enqueued on
 * `stream` (NULL = the first handle's own stream) and nothing synchronises.  All handles on one device.
 * As with eb200_gp_eval_device the caller reads `info` from the result vector before predicting. */
int eb200_gp_eval_multi_device(eb200_gp** gps, int nb, const double* p_dev, int p_stride, int mode, double* result_dev,
                               int result_stride, void* stream);

/* nb value-only evaluations (EB200_EVAL_VALUE_ONLY semantics) at the parameter vectors p_host[nb][d+1];
 * result_host[nb][EB200_RESULT_LEN(d)].  A value-only factorisation is bound by its dependency chain
 * and leaves most SMs idle, so up to eight share one launch.  This is the half-ensemble of proposals
 * the ensemble sampler evaluates per move (gaussian_process_mcmc.py:259-265; emcee's `vectorize`). */
int eb200_gp_eval_value_batch_host(eb200_gp* gp, int nb, const double* p_host, double* result_host);

/* Replaces gp.predict(y, t, return_cov=False, return_var=bool)  (gaussian_process.py:285-287,
 * :344-346; gaussian_process_bins.py:354-359, :452-457).  Requires a preceding eval at the
 * hyperparameters to predict with in mode EB200_EVAL_LML or EB200_EVAL_GRAD (L^-1 and alpha stay
 * resident).  mean-model values
 * are added by the host.  t is [m][d]; mu and var are [m]; var may be NULL if !want_var. */
int eb200_gp_predict_host(eb200_gp* gp, int m, const double* t_host, int want_var, double* mu_host,
                          double* var_host);
int eb200_gp_predict_device(eb200_gp* gp, int m, const double* t_dev, int want_var, double* mu_dev,
                            double* var_dev, void* stream);

/* Prediction on a grid: every parameter row theta[r][0..d-2] at every independent value ind[b]; query row
 * r * n_ind + b is (ind[b], theta[r][:]) -- the rows mock_sweep / mock_hypercube and the Sobol sweep
 * build on the host one emulator call at a time (mocking/__init__.py:84-95, :187-198).  The rows are
 * formed on the device (round_f32: rounded to float32 first, as the reference's emulators do with
 * their query points); only theta, ind and the results cross the bus.  mu, var: [n_theta * n_ind]. */
int eb200_gp_predict_grid_host(eb200_gp* gp, int n_theta, const double* theta_host, int n_ind, const double* ind_host,
                               int round_f32, int want_var, double* mu_host, double* var_host);

/* The same with theta, ind and the results in device memory (enqueued on `stream`, no synchronisation): what a
 * sharded sweep uses, whose per-rank results are all-gathered over NVLink from where they are
 *  and copied to the host once. */
int eb200_gp_predict_grid_device(eb200_gp* gp, long long n_theta, const double* theta_dev, int n_ind, const double* ind_dev,
                                 int round_f32, int want_var, double* mu_dev, double* var_dev, void* stream);

int eb200_gp_synchronize(eb200_gp* gp);

/* Test / diagnostics hooks (not used by the product path).
 * stage: 1 = K built into matrix 0 (diagnostic kernel; the product path generates K inside the
 * factorisation), 2 = Cholesky + triangular inverse (L in the lower triangle of matrix 0,
 * X = L^-1 in matrix 1), 3 = + alpha, 4 = + K^-1 lower triangle stored into matrix 0.
 * eb200_gp_debug_matrix copies matrix `which` (0 or 1), np*np doubles. */
int eb200_gp_debug_stage(eb200_gp* gp, const double* p_host, int stage);
int eb200_gp_padded_n(const eb200_gp* gp);
int eb200_gp_debug_matrix(eb200_gp* gp, int which, double* out_host);
int eb200_gp_debug_alpha(eb200_gp* gp, double* alpha_host);
/* number of kernel launches issued by one eval (graph nodes) */
int eb200_gp_eval_launches(const eb200_gp* gp, int mode);

/* Per-task globaltimer stamps (ns) of the dataflow factorisation kernel: out[task][10] =
 * {claimed, K generated / panels accumulated, C formed / diagonal inverse visible, diagonal
 * factor visible, tile stored, published, SM id, CTA id, ns spent waiting for input flags, ns yielded
 * to chain tiles}; tasks_out[task][6] = (type 0 Cholesky / 1 inverse / 2, 3 helpers / 4 diagonal inverse,
 * tile row, tile column, 0).  The number of tasks is eb200_gp_factor_tasks(). */
int eb200_gp_debug_factor_trace(eb200_gp* gp, const double* p_host, unsigned long long* out_host, int* tasks_out);
int eb200_gp_factor_tasks(const eb200_gp* gp);
/* the same for a group of nb equally shaped handles sharing one factorisation launch: out[nb][task][10] */
int eb200_gp_debug_group_trace(eb200_gp** gps, int nb, const double* p_host, int p_stride, unsigned long long* out_host,
                               int* tasks_out);
/* the same for a value-only evaluation (Cholesky tiles + the residual row, tile row nt) */
int eb200_gp_debug_value_trace(eb200_gp* gp, const double* p_host, unsigned long long* out_host, int* tasks_out);
int eb200_gp_value_tasks(const eb200_gp* gp);

/* Per-stage device time of one LML+grad evaluation (milliseconds, averaged over `reps` after
 * one warm-up), launched kernel by kernel outside the CUDA graph with events in between:
 * ms_out[0] = parameter prep, [1] = dataflow factorisation (kernel-matrix build + Cholesky +
 * triangular inverse), [2] = alpha, [3] = K^-1 tiles + gradient traces, [4] = final reduction. */
int eb200_gp_profile_stages(eb200_gp* gp, const double* p_host, int reps, float* ms_out);

/* The same for one grouped LML+grad evaluation of nb equally shaped handles (eb200_gp_eval_multi_host's shared
 * launches); ms_out as above, each figure covering the whole group. */
int eb200_gp_profile_group_stages(eb200_gp** gps, int nb, const double* p_host, int p_stride, int reps, float* ms_out);

/* Micro-benchmark of the dataflow kernel's 64x64 main loop in isolation: C[m][n] = A[m][k] * B[n][k]^T
 * (b_mcontig = 0) or A[m][k] * B[k][n] (b_mcontig = 1) with `grid` persistent CTAs. */
int eb200_bench_accumulate(int m, int n, int k, int b_mcontig, int grid, const double* a_dev, const double* b_dev,
                           double* c_dev, void* stream);

/* Engine micro-benchmark: C[m][n] = A[m][k] * B[n][k]^T through the tile engine, device
 * pointers, m, n multiples of 128, k multiple of 16.  Used by bench.py to report the engine's
 * DMMA throughput next to cuBLAS DGEMM. */
int eb200_bench_gemm_nt(int m, int n, int k, const double* a_dev, const double* b_dev, double* c_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EMULATOR_B200_H */
