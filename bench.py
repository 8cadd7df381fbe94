#!/usr/bin/env python
"""
Benchmark of the GP hot path on BASELINE.json's configurations, synthetic seeded data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--skip-secondary]

HEADLINE (the JSON line's metric/value/e2e): log-marginal-likelihood + gradient evaluations per second at
configuration C2 (N = 4096 training rows, n_par = 8 -> kernel dimension d = 9, P = 10 hyperparameters, fp64).
One "step" = one pass of the hot path over a batch of BATCH independent hyperparameter vectors (kernel matrix
build, Cholesky, triangular inverse, alpha, K^-1 tiles fused with the gradient traces, final reduction) -- what
the reference does as one `fun(p)` + one `jac(p)` call per vector
(swiftemulator/emulators/gaussian_process.py:208-221).  The vectors of a step are independent (optimiser
restarts, leave-one-out refits): they are evaluated on BATCH GP handles holding the same data, all eight
sharing launches (eb200_gp_eval_multi_*: interleaved factorisations; bit-identical to stand-alone evaluations,
asserted below).

value  : whole-job evals/s with inputs resident in HBM (hyperparameter vectors on the device, results left on
         the device), timed with CUDA events on the launching stream.
e2e    : the same metric through the host-buffer C ABI call (eb200_gp_eval_multi_host): every step copies its
         hyperparameter vectors host->device from pinned memory and reads the result vectors device->host inside
         the timed region.
N > 1  : one process per GPU (torchrun); every rank owns independent problems of the same shape -- weak scaling,
         no data-path collective; barriers on both sides, max over ranks.

SECONDARY BLOCKS in the same JSON line (BASELINE.json's other two metrics, the sharded configurations; both are
STRONG scaling: the same total work at every N, through the product's drivers, collectives inside the timed
region, max over ranks):
sweep_c5 : configuration C5, the Sobol sweep -- 2^20 parameter rows x 32 independent values predicted (mean +
           variance) against the C2 emulator through `sensitivity.sobol.sweep_predict` (rows split over ranks,
           results all-gathered from device memory over NCCL, one copy to the host).
loo_c4   : configuration C4, leave-one-out cross check at N = 8176 -- `CrossCheck.build_emulators` over a fixed
           subset of the 512 simulations (refits handed out dynamically to ranks, scipy BFGS per refit, likelihood
           evaluations of concurrent refits combined into shared launches).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core,
# and BLAS reads the variable when numpy is imported.
# (only that arm: eight ranks of OUR arm with 16 BLAS threads each would fight over the host cores their optimisers need)
if os.environ.get("OMP_NUM_THREADS") == "1" and "TORCHELASTIC_RUN_ID" in os.environ and "reference" in sys.argv:
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAIN, N_PAR = 4096, 8
DIM = N_PAR + 1
BATCH = 8  # independent evaluations per step
WORKLOAD = "C2: LML+grad, N=4096 (256 sims x 16 points), n_par=8 (kernel d=9, P=10), fp64"
METRIC = "GP LML+grad evals/s (N=4096, d=8, fp64)"
# identical in both arms (the driver compares the dicts); what an arm samples of it is in its cpu_baseline.sample
CONFIG = {
    "workload": WORKLOAD,
    "batch": "8 independent hyperparameter vectors per step, drawn N(x0, 0.5^2), seed 7",
    "per_gpu": "independent problems of this shape on every rank (weak scaling)",
    "cache": "working set 2 x 134 MB per problem exceeds the 126 MB L2 (inputs larger than L2)",
}
SWEEP_LOG2_ROWS, SWEEP_IND = 20, 32
LOO_REFITS = 64  # 8 per rank on 8 GPUs (C4 itself: 64 per rank); BFGS takes 30 .. 100+ evaluations per refit, so fewer per rank mostly measures that spread


def make_inputs(seed_offset=0):
    from emulator_b200 import synthetic

    x, y, yerr = synthetic.design_arrays(N_TRAIN, DIM, seed=1235 + seed_offset)
    yerr2 = np.ascontiguousarray(yerr ** 2, dtype=np.float64)  # squared in float32, as george does
    noise = np.sqrt(yerr2 + 1.25e-12) ** 2
    ps = synthetic.perturbed_parameter_vectors(DIM, count=32, seed=7)[1:]  # 32 vectors around x0
    return x.astype(np.float64), y.astype(np.float64), noise, yerr, ps


class ClockSampler(threading.Thread):
    """nvidia-smi clock / throttle-reason samples while a timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, enabled=True, interval=0.2):
        super().__init__(daemon=True)
        self.interval = interval
        # (rank 0 only: every nvidia-smi call takes the driver's global lock for milliseconds -- eight ranks polling five
        #  times a second stretched the launch -> synchronise round trips of the leave-one-out block several-fold)
        self.index, self.samples, self._stop_evt, self.enabled = index, [], threading.Event(), enabled

    def run(self):
        while self.enabled and not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(self.interval)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(len(s) > col and s[col].lower().startswith("active") for s in self.samples):
                reasons.append(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(self.samples[0][1]) if self.samples else None,
            "power_w_max": max((float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()), default=None),
            "reasons": reasons,
            "samples": len(self.samples),
        }


def cpu_reference_evals(x, y, noise, ps, evals, warm=1, lean=False):
    """George-faithful oracle (or, lean=True, the same algorithm as the GPU path: no N x N x P tensor, one
    triangular inverse) on the host cores: returns (evals/s, seconds per eval list)."""
    from oracle import gp_oracle

    fn = gp_oracle.lml_and_grad_lean if lean else gp_oracle.lml_and_grad
    times = []
    for i in range(warm + evals):
        p = ps[i % len(ps)]
        t0 = time.perf_counter()
        fn(p, x, noise, y)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return len(times) / sum(times), times


def blas_threads():
    try:
        from threadpoolctl import threadpool_info

        return max((i.get("num_threads", 1) for i in threadpool_info()), default=os.cpu_count())
    except Exception:
        return os.cpu_count()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    x, y, noise, _, ps = make_inputs()
    cores = blas_threads()
    t_all = time.perf_counter()
    rate, times = cpu_reference_evals(x, y, noise, ps, evals=args.steps, warm=args.warmup)
    sample = (f"{len(times)} LML+grad evals (1 per step) of the george-faithful numpy/scipy oracle at N=4096, d=9 "
              f"(full K^-1 by N solves, materialised N x N x P gradient tensor), OpenBLAS on {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": dict(CONFIG),
        "cpu_baseline": {"value": rate, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "george is not installable here; oracle/gp_oracle.py restates it (checked against scikit-learn, tests/test_oracle_vs_sklearn.py)"},
        "e2e": {"value": rate, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------
def dgemm_peak(torch, dev):
    """cuBLAS DGEMM 8192^3 on this GPU, TFLOP/s (MEASURED_PEAKS.json has no fp64 entry)."""
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    c = torch.empty(8192, 8192, dtype=torch.float64, device=dev)
    best = 1e30
    for it in range(4):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); torch.matmul(a, b, out=c); s1.record(); torch.cuda.synchronize()
        if it:
            best = min(best, s0.elapsed_time(s1))
    del a, b, c
    return 2 * 8192.0 ** 3 / best / 1e9


def ncu_traffic(kernel_key):
    """DRAM bytes per launch of a kernel from the committed ncu --set full summary (None if absent)."""
    import csv

    for name in ("r02_top_kernels_ncu.csv", "r01_top_kernels_ncu.csv"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                rows = list(csv.DictReader(fh))
            for row in rows[1:]:
                if kernel_key in row["Kernel Name"]:
                    return (float(row["dram__bytes_read.sum"]) + float(row["dram__bytes_write.sum"])) * 1e6, name
        except Exception:
            continue
    return None, None


def timed_collective_region(torch, td, world, dev, fn):
    """Run fn() between barriers; returns (result, device milliseconds by CUDA events on the current stream, wall
    milliseconds), both as the max over ranks."""
    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    out = fn()
    e1.record()
    barrier()
    wall = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e0.elapsed_time(e1), wall], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    ev_ms, wall_ms = (float(v) for v in t.cpu())
    return out, ev_ms, wall_ms


def block_sweep_c5(torch, td, world, rank, local, dev, peak_tflops, log2_rows):
    """BASELINE.json configuration C5 through the product's sweep driver (strong scaling)."""
    from emulator_b200 import synthetic
    from emulator_b200.emulators import GaussianProcessEmulator
    from emulator_b200.emulators.base import build_gp, default_kernel
    from emulator_b200.sensitivity import sobol

    spec, params, values = synthetic.make_config("C2")
    emu = GaussianProcessEmulator(device=local)
    emu._build_arrays(spec, params, values)
    emu.kernel = default_kernel(spec.number_of_parameters + 1)
    gp = build_gp(emu.kernel, None, emu.independent_variables, emu.dependent_variables, emu.dependent_variable_errors, device=local)
    p_star = synthetic.perturbed_parameter_vectors(DIM, 1)[1]  # the same hyperparameters on every rank: bitwise-equal replicas
    gp.set_parameter_vector(p_star)
    emu.emulator = gp
    rows = 1 << log2_rows
    base = max(1, min(1 << 16, rows // (spec.number_of_parameters + 2)))
    design = sobol.saltelli_design(spec, base, seed=11)  # A, B and the 8 AB_i blocks: 10 * base rows
    top_up = rows - len(design)
    if top_up > 0:  # further rows of the same kind up to exactly 2^k
        extra = sobol.saltelli_design(spec, -(-top_up // (spec.number_of_parameters + 2)), seed=12)[:top_up]
        design = np.vstack([design, extra])
    theta = np.ascontiguousarray(design[:rows])
    ind = np.linspace(0.0, 1.0, SWEEP_IND)
    from emulator_b200 import dist

    sobol.sweep_predict(emu, theta[: 1 << 12], ind, want_var=True)  # warm-up: factorisation, allocations, NCCL
    # host staging of the two result arrays pinned ahead of time -- what any sweep after a process's first one finds
    dist.pinned_pool.reserve((rows, SWEEP_IND), 2)
    sampler = ClockSampler(local, enabled=rank == 0, interval=0.2 if world == 1 else 0.5)
    sampler.start()
    (mean, var), ev_ms, wall_ms = timed_collective_region(torch, td, world, dev, lambda: sobol.sweep_predict(emu, theta, ind, want_var=True))
    clocks = sampler.stop()
    m = rows * SWEEP_IND
    n = N_TRAIN
    flops = float(n) * n * m  # variance: one triangular product per query row
    out = {
        "workload": f"C5: Sobol sweep, 2^{log2_rows} parameter rows x {SWEEP_IND} independent values (mean+var) against the C2 GP (N=4096, d=9)",
        "metric": "predictions/s (mean+var)", "value": m / (ev_ms / 1e3), "unit": "query rows/s", "seconds": ev_ms / 1e3,
        "wall_seconds": wall_ms / 1e3, "query_rows": m, "scaling": "strong", "n_gpus": world,
        "timed": "sensitivity.sobol.sweep_predict between barriers: H2D of this rank's parameter rows, query rows formed on the device, "
                 "mean + variance kernels, NCCL all-gather of both results from device memory, one D2H copy each; CUDA events, max over ranks",
        "h2d_bytes": int(theta.nbytes // world + ind.nbytes), "d2h_bytes": int(2 * m * 8),
        "roofline": {"bound": "tensor", "kernel": "var_grid_fused_kernel (K* rebuilt in shared memory from the factorised tables, V = K* L^-T by TMA-fed "
                     "DMMA.8x8x4 128x128 tiles, reduced to row norms; K* and V never in global memory)",
                     "flops": flops, "achieved": flops / ev_ms / 1e9 / world, "peak": peak_tflops, "unit": "TFLOP/s per GPU",
                     "frac": flops / ev_ms / 1e9 / world / peak_tflops, "traffic": ncu_traffic("var_grid_fused")[0],
                     "note": "N^2 flop per query row over the WHOLE sweep (host copies, mean kernel, all-gather included)"},
        "checksum": float(mean.sum() + var.sum()), "clocks": clocks,
    }
    if world == 1 and rank == 0:
        from oracle import gp_oracle

        x = emu.independent_variables.astype(np.float64)
        y = emu.dependent_variables.astype(np.float64)
        diag = gp_oracle.noise_diagonal(emu.dependent_variable_errors)
        sl = 128  # theta rows -> 4096 query rows
        t_rows = np.empty((sl * SWEEP_IND, DIM), dtype=np.float32)
        t_rows[:, 0] = np.tile(ind, sl)
        t_rows[:, 1:] = np.repeat(theta[:sl], SWEEP_IND, axis=0)
        t0 = time.perf_counter()
        ref_mu, ref_var = gp_oracle.predict(p_star, x, diag, y, t_rows.astype(np.float64))
        dt = time.perf_counter() - t0
        amp = gp_oracle.amplitude(p_star, DIM)
        out["cpu_baseline"] = {"value": sl * SWEEP_IND / dt, "unit": "query rows/s", "cores": blas_threads(), "kind": "port",
                               "sample": f"{sl * SWEEP_IND} query rows (the first {sl} parameter rows) through oracle predict, factorisation included; {dt:.2f} s"}
        out["parity_on_cpu_sample"] = {
            "mean_rel": float(np.max(np.abs(mean[:sl].ravel() - ref_mu)) / np.max(np.abs(ref_mu))),
            "var_rel_amp": float(np.max(np.abs(var[:sl].ravel() - ref_var)) / amp),
        }
        assert out["parity_on_cpu_sample"]["mean_rel"] <= 1e-9 and out["parity_on_cpu_sample"]["var_rel_amp"] <= 1e-9
    gp.close()
    return out


def block_loo_c4(torch, td, world, rank, local, dev, peak_tflops, refits):
    """BASELINE.json configuration C4 (a fixed subset of its 512 refits) through CrossCheck (strong scaling)."""
    from emulator_b200 import dist, synthetic
    from emulator_b200.sensitivity import CrossCheck

    spec, params, values = synthetic.make_config("C4")
    uids = list(values.model_values.keys())[:refits]
    cc = CrossCheck(device=local, leave_out=uids)
    sampler = ClockSampler(local, enabled=rank == 0, interval=0.2 if world == 1 else 0.5)
    sampler.start()
    _, ev_ms, wall_ms = timed_collective_region(torch, td, world, dev, lambda: cc.build_emulators(spec, params, values))
    clocks = sampler.stop()
    stats = cc.build_stats
    totals = dist.all_reduce_sum(np.array([stats["device_evaluations"], stats["combined_calls"], stats["refits"]], dtype=np.float64))
    native_s = dist.all_reduce_sum(np.eye(world)[rank] * stats["combined_seconds"])
    per_rank = dist.all_reduce_sum(np.eye(world)[rank] * stats["refits"])
    sizes = [len(v["independent"]) for v in values.model_values.values()]
    n_loo = sum(sizes) - sizes[0]  # training rows of one leave-one-out refit
    flops = totals[0] * float(n_loo) ** 3
    out = {
        "workload": f"C4: CrossCheck.build_emulators, {refits} of the 512 leave-one-out refits at N={n_loo} (d=9), scipy BFGS each",
        "metric": "LOO refits/s", "value": refits / (ev_ms / 1e3), "unit": "refits/s", "seconds": ev_ms / 1e3, "wall_seconds": wall_ms / 1e3,
        "scaling": "strong", "n_gpus": world, "refits": refits, "refits_per_rank": [int(v) for v in per_rank],
        "device_evaluations": int(totals[0]), "combined_native_calls": int(totals[1]),
        "seconds_inside_native_calls_per_rank": [round(float(v), 3) for v in native_s],
        "timed": "CrossCheck.build_emulators between barriers: array building, handle creation, BFGS (host), likelihood evaluations "
                 "(device, concurrent refits share launches), all-reduce of the optima; CUDA events on an idle stream = wall clock, max over ranks",
        "roofline": {"bound": "tensor", "flops": flops, "achieved": flops / ev_ms / 1e9 / world, "peak": peak_tflops, "unit": "TFLOP/s per GPU",
                     "frac": flops / ev_ms / 1e9 / world / peak_tflops, "note": "N^3 flop per LML+grad evaluation, host time included"},
        "projected_512_refits_s": 512.0 / (refits / (ev_ms / 1e3)),
        "checksum": float(cc.cross_fit_parameters.sum()), "clocks": clocks,
    }
    if world == 1 and rank == 0:
        from oracle import gp_oracle

        emu = cc.cross_emulators.built_items()[0][1]
        x, y = emu.independent_variables.astype(np.float64), emu.dependent_variables.astype(np.float64)
        diag = gp_oracle.noise_diagonal(emu.dependent_variable_errors)
        # (not at the optimum, where the gradient is ~0 and a relative comparison means nothing)
        p = emu.emulator.get_parameter_vector() + 0.05
        t0 = time.perf_counter()
        ref_lml, ref_grad, _, _ = gp_oracle.lml_and_grad_lean(p, x, diag, y)
        dt = time.perf_counter() - t0
        evals_per_refit = totals[0] / refits
        out["cpu_baseline"] = {"value": 1.0 / (dt * evals_per_refit), "unit": "refits/s", "cores": blas_threads(), "kind": "port",
                               "sample": f"1 LML+grad evaluation of the lean oracle (the GPU path's algorithm; the george-faithful one needs a 5.3 GB "
                                         f"gradient tensor here) at N={n_loo}: {dt:.1f} s, times the {evals_per_refit:.1f} evaluations BFGS took per refit"}
        emu.emulator.set_parameter_vector(p)
        lml = emu.emulator.log_likelihood(emu.dependent_variables)
        grad = emu.emulator.grad_log_likelihood(emu.dependent_variables)
        out["parity_on_cpu_sample"] = {"lml_rel": float(abs(lml - ref_lml) / abs(ref_lml)),
                                       "grad_rel": float(np.max(np.abs(grad - ref_grad)) / np.max(np.abs(ref_grad)))}
    for _, e in cc.cross_emulators.built_items():
        e.emulator.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as td

    from emulator_b200 import _lib
    from emulator_b200.device import DeviceProblem, evaluate_many, evaluate_many_device, profile_group_stages

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _lib.require_gpu()
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    x, y, noise, _, ps = make_inputs(seed_offset=rank)
    gps = []
    for _ in range(BATCH):
        gp = DeviceProblem(x, noise, device=local)
        gp.set_residual(y)
        gps.append(gp)
    P, L = DIM + 1, DIM + 5
    n_ps = len(ps)
    assert n_ps % BATCH == 0
    p_dev = torch.from_numpy(ps).to(dev)
    out_dev = torch.zeros((n_ps, L), dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream(device=dev)  # explicit stream: handle 0 would select the GP's own stream
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream

    def step_device(step):
        i = (step * BATCH) % n_ps
        evaluate_many_device(gps, p_dev[i].data_ptr(), P, True, out_dev[i].data_ptr(), L, sptr)

    host_results = np.zeros((n_ps, L))

    def step_host(step):
        i = (step * BATCH) % n_ps
        lml, grad, info = evaluate_many(gps, ps[i : i + BATCH], want_grad=True)
        host_results[i : i + BATCH, 0] = lml
        host_results[i : i + BATCH, 1 : P + 1] = grad
        host_results[i : i + BATCH, P + 1] = info

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) ----
    for w in range(args.warmup):
        step_device(w)
    barrier()
    sampler = ClockSampler(local, enabled=rank == 0, interval=0.2 if world == 1 else 0.5)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s in range(args.steps):
        step_device(s)
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    # ---- end-to-end timing through the host-buffer C ABI ----
    for w in range(min(args.warmup, 2)):
        step_host(w)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_host(s)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()

    times = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(times, op=td.ReduceOp.MAX)
    dev_ms, e2e_ms = (float(v) for v in times.cpu())
    evals = args.steps * BATCH * world

    # sanity: the device- and host-pointer paths agree bit for bit, and a grouped evaluation equals a stand-alone one
    res = out_dev.cpu().numpy()
    used = sorted({(s * BATCH + b) % n_ps for s in range(args.steps) for b in range(BATCH)})
    assert all(res[i, DIM + 2] == 0 for i in used), "Cholesky failed in the timed region"
    assert np.array_equal(res[used, : P + 2], host_results[used, : P + 2]), "device/host entry points disagree"
    alone = gps[3].evaluate(ps[used[3]], True)
    assert alone[0] == res[used[3], 0] and np.array_equal(alone[1], res[used[3], 1 : P + 1]), "grouped != stand-alone evaluation"

    line = None
    peak_tflops = dgemm_peak(torch, dev) if rank == 0 else 0.0
    pk = torch.tensor([peak_tflops], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(pk, op=td.ReduceOp.MAX)
    peak_tflops = float(pk.cpu()[0])
    if rank == 0:
        # ---- roofline of the dominant stage, live: per-stage CUDA-event times of one group + cuBLAS DGEMM peak ----
        group = BATCH  # the eight evaluations of a step share their launches
        stage_ms = profile_group_stages(gps[:group], ps[:group], reps=3) / group  # per evaluation
        single_ms = gps[0].profile_stages(ps[0], reps=3)
        names = ["params", "kbuild_cholesky_inverse", "alpha", "kinv_traces", "finalize"]
        n3 = float(N_TRAIN) ** 3
        stage_flops = [0.0, 2 * n3 / 3, 4.0 * N_TRAIN * N_TRAIN, n3 / 3, 0.0]
        dominant = int(np.argmax(stage_ms))
        eval_ms = dev_ms / (args.steps * BATCH)
        achieved_eval = n3 / eval_ms / 1e9
        achieved_stage = stage_flops[dominant] / stage_ms[dominant] / 1e9 if stage_ms[dominant] > 0 else 0.0
        traffic, traffic_src = ncu_traffic("factor_dataflow" if dominant == 1 else "lauum_trace")
        kernel_names = {1: f"factor_dataflow_kernel<9,true> (K build + Cholesky + triangular inverse of {group} interleaved problems)",
                        3: "lauum_trace_kernel (K^-1 tiles + traces)"}
        roofline = {
            "bound": "tensor", "kernel": kernel_names.get(dominant, names[dominant]) + f" -- DMMA.8x8x4 tile engine, one launch per group of {group} evaluations",
            "achieved": achieved_stage, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved_stage / peak_tflops,
            "launch": {"evaluations": group, "flops": stage_flops[dominant] * group, "ms": float(stage_ms[dominant] * group),
                       "note": f"achieved = flops / ms of one launch of the dominant kernel (a group of {group} evaluations), CUDA events"},
            "traffic": traffic, "traffic_source": f"profiles/{traffic_src} (ncu --set full, DRAM bytes per launch of {group} evaluations; algorithmic: "
                                                  f"{group} x 134 MB of L and Y written once -- interleaved problems no longer fit the 126 MB L2)" if traffic_src else None,
            "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul fp64) measured in this run; MEASURED_PEAKS.json has no fp64 entry",
            "whole_eval": {"flops": n3, "ms": eval_ms, "achieved": achieved_eval, "frac": achieved_eval / peak_tflops},
            "stage_ms_per_eval_grouped": {k: float(v) for k, v in zip(names, stage_ms)},
            "stage_ms_stand_alone": {k: float(v) for k, v in zip(names, single_ms)},
        }
        # ---- CPU baseline on this box's host cores (bounded sample) ----
        cpu_baseline = None
        if world == 1:  # reported at N=1 only (torchrun pins OMP_NUM_THREADS=1 on multi-rank launches)
            cores = blas_threads()
            rate, ctimes = cpu_reference_evals(x, y, noise, ps, evals=2, warm=1)
            lean_rate, lean_times = cpu_reference_evals(x, y, noise, ps, evals=2, warm=1, lean=True)
            cpu_baseline = {
                "value": rate, "unit": "evals/s", "cores": cores, "kind": "port",
                "sample": f"2 LML+grad evals (after 1 warm-up) of the george-faithful oracle at N=4096, d=9; {np.mean(ctimes):.2f} s each",
                # also the CPU running the GPU path's algorithm (no N x N x P tensor, one triangular inverse)
                "optimised_cpu_value": lean_rate, "optimised_cpu_sample": f"2 evals of the lean oracle, {np.mean(lean_times):.2f} s each",
            }
        kernels_per_group = 7  # parameters, factorisations, z, alpha (2), K^-1 + traces, final sums
        groups_per_step = BATCH // group
        line = {
            "metric": METRIC, "value": evals / (dev_ms / 1e3), "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": dict(CONFIG),
            "e2e": {"value": evals / (e2e_ms / 1e3), "unit": "evals/s", "h2d_bytes_per_step": BATCH * P * 8,
                    "d2h_bytes_per_step": BATCH * L * 8, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": kernels_per_group * groups_per_step * args.steps, "launches_per_step": kernels_per_group * groups_per_step,
            "evals_per_step": BATCH, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
    for gp in gps:
        gp.close()
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
    # ---- the sharded configurations (strong scaling; collectives inside the timed regions) ----
    if not args.skip_secondary:
        sweep = block_sweep_c5(torch, td, world, rank, local, dev, peak_tflops, args.sweep_log2_rows)
        loo = block_loo_c4(torch, td, world, rank, local, dev, peak_tflops, args.loo_refits)
        if rank == 0:
            line["sweep_c5"] = sweep
            line["loo_c4"] = loo
    if rank == 0:
        emit(line)
    if world > 1:
        td.barrier()
        td.destroy_process_group()


def emit(line: dict):
    """Write the one JSON line to the real stdout (fd saved before anything could print to it)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    # Libraries (NCCL prints its version banner at communicator creation) write to fd 1; keep the
    # contract of exactly one JSON line on stdout by pointing fd 1 at stderr for the run.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-secondary", action="store_true", help="headline metric only (no sweep_c5 / loo_c4 blocks)")
    ap.add_argument("--sweep-log2-rows", type=int, default=SWEEP_LOG2_ROWS)
    ap.add_argument("--loo-refits", type=int, default=LOO_REFITS)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
