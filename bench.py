#!/usr/bin/env python
"""
Headline benchmark: GP log-marginal-likelihood + gradient evaluations per second at
BASELINE.json configuration C2 (N = 4096 training rows, n_par = 8 -> kernel dimension d = 9,
P = 10 hyperparameters, fp64), synthetic seeded data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over a batch of BATCH hyperparameter vectors (kernel
matrix build, Cholesky, triangular inverse, alpha, K^-1 tiles fused with the gradient
traces, final reduction) -- what the reference does as one `fun(p)` + one `jac(p)` call per
vector (swiftemulator/emulators/gaussian_process.py:208-221).

value  : whole-job evals/s with inputs resident in HBM (hyperparameter vectors on the device,
         results left on the device), timed with CUDA events on the launching stream.
e2e    : the same metric through the host-buffer C ABI call (eb200_gp_eval_host): every
         evaluation copies its hyperparameter vector host->device from pinned memory and reads
         the result vector device->host inside the timed region.
N > 1  : one process per GPU (torchrun); every rank owns an independent problem of the same
         shape (optimiser restarts / leave-one-out refits shard like this) -- weak scaling, no
         data-path collective; the timed region is bracketed by barriers and the max over
         ranks is taken.
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
if os.environ.get("OMP_NUM_THREADS") == "1" and "TORCHELASTIC_RUN_ID" in os.environ:
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAIN, N_PAR = 4096, 8
DIM = N_PAR + 1
BATCH = 8  # evaluations per step
WORKLOAD = "C2: LML+grad, N=4096 (256 sims x 16 points), n_par=8 (kernel d=9, P=10), fp64"
METRIC = "GP LML+grad evals/s (N=4096, d=8, fp64)"


def make_inputs(seed_offset=0):
    from emulator_b200 import synthetic

    x, y, yerr = synthetic.design_arrays(N_TRAIN, DIM, seed=1235 + seed_offset)
    yerr2 = np.ascontiguousarray(yerr ** 2, dtype=np.float64)  # squared in float32, as george does
    noise = np.sqrt(yerr2 + 1.25e-12) ** 2
    ps = synthetic.perturbed_parameter_vectors(DIM, count=32, seed=7)
    return x.astype(np.float64), y.astype(np.float64), noise, yerr, ps


class ClockSampler(threading.Thread):
    """nvidia-smi clock / throttle-reason samples while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

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
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "evals_per_step": 1, "note": "george is not installable here; oracle/gp_oracle.py restates it"},
        "cpu_baseline": {"value": rate, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    emit(line)


def run_ours(args):
    import torch
    import torch.distributed as td

    from emulator_b200 import _lib
    from emulator_b200.device import DeviceProblem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _lib.require_gpu()
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    x, y, noise, _, ps = make_inputs(seed_offset=rank)
    gp = DeviceProblem(x, noise, device=local)
    gp.set_residual(y)
    P = DIM + 1
    n_ps = len(ps)
    p_dev = torch.from_numpy(ps).to(dev)
    out_dev = torch.zeros((n_ps, DIM + 5), dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream(device=dev)  # explicit stream: handle 0 would select the GP's own stream
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream

    def step_device(step):
        for b in range(BATCH):
            i = (step * BATCH + b) % n_ps
            gp.evaluate_device(p_dev[i].data_ptr(), True, out_dev[i].data_ptr(), sptr)

    host_results = np.zeros((n_ps, DIM + 5))

    def step_host(step):
        for b in range(BATCH):
            i = (step * BATCH + b) % n_ps
            lml, grad, info, _, _ = gp.evaluate(ps[i], True)
            host_results[i, 0] = lml
            host_results[i, 1 : P + 1] = grad

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) ----
    for w in range(args.warmup):
        step_device(w)
    barrier()
    sampler = ClockSampler(local)
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

    # sanity: device and host paths agree bit for bit with each other
    res = out_dev.cpu().numpy()
    used = sorted({(s * BATCH + b) % n_ps for s in range(args.steps) for b in range(BATCH)})
    assert all(res[i, DIM + 2] == 0 for i in used), "Cholesky failed in the timed region"
    assert np.array_equal(res[used, : P + 1], host_results[used, : P + 1]), "device/host entry points disagree"

    if rank == 0:
        # ---- roofline of the dominant stage, live: per-stage CUDA-event times + cuBLAS DGEMM peak ----
        stage_ms = gp.profile_stages(ps[0], reps=3)
        names = ["params", "kbuild_cholesky_inverse", "alpha", "kinv_traces", "finalize"]
        n3 = float(N_TRAIN) ** 3
        stage_flops = [0.0, 2 * n3 / 3, 4.0 * N_TRAIN * N_TRAIN, n3 / 3, 0.0]
        dominant = int(np.argmax(stage_ms))
        a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        c = torch.empty(8192, 8192, dtype=torch.float64, device=dev)
        best = 1e30
        for it in range(4):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(); torch.matmul(a, b, out=c); s1.record(); torch.cuda.synchronize()
            if it:
                best = min(best, s0.elapsed_time(s1))
        peak_tflops = 2 * 8192.0 ** 3 / best / 1e9
        del a, b, c
        eval_ms = dev_ms / (args.steps * BATCH)
        achieved_eval = n3 / eval_ms / 1e9
        achieved_stage = stage_flops[dominant] / stage_ms[dominant] / 1e9 if stage_ms[dominant] > 0 else 0.0
        traffic = None
        try:  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
            import csv
            with open(os.path.join(ROOT, "profiles", "r01_top_kernels_ncu.csv")) as fh:
                rows = list(csv.DictReader(fh))
            key = "factor_dataflow" if dominant == 1 else "lauum_trace"
            for row in rows[1:]:
                if key in row["Kernel Name"]:
                    traffic = (float(row["dram__bytes_read.sum"]) + float(row["dram__bytes_write.sum"])) * 1e6
        except Exception:
            traffic = None
        kernel_names = {1: "factor_dataflow_kernel (K build + Cholesky + triangular inverse)", 3: "lauum_trace_kernel (K^-1 tiles + traces)"}
        roofline = {
            "bound": "tensor", "kernel": kernel_names.get(dominant, names[dominant]) + " -- DMMA.8x8x4 tile engine, one launch per eval",
            "achieved": achieved_stage, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved_stage / peak_tflops,
            "traffic": traffic, "traffic_source": "profiles/r01_top_kernels_ncu.csv (ncu --set full, bytes per launch)",
            "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul fp64) measured in this run; MEASURED_PEAKS.json has no fp64 entry",
            "whole_eval": {"flops": n3, "ms": eval_ms, "achieved": achieved_eval, "frac": achieved_eval / peak_tflops},
            "stage_ms": {k: float(v) for k, v in zip(names, stage_ms)},
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
        launches = gp.launches_per_eval(True)
        line = {
            "metric": METRIC, "value": evals / (dev_ms / 1e3), "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "evals_per_step": BATCH, "per_gpu": "one independent problem per rank",
                       "cache": "working set 2 x 134 MB per problem exceeds the 126 MB L2 (inputs larger than L2)"},
            "e2e": {"value": evals / (e2e_ms / 1e3), "unit": "evals/s", "h2d_bytes_per_step": BATCH * P * 8,
                    "d2h_bytes_per_step": BATCH * (DIM + 5) * 8, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches * args.steps * BATCH, "launches_per_eval": launches,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        emit(line)
    gp.close()
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
