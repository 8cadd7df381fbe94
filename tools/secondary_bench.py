"""
Secondary metrics of BASELINE.json on one GPU (development / reporting tool):
predictions/s against the C2 GP, fit times (C1, one C2 fit, C3 bins), one C4-sized refit evaluation.
Writes one JSON document to stdout.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from emulator_b200 import synthetic
from emulator_b200.device import DeviceProblem
from emulator_b200.emulators import GaussianProcessEmulator, GaussianProcessEmulatorBins

out = {}
# ---- predictions/s against a C2-sized GP (device-resident queries, CUDA events) ----
n, d = 4096, 9
x, y, yerr = synthetic.design_arrays(n, d, seed=1235)
noise = np.sqrt(np.ascontiguousarray(yerr ** 2, dtype=np.float64) + 1.25e-12) ** 2
gp = DeviceProblem(x.astype(np.float64), noise); gp.set_residual(y.astype(np.float64))
p = synthetic.perturbed_parameter_vectors(d, 1)[0]
gp.evaluate(p, False)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
for m, want_var in ((1 << 16, True), (1 << 18, False)):
    t = torch.rand((m, d), dtype=torch.float64, device=dev)
    mu = torch.empty(m, dtype=torch.float64, device=dev); var = torch.empty(m, dtype=torch.float64, device=dev)
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        gp.predict_device(m, t.data_ptr(), want_var, mu.data_ptr(), var.data_ptr(), stream.cuda_stream)
        e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    key = "predict_mean_var" if want_var else "predict_mean"
    out[key] = {"rows": m, "ms": ms, "rows_per_s": m / ms * 1e3}
    if want_var:
        out[key]["tflops_variance_gemm"] = float(n) * n * m / ms / 1e9
# host-buffer path (what predict_values uses)
th = np.random.default_rng(0).random((1 << 14, d))
t0 = time.perf_counter(); gp.predict(th, True); out["predict_mean_var_host"] = {"rows": len(th), "rows_per_s": len(th) / (time.perf_counter() - t0)}
gp.close()
# ---- fits ----
def fit(name, cls, cfg):
    spec, params, values = synthetic.make_config(cfg)
    emu = cls()
    t0 = time.perf_counter(); emu.fit_model(spec, params, values); dt = time.perf_counter() - t0
    r = {"seconds": dt}
    if hasattr(emu, "fit_result") and emu.fit_result is not None:
        r.update(nfev=int(emu.fit_result.nfev), njev=int(emu.fit_result.njev), nit=int(emu.fit_result.nit),
                 device_evals=int(emu.emulator.n_device_evaluations))
    out[name] = r
fit("fit_C1_gpe_N1000", GaussianProcessEmulator, "C1")
fit("fit_C2_gpe_N4096", GaussianProcessEmulator, "C2")
fit("fit_C3_bins_64xN256", GaussianProcessEmulatorBins, "C3")
# the same with all bins advancing in lock step (emulator_b200/optimise.py)
spec3, params3, values3 = synthetic.make_config("C3")
GaussianProcessEmulatorBins(fit_mode="lockstep").fit_model(spec3, params3, values3)  # warm-up
emu3 = GaussianProcessEmulatorBins(fit_mode="lockstep")
t0 = time.perf_counter(); emu3.fit_model(spec3, params3, values3); dt3 = time.perf_counter() - t0
out["fit_C3_bins_64xN256_lockstep"] = {"seconds": dt3, "rounds": int(emu3.lockstep_result.rounds),
                                       "device_evals": int(emu3.lockstep_result.nfev.sum())}
# ---- one C4-sized evaluation (N = 8176) ----
x, y, yerr = synthetic.design_arrays(8176, 9, seed=1238)
noise = np.sqrt(np.ascontiguousarray(yerr ** 2, dtype=np.float64) + 1.25e-12) ** 2
gp = DeviceProblem(x.astype(np.float64), noise); gp.set_residual(y.astype(np.float64))
p = synthetic.perturbed_parameter_vectors(9, 1)[0]
gp.evaluate(p, True)
t0 = time.perf_counter()
for _ in range(3): gp.evaluate(p, True)
ms = (time.perf_counter() - t0) / 3 * 1e3
out["eval_C4_N8176"] = {"ms": ms, "tflops": 8176.0 ** 3 / ms / 1e9, "projected_refit_s_at_150_evals": 150 * ms / 1e3,
                        "projected_loo_512_refits_8gpu_min": 512 * 150 * ms / 1e3 / 8 / 60}
gp.close()
# ---- value-only evaluations (what the ensemble sampler asks for): Cholesky factor alone ----
for n_, d_ in ((4096, 9), (8176, 9), (256, 8)):
    x, y, yerr = synthetic.design_arrays(n_, d_, seed=1239)
    noise = np.sqrt(np.ascontiguousarray(yerr ** 2, dtype=np.float64) + 1.25e-12) ** 2
    gp = DeviceProblem(x.astype(np.float64), noise); gp.set_residual(y.astype(np.float64))
    p = synthetic.perturbed_parameter_vectors(d_, 1)[0]
    res = {}
    for kind, kw in (("value_only", dict(want_grad=False, value_only=True)), ("lml_with_prediction_state", dict(want_grad=False)),
                     ("lml_and_grad", dict(want_grad=True))):
        gp.evaluate(p, **kw)
        reps = 20 if n_ <= 4096 else 5
        t0 = time.perf_counter()
        for _ in range(reps): gp.evaluate(p, **kw)
        res[kind + "_ms"] = (time.perf_counter() - t0) / reps * 1e3
    out[f"eval_modes_N{n_}"] = res
    gp.close()
    gp = DeviceProblem(x.astype(np.float64), noise); gp.set_residual(y.astype(np.float64))
    ps = p + 0.05 * np.random.default_rng(2).standard_normal((40, d_ + 1))
    gp.evaluate_value_batch(ps)
    t0 = time.perf_counter(); gp.evaluate_value_batch(ps); dt = time.perf_counter() - t0
    out[f"eval_modes_N{n_}"]["value_only_batched_ms_per_eval"] = dt / 40 * 1e3
    gp.close()
# ---- ensemble-MCMC emulator (reference defaults: 40 walkers, 50 + 100 steps = 6000 value-only likelihoods) ----
from emulator_b200.emulators import GaussianProcessEmulatorMCMC
spec, params, values = synthetic.make_config("C1")
emu = GaussianProcessEmulatorMCMC(seed=0)
t0 = time.perf_counter(); emu.fit_model(spec, params, values); dt = time.perf_counter() - t0
out["fit_C1_mcmc_N1000_6000_likelihoods"] = {"seconds": dt, "device_evals": int(emu.emulator.n_device_evaluations)}
# ---- leave-one-out refits at C4 size (512 sims x 16 points; each refit N = 8176): two measured refits,
#      sequentially and from two host threads (CrossCheck's `workers`), projected to all 512 ----
from concurrent.futures import ThreadPoolExecutor
from emulator_b200.backend.model_values import ModelValues
spec, params, values = synthetic.make_config("C4")
uids = list(values.model_values.keys())
def refit(uid):
    emu = GaussianProcessEmulator()
    t0 = time.perf_counter()
    emu.fit_model(spec, params, ModelValues({k: v for k, v in values.model_values.items() if k != uid}))
    return time.perf_counter() - t0, int(emu.fit_result.nfev), int(emu.emulator.n_device_evaluations)
r0 = refit(uids[0])
t0 = time.perf_counter()
with ThreadPoolExecutor(max_workers=2) as pool:
    r12 = list(pool.map(refit, uids[1:3]))
pair = time.perf_counter() - t0
out["loo_C4_refit_N8176"] = {"one_refit_seconds": r0[0], "nfev": r0[1], "device_evals": r0[2],
                             "two_refits_two_threads_seconds": pair, "their_device_evals": [r[2] for r in r12],
                             "projected_512_refits_8gpu_min": 512 / 8 * min(r0[0], pair / 2) / 60}
print(json.dumps(out, indent=1))
