"""C3 (64 bins x N = 256, 8 parameters): per-bin fits by scipy BFGS from host threads vs all bins in lock step
(development tool)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from emulator_b200 import synthetic
from emulator_b200.emulators import GaussianProcessEmulatorBins
spec, params, values = synthetic.make_config("C3")
out = {}
for mode, kw in (("scipy_8_threads", dict(fit_mode="scipy", workers=8)), ("scipy_1_thread", dict(fit_mode="scipy", workers=1)),
                 ("lockstep", dict(fit_mode="lockstep"))):
    emu = GaussianProcessEmulatorBins(**kw)
    emu.fit_model(spec, params, values)  # warm-up (library load, allocations)
    emu = GaussianProcessEmulatorBins(**kw)
    t0 = time.perf_counter(); emu.fit_model(spec, params, values); dt = time.perf_counter() - t0
    lml = [g.log_likelihood(v["dependent"]) for g, v in zip(emu.bin_gaussian_process, emu.bin_model_values)]
    out[mode] = {"seconds": dt, "sum_lml_at_optimum": float(np.sum(lml)), "device_evals": int(sum(g.n_device_evaluations for g in emu.bin_gaussian_process))}
    if emu.lockstep_result is not None:
        r = emu.lockstep_result
        out[mode].update(rounds=int(r.rounds), status_counts=[int((r.status == k).sum()) for k in range(3)], max_nit=int(r.nit.max()))
print(json.dumps(out, indent=1))
