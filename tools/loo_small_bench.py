"""Leave-one-out cross-check at C1 size (50 simulations x 20 points: 50 refits of N = 980, d = 3): per-refit
scipy BFGS vs all refits in lock step (development tool)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from emulator_b200 import synthetic
from emulator_b200.sensitivity import CrossCheck
spec, params, values = synthetic.make_config("C1")
out = {}
for mode, kw in (("scipy_2_threads", dict(fit_mode="scipy", workers=2)), ("lockstep", dict(fit_mode="lockstep"))):
    CrossCheck(**kw).build_emulators(spec, params, values)  # warm-up
    cc = CrossCheck(**kw)
    t0 = time.perf_counter(); cc.build_emulators(spec, params, values); dt = time.perf_counter() - t0
    t0 = time.perf_counter(); ms, _ = cc.get_mean_squared(); dt2 = time.perf_counter() - t0
    out[mode] = {"build_seconds": dt, "score_seconds": dt2, "mean_squared": float(ms),
                 "device_evals": int(sum(e.emulator.n_device_evaluations for e in cc.cross_emulators.values()))}
    if cc.lockstep_result is not None:
        out[mode]["rounds"] = int(cc.lockstep_result.rounds)
print(json.dumps(out, indent=1))
