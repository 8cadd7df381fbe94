"""Run a few LML+grad evaluations at (n, d) -- profiling target (development tool)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from emulator_b200.device import DeviceProblem
from oracle import gp_oracle as orc
n, d, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kind = sys.argv[4] if len(sys.argv) > 4 else "grad"   # grad | lml | value
want_grad, value_only = kind == "grad", kind == "value"
rng = np.random.default_rng(0)
x = rng.random((n, d)).astype(np.float32).astype(np.float64)
y = np.sin(x.sum(1) * 2.0)
diag = orc.noise_diagonal((0.02 * np.abs(y) + 1e-3).astype(np.float32))
gp = DeviceProblem(x, diag); gp.set_residual(y)
p = orc.default_parameter_vector(d)
if kind.startswith("batch"):   # batchNB: NB value-only evaluations per call
    nb = int(kind[5:])
    ps = p + 0.05 * np.random.default_rng(1).standard_normal((nb, d + 1))
    gp.evaluate_value_batch(ps)
    t0 = time.perf_counter()
    for _ in range(reps):
        out = gp.evaluate_value_batch(ps)
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"n": n, "d": d, "kind": kind, "ms_per_call": dt * 1e3, "ms_per_eval": dt * 1e3 / nb, "info": int(out[1].max())}))
    sys.exit(0)
gp.evaluate(p, want_grad, value_only=value_only)  # warm-up
t0 = time.perf_counter()
for _ in range(reps):
    out = gp.evaluate(p, want_grad, value_only=value_only)
dt = (time.perf_counter() - t0) / reps
print(json.dumps({"n": n, "d": d, "ms": dt * 1e3, "lml": out[0], "info": out[2], "kind": kind, "launches": gp.launches_per_eval(want_grad, value_only)}))
