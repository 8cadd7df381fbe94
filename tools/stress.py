"""Repeat evaluations at several sizes and modes; every result must be bit-identical to the first and
info must stay 0 (development tool: exercises the dataflow kernel's claiming under timing noise)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from emulator_b200.device import DeviceProblem
from emulator_b200 import synthetic
from oracle import gp_oracle as orc
bad = 0
for n, d, reps in [(int(a), int(b), int(c)) for a, b, c in (t.split(":") for t in (sys.argv[1] if len(sys.argv) > 1 else "4096:9:300,2500:8:300,640:9:1500,1536:5:800,200:3:3000").split(","))]:
    x, y, yerr = synthetic.design_arrays(n, d, seed=n)
    gp = DeviceProblem(x.astype(np.float64), orc.noise_diagonal(yerr)); gp.set_residual(y.astype(np.float64))
    p = orc.default_parameter_vector(d) + 0.1
    ref_g = gp.evaluate(p, True); ref_v = gp.evaluate(p, False, value_only=True)
    ps = p + 0.05 * np.random.default_rng(0).standard_normal((11, d + 1))
    ref_b = gp.evaluate_value_batch(ps)
    t0 = time.perf_counter()
    for k in range(reps):
        g = gp.evaluate(p, True)
        if g[2] != 0 or g[0] != ref_g[0] or not np.array_equal(g[1], ref_g[1]):
            bad += 1; print('grad mismatch', k, g[2], g[0] - ref_g[0], np.max(np.abs(g[1] - ref_g[1])), g[3] - ref_g[3], g[4] - ref_g[4], flush=True)
        if k % 3 == 0:
            v = gp.evaluate(p, False, value_only=True)
            if v[2] != 0 or v[0] != ref_v[0]:
                bad += 1; print('value mismatch', k, v[2], v[0] - ref_v[0], v[3] - ref_v[3], v[4] - ref_v[4], flush=True)
        if k % 7 == 0:
            b = gp.evaluate_value_batch(ps)
            if np.any(b[1] != 0) or not np.array_equal(b[0], ref_b[0]):
                bad += 1; print('batch mismatch', k, b[1], b[0] - ref_b[0], b[2] - ref_b[2], b[3] - ref_b[3], flush=True)
    print(json.dumps({"n": n, "d": d, "reps": reps, "seconds": time.perf_counter() - t0, "bad": bad}), flush=True)
    gp.close()
print("STRESS", "OK" if bad == 0 else f"FAILED ({bad})")
