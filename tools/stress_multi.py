"""Repeat batched multi-handle evaluations (shared factorisation launches) and compare with each handle's
stand-alone evaluation, bit for bit (development tool)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from emulator_b200.device import DeviceProblem, evaluate_many
from emulator_b200 import synthetic
from oracle import gp_oracle as orc
bad = 0
for n, d, nb, reps in ((1000, 3, 11, 300), (2048, 9, 5, 60), (900, 16, 8, 200), (1500, 5, 16, 100)):
    probs, ps, refs = [], [], []
    for k in range(nb):
        x, y, yerr = synthetic.design_arrays(n, d, seed=31 * n + k)
        q = DeviceProblem(x.astype(np.float64), orc.noise_diagonal(yerr)); q.set_residual(y.astype(np.float64))
        p = orc.default_parameter_vector(d) + 0.2 * np.random.default_rng(k).standard_normal(d + 1)
        probs.append(q); ps.append(p); refs.append(q.evaluate(p, True))
    ps = np.array(ps)
    t0 = time.perf_counter()
    for r in range(reps):
        lml, grad, info = evaluate_many(probs, ps, want_grad=True)
        for k in range(nb):
            if info[k] != 0 or lml[k] != refs[k][0] or not np.array_equal(grad[k], refs[k][1]):
                bad += 1
                print("mismatch", n, d, r, k, info[k], lml[k] - refs[k][0], flush=True)
    print(json.dumps({"n": n, "d": d, "nb": nb, "reps": reps, "seconds": time.perf_counter() - t0, "bad": bad}), flush=True)
    for q in probs: q.close()
print("STRESS", "OK" if bad == 0 else f"FAILED ({bad})")
