"""Randomised parity sweep against the oracle: random sizes, dimensions, hyperparameters, all evaluation
modes, prediction and grid prediction (development tool; the fixed cases live in tests/)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from emulator_b200.device import DeviceProblem
from emulator_b200 import synthetic
from oracle import gp_oracle as orc
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst = {"lml": 0.0, "grad": 0.0, "value_only": 0.0, "batch": 0.0, "mu": 0.0, "var": 0.0}
bad = 0
for c in range(cases):
    n = int(rng.choice([rng.integers(1, 70), rng.integers(60, 300), rng.integers(250, 1400)]))
    d = int(rng.integers(1, 17))
    x = rng.random((n, d)).astype(np.float32).astype(np.float64)
    y = np.sin(3.0 * x.sum(axis=1)) + 0.1 * rng.standard_normal(n)
    yerr = (0.02 * np.abs(y) + 1e-3).astype(np.float32)
    noise = orc.noise_diagonal(yerr)
    gp = DeviceProblem(x, noise); gp.set_residual(y)
    p = orc.default_parameter_vector(d) + 0.5 * rng.standard_normal(d + 1)
    try:
        ref = orc.lml_and_grad(p, x, noise, y)
    except np.linalg.LinAlgError:
        info = gp.evaluate(p, True)[2]
        if info == 0: bad += 1; print("case", c, n, d, "oracle failed, device did not")
        gp.close(); continue
    lml, grad, info, logdet, quad = gp.evaluate(p, True)
    if info != 0: bad += 1; print("case", c, n, d, "device info", info); gp.close(); continue
    worst["lml"] = max(worst["lml"], abs(lml - ref[0]) / abs(ref[0]))
    gerr = np.max(np.abs(grad - ref[1])) / np.max(np.abs(ref[1]))
    worst["grad"] = max(worst["grad"], gerr)
    if gerr > 1e-9 or abs(lml - ref[0]) / abs(ref[0]) > 1e-10:
        # how well is the problem conditioned, and how much does the ORACLE move under a row permutation?
        k = orc.kernel_value(p, x); k[np.diag_indices(n)] += noise
        perm = rng.permutation(n)
        ref2 = orc.lml_and_grad(p, x[perm], noise[perm], y[perm])
        print(json.dumps({"case": c, "n": n, "d": d, "grad_err": gerr, "lml_err": abs(lml - ref[0]) / abs(ref[0]), "cond_K": float(np.linalg.cond(k)),
                          "oracle_self_noise_grad": float(np.max(np.abs(ref2[1] - ref[1])) / np.max(np.abs(ref[1]))),
                          "oracle_self_noise_lml": float(abs(ref2[0] - ref[0]) / abs(ref[0]))}), flush=True)
    v = gp.evaluate(p, False, value_only=True)
    worst["value_only"] = max(worst["value_only"], abs(v[0] - ref[0]) / abs(ref[0]))
    ps = p + 0.1 * rng.standard_normal((3, d + 1))
    b = gp.evaluate_value_batch(ps)
    for k in range(3):
        try:
            rb = orc.lml_and_grad(ps[k], x, noise, y)[0]
            if b[1][k] == 0: worst["batch"] = max(worst["batch"], abs(b[0][k] - rb) / abs(rb))
        except np.linalg.LinAlgError:
            pass
    gp.evaluate(p, False)
    m = int(rng.integers(1, 40))
    t = rng.random((m, d)).astype(np.float32).astype(np.float64)
    mu, var = gp.predict(t, True)
    rmu, rvar = orc.predict(p, x, noise, y, t)
    worst["mu"] = max(worst["mu"], np.max(np.abs(mu - rmu)) / max(np.max(np.abs(rmu)), 1e-300))
    worst["var"] = max(worst["var"], np.max(np.abs(var - rvar)) / orc.amplitude(p, d))
    if d >= 2:
        th = rng.random((3, d - 1)); ind = rng.random(4)
        rows = np.column_stack([np.tile(ind, 3), np.repeat(th, 4, axis=0)]).astype(np.float32).astype(np.float64)
        g_mu, g_var = gp.predict_grid(th, ind, True)
        r_mu, r_var = gp.predict(rows, True)
        if not (np.array_equal(g_mu.ravel(), r_mu) and np.array_equal(g_var.ravel(), r_var)):
            bad += 1; print("case", c, n, d, "grid prediction differs")
    gp.close()
print(json.dumps({"cases": cases, "bad": bad, "worst_relative": worst}))
