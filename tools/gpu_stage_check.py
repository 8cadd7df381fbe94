"""Stage-by-stage check of the CUDA path against numpy on a real GPU (development tool)."""
import sys, time, json, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emulator_b200.device import DeviceProblem
from oracle import gp_oracle as orc
from scipy.linalg import cholesky, solve_triangular


def make(n, d, seed=0, err=0.02):
    rng = np.random.default_rng(seed)
    x = rng.random((n, d)).astype(np.float32).astype(np.float64)
    y = np.sin(x.sum(1) * 2.0) + 0.3 * x[:, 0] ** 2
    yerr = (err * np.abs(y) + 1e-3).astype(np.float32)
    y = y.astype(np.float32).astype(np.float64)
    return x, y, orc.noise_diagonal(yerr)


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def check(n, d, seed=0, stages=True):
    x, y, diag = make(n, d, seed)
    p = orc.default_parameter_vector(d) + 0.1 * np.random.default_rng(seed + 1).standard_normal(d + 1)
    gp = DeviceProblem(x, diag)
    gp.set_residual(y)
    npad = gp.padded_n
    out = {"n": n, "d": d, "np": npad}
    K = orc.kernel_value(p, x); K[np.diag_indices(n)] += diag
    if stages:
        gp.debug_stage(p, 1)
        A = gp.debug_matrix(0)
        out["K_err"] = rel(np.tril(A[:n, :n]), np.tril(K))
        L = cholesky(K, lower=True)
        gp.debug_stage(p, 2)
        A = gp.debug_matrix(0)
        out["L_err"] = rel(np.tril(A[:n, :n]), L)
        out["L_pad_ok"] = bool(np.allclose(np.tril(A[n:, :]), np.eye(npad)[n:, :]))
        Xi = solve_triangular(L, np.eye(n), lower=True)
        Wm = gp.debug_matrix(1)
        out["X_err"] = rel(np.tril(Wm[:n, :n]), Xi)
        out["X_upper_zero"] = bool(np.all(Wm[np.triu_indices(npad, 1)] == 0))
        gp.debug_stage(p, 3)
        al = gp.debug_alpha()
        out["alpha_err"] = rel(al, np.linalg.solve(K, y))
        gp.debug_stage(p, 4)
        Wm = gp.debug_matrix(0)
        Kinv = Xi.T @ Xi
        out["Kinv_err"] = rel(np.tril(Wm[:n, :n]), np.tril(Kinv))
    lml, grad, info, logdet, quad = gp.evaluate(p, True)
    if n <= 2048:
        rl, rg, rld, rq = orc.lml_and_grad(p, x, diag, y)
    else:
        rl, rg, rld, rq = orc.lml_and_grad_lean(p, x, diag, y)
    out.update(info=info, lml=lml, lml_ref=float(rl), lml_rel=abs(lml - rl) / abs(rl), logdet_rel=abs(logdet - rld) / abs(rld),
               quad_rel=abs(quad - rq) / abs(rq), grad_rel=float(np.max(np.abs(grad - rg)) / np.max(np.abs(rg))))
    lml2, _, _, _, _ = gp.evaluate(p, False)
    out["lml_nograd_same"] = bool(lml2 == lml)
    # predict
    t = np.random.default_rng(seed + 2).random((37 if n < 2000 else 300, d)).astype(np.float32).astype(np.float64)
    mu, var = gp.predict(t, True)
    rmu, rvar = orc.predict(p, x, diag, y, t)
    amp = orc.amplitude(p, d)
    out["mu_rel"] = rel(mu, rmu); out["var_abs_over_amp"] = float(np.max(np.abs(var - rvar)) / amp)
    mu2 = gp.predict(t, False)
    out["mu_novar_same"] = bool(np.array_equal(mu, mu2))
    # timing
    for wg in (True, False):
        gp.evaluate(p, wg)
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            gp.evaluate(p, wg)
        out["eval_ms_grad" if wg else "eval_ms_lml"] = (time.perf_counter() - t0) / reps * 1e3
    out["launches"] = gp.launches_per_eval(True)
    gp.close()
    return out


if __name__ == "__main__":
    cases = [(100, 2), (200, 3), (1000, 3), (1500, 9)]
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        cases += [(4096, 9)]
    for n, d in cases:
        r = check(n, d, stages=(n <= 2048))
        print(json.dumps(r), flush=True)
