"""
Lock-step BFGS over many independent, equally shaped problems.

The reference fits its per-bin GPs one after the other with `scipy.optimize.minimize`
(/root/reference/swiftemulator/emulators/gaussian_process_bins.py:206-262): for 64 bins of 256 points
that is ~4500 evaluations of ~0.2 ms each on the device, every one wrapped in ~0.15 ms of Python
(scipy's BFGS bookkeeping), serialised by the interpreter lock whatever the number of threads.
Here all problems advance together: one round = one native call that evaluates value and gradient of
every active problem at its current trial point (the evaluations overlap on the device), followed by
one vectorised numpy update of all BFGS states.

The optimiser is a textbook BFGS on the inverse Hessian with an Armijo backtracking line search
(quadratic interpolation, safeguarded) and scipy's stopping rule (max |g| < gtol, default 1e-5, at
most 200 * P iterations).  It is NOT scipy's line search (Moré-Thuente), so iterates differ; optima
agree to the optimiser's own tolerance, which is what the parity tests compare.  It is opt-in
(`GaussianProcessEmulatorBins(fit_mode="lockstep")`); the default path stays scipy.
"""

from __future__ import annotations

from typing import Callable, NamedTuple

import numpy as np


class LockstepResult(NamedTuple):
    x: np.ndarray        # [nb, P] final points
    fun: np.ndarray      # [nb] objective values there
    jac: np.ndarray      # [nb, P]
    nit: np.ndarray      # [nb] BFGS iterations
    nfev: np.ndarray     # [nb] evaluations (value + gradient together)
    status: np.ndarray   # [nb] 0 converged, 1 iteration limit, 2 line search could not improve (precision loss)
    rounds: int          # lock-step rounds = batched evaluation calls


def minimise_lockstep(fun_and_grad: Callable, x0: np.ndarray, gtol: float = 1e-5, maxiter: int = None,
                      c1: float = 1e-4, min_step: float = 1e-14) -> LockstepResult:
    """
    Minimise nb functions at once.  `fun_and_grad(indices, points)` evaluates problems `indices`
    (int array) at `points` [len(indices), P] and returns (values, gradients); a failed evaluation is
    signalled by a non-finite value (treated as "step too long").
    """
    x = np.array(x0, dtype=np.float64)
    nb, P = x.shape
    maxiter = 200 * P if maxiter is None else maxiter
    idx_all = np.arange(nb)
    f, g = fun_and_grad(idx_all, x)
    f = np.array(f, dtype=np.float64)
    g = np.array(g, dtype=np.float64)
    if not np.all(np.isfinite(f)):
        raise ValueError("the objective is not finite at the starting point of every problem")
    H = np.tile(np.eye(P), (nb, 1, 1))
    nit = np.zeros(nb, dtype=np.int64)
    nfev = np.ones(nb, dtype=np.int64)
    status = np.full(nb, -1, dtype=np.int64)
    status[np.max(np.abs(g), axis=1) < gtol] = 0
    d = -g.copy()
    slope = np.einsum("bi,bi->b", g, d)
    # first step as in scipy's line search: alpha0 = min(1, 1 / |g|_1-ish); here 1 / max(1, |g|_2)
    alpha = 1.0 / np.maximum(1.0, np.linalg.norm(g, axis=1))
    rounds = 1
    while True:
        active = np.flatnonzero(status < 0)
        if len(active) == 0:
            break
        trial = x[active] + alpha[active, None] * d[active]
        ft, gt = fun_and_grad(active, trial)
        ft = np.array(ft, dtype=np.float64)
        gt = np.array(gt, dtype=np.float64)
        rounds += 1
        nfev[active] += 1
        a = alpha[active]
        ok = np.isfinite(ft) & (ft <= f[active] + c1 * a * slope[active])
        # ---- rejected steps: shrink (safeguarded quadratic interpolation; plain halving if not finite) ----
        rej = active[~ok]
        if len(rej):
            ar, fr = alpha[rej], ft[~ok]
            denom = 2.0 * (fr - f[rej] - slope[rej] * ar)
            with np.errstate(divide="ignore", invalid="ignore"):
                quad = -slope[rej] * ar * ar / denom
            new = np.where(np.isfinite(quad) & (denom > 0), np.clip(quad, 0.1 * ar, 0.5 * ar), 0.5 * ar)
            alpha[rej] = new
            stalled = new * np.max(np.abs(d[rej]), axis=1) < min_step * np.maximum(1.0, np.max(np.abs(x[rej]), axis=1))
            status[rej[stalled]] = 2
        # ---- accepted steps: BFGS update, new direction ----
        acc = active[ok]
        if len(acc):
            s = trial[ok] - x[acc]
            y = gt[ok] - g[acc]
            x[acc], f[acc], g[acc] = trial[ok], ft[ok], gt[ok]
            nit[acc] += 1
            ys = np.einsum("bi,bi->b", y, s)
            good = ys > 1e-10 * np.linalg.norm(y, axis=1) * np.linalg.norm(s, axis=1)
            if np.any(good):
                u = acc[good]
                rho = 1.0 / ys[good]
                sg, yg = s[good], y[good]
                Hy = np.einsum("bij,bj->bi", H[u], yg)
                yHy = np.einsum("bi,bi->b", yg, Hy)
                # H <- (I - rho s y^T) H (I - rho y s^T) + rho s s^T
                H[u] = (H[u] - rho[:, None, None] * (sg[:, :, None] * Hy[:, None, :] + Hy[:, :, None] * sg[:, None, :])
                        + ((rho * rho * yHy + rho)[:, None, None]) * (sg[:, :, None] * sg[:, None, :]))
            conv = np.max(np.abs(g[acc]), axis=1) < gtol
            status[acc[conv]] = 0
            status[acc[(~conv) & (nit[acc] >= maxiter)]] = 1
            d[acc] = -np.einsum("bij,bj->bi", H[acc], g[acc])
            slope[acc] = np.einsum("bi,bi->b", g[acc], d[acc])
            # a direction that is not a descent direction (round-off in H): restart from steepest descent
            bad = acc[slope[acc] >= 0]
            if len(bad):
                H[bad] = np.eye(P)
                d[bad] = -g[bad]
                slope[bad] = -np.einsum("bi,bi->b", g[bad], g[bad])
            alpha[acc] = 1.0
    return LockstepResult(x, f, g, nit, nfev, status, rounds)
