"""
Extended-precision (numpy longdouble, 64-bit mantissa on x86) restatement of oracle/gp_oracle.py's
likelihood and gradient -- TEST INFRASTRUCTURE ONLY, small N only (pure numpy loops, O(N^3)).

Purpose: the fp64 oracle and the device path agree to 1e-13 ... 1e-11 on well-conditioned problems, but
on ill-conditioned ones (cond(K) ~ 1e8 ... 1e9: few kernel dimensions, dense designs) BOTH carry errors
of order cond(K) * 2^-53 and differ from each other by that much.  This module provides the
better-conditioned yardstick that tells an implementation difference from the conditioning floor:
same formulas (george BasicSolver semantics, see gp_oracle.py), every step in longdouble.
"""

from __future__ import annotations

import numpy as np

LD = np.longdouble


def _cholesky_lower(k: np.ndarray) -> np.ndarray:
    n = k.shape[0]
    lo = np.zeros((n, n), dtype=LD)
    for j in range(n):
        s = k[j, j] - np.dot(lo[j, :j], lo[j, :j])
        if not s > 0:
            raise np.linalg.LinAlgError(f"{j + 1}-th leading minor of the array is not positive definite")
        lo[j, j] = np.sqrt(s)
        if j + 1 < n:
            lo[j + 1 :, j] = (k[j + 1 :, j] - lo[j + 1 :, :j] @ lo[j, :j]) / lo[j, j]
    return lo


def _inverse_lower(lo: np.ndarray) -> np.ndarray:
    n = lo.shape[0]
    xi = np.zeros((n, n), dtype=LD)
    for i in range(n):
        xi[i, i] = LD(1) / lo[i, i]
        if i:
            xi[i, :i] = -(lo[i, :i] @ xi[:i, :i]) * xi[i, i]
    return xi


def lml_and_grad(p, x, diag, r):
    """(lml, grad[P], logdet, quad) in longdouble; arguments as gp_oracle.lml_and_grad."""
    p = np.asarray(p, dtype=LD)
    x = np.asarray(x, dtype=LD)
    diag = np.asarray(diag, dtype=LD)
    r = np.asarray(r, dtype=LD)
    n, d = x.shape
    amp = LD(d) * np.exp(p[0])
    inv_m = LD(1) / np.exp(p[1:])
    sq = np.zeros((n, n), dtype=LD)
    per_axis = []
    for m in range(d):  # accumulate in axis order, as george's metric does
        dm = (x[:, m, None] - x[None, :, m]) ** 2 * inv_m[m]
        per_axis.append(dm)
        sq = sq + dm
    k0 = amp * np.exp(LD(-0.5) * sq)
    k = k0.copy()
    k[np.diag_indices(n)] += diag
    lo = _cholesky_lower(k)
    xi = _inverse_lower(lo)
    z = xi @ r
    alpha = xi.T @ z
    logdet = LD(2) * np.sum(np.log(np.diag(lo)))
    quad = np.dot(z, z)
    lml = LD(-0.5) * (LD(n) * np.log(LD(2) * LD(np.pi)) + logdet) - LD(0.5) * quad
    a = np.outer(alpha, alpha) - xi.T @ xi
    grad = np.zeros(d + 1, dtype=LD)
    grad[0] = LD(0.5) * np.sum(a * k0)
    for m in range(d):
        grad[m + 1] = LD(0.5) * np.sum(a * k0 * (LD(0.5) * per_axis[m]))
    return lml, grad, logdet, quad
