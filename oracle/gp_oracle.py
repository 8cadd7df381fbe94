"""
CPU oracle for the Gaussian-process hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm
may import this module.  The product path (``emulator_b200``) never does; it fails loudly
when the CUDA library is missing.

PARITY UNPINNED.  The arithmetic of this path lives in the third-party package ``george``
(dfm/george, 0.4.x line -- the reference pins no version: /root/reference/requirements.txt:4,
setup.py:34), which is absent from /root/reference and from this image, and the reference's
own tests hold no golden vectors, known-answer tests or fixtures for it
(/root/reference/tests/test_emulator_gpe.py:60-66 and siblings are assertion-free smoke
tests on unseeded data).  This file therefore restates george's published algorithm
(``george/gp.py``: GP.compute / log_likelihood / grad_log_likelihood / predict;
``george/solvers/basic.py``: BasicSolver; ``george/kernels.py`` + ``include/george/kernels.h``,
``metrics.h``: ConstantKernel, ExpSquaredKernel with an axis-aligned metric) and is anchored
on the reference's call sites:

  * kernel construction   swiftemulator/emulators/gaussian_process.py:181-183
  * GP(...) / compute     swiftemulator/emulators/gaussian_process.py:191-206
  * nll / grad closures   swiftemulator/emulators/gaussian_process.py:208-214
  * predict               swiftemulator/emulators/gaussian_process.py:285-287, 344-346

Self-checks that stand in for the missing known-answer tests live in
``tests/test_oracle.py`` (finite-difference gradient, slogdet / solve cross-checks, symmetry,
variance >= 0, row-permutation invariance) and in the committed ``tests/golden`` vectors.

What narrows "unpinned": ``tests/test_oracle_vs_sklearn.py`` checks this file against an independently written
implementation of the same quantities, scikit-learn's ``GaussianProcessRegressor`` with
``ConstantKernel(d e^p0) * RBF(length_scale = sqrt(M))``, ``alpha`` = george's diagonal term, ``optimizer=None``:
log marginal likelihood, its gradient (chain rule d/dlog M = 1/2 d/dlog l), predictive mean and variance agree to
1e-13 .. 2e-11 on seven shapes (d = 1 .. 16, N up to 1024), asserted at 1e-10.  It remains unpinned against george
ITSELF (george's own conventions -- amp = d e^p0, the sqrt(yerr^2 + TINY)^2 diagonal, no noise in the predictive
variance -- are restated from its published source, not executed).

Everything is IEEE fp64 through numpy / scipy.linalg (LAPACK dpotrf / dpotrs), exactly the
libraries george's BasicSolver calls.
"""

from __future__ import annotations

import numpy as np
from scipy.linalg import cho_solve, cholesky

# george.gp.TINY: default white-noise variance, exp(log(TINY)); frozen in the reference's use.
TINY = 1.25e-12


def default_parameter_vector(ndim: int) -> np.ndarray:
    """
    Parameter vector of ``1**2 * ExpSquaredKernel(np.ones(ndim), ndim=ndim)``
    (gaussian_process.py:181-183): george turns the scalar into
    ``ConstantKernel(log_constant=log(1.0 / ndim))`` and the metric into ``log_M_i_i = 0``.
    """
    return np.concatenate([[np.log(1.0 / ndim)], np.zeros(ndim)])


def parameter_names(ndim: int):
    """george's names for the unfrozen parameters of GP(constant * ExpSquared)."""
    return tuple(
        ["kernel:k1:log_constant"] + [f"kernel:k2:metric:log_M_{i}_{i}" for i in range(ndim)]
    )


def amplitude(p: np.ndarray, ndim: int) -> float:
    """ConstantKernel is evaluated per axis and summed over its ndim axes (kernels.h)."""
    return ndim * np.exp(p[0])


def squared_distance(p: np.ndarray, x1: np.ndarray, x2: np.ndarray) -> np.ndarray:
    """
    AxisAlignedMetric::value (metrics.h): r2 = sum_i d_i * d_i / M_i, accumulated in axis
    order, M_i = exp(log_M_i_i).
    """
    metric = np.exp(p[1:])
    r2 = np.zeros((x1.shape[0], x2.shape[0]))
    for i in range(x1.shape[1]):
        diff = x1[:, i][:, None] - x2[:, i][None, :]
        r2 += diff * diff / metric[i]
    return r2


def kernel_value(p: np.ndarray, x1: np.ndarray, x2: np.ndarray | None = None) -> np.ndarray:
    """Kernel.get_value(x1, x2): constant * exp(-0.5 r2)."""
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    x2 = x1 if x2 is None else np.ascontiguousarray(x2, dtype=np.float64)
    return amplitude(p, x1.shape[1]) * np.exp(-0.5 * squared_distance(p, x1, x2))


def kernel_gradient(p: np.ndarray, x: np.ndarray) -> np.ndarray:
    """
    Kernel.get_gradient(x): the materialised N x N x P tensor george builds.
    d/d log c = k;  d/d log M_i = k * 0.5 * d_i^2 / M_i.
    """
    x = np.ascontiguousarray(x, dtype=np.float64)
    n, ndim = x.shape
    k = kernel_value(p, x)
    metric = np.exp(p[1:])
    grad = np.empty((n, n, ndim + 1))
    grad[:, :, 0] = k
    for i in range(ndim):
        diff = x[:, i][:, None] - x[:, i][None, :]
        grad[:, :, i + 1] = k * (0.5 * diff * diff / metric[i])
    return grad


def noise_diagonal(yerr) -> np.ndarray:
    """
    The diagonal term as george forms it.  GP.compute squares ``yerr`` in its input dtype
    (float32 for GaussianProcessEmulator, gaussian_process.py:93-95), stores it as float64,
    adds the white noise, takes the square root, and BasicSolver.compute squares it again.
    """
    yerr2 = np.ascontiguousarray(np.atleast_1d(yerr) ** 2, dtype=np.float64)
    return np.sqrt(yerr2 + TINY) ** 2


class OracleGP:
    """
    ``george.GP``-shaped object restricted to what the reference calls.  Zero mean unless a
    callable ``mean`` (CallableModel.get_value) is given; the white-noise term is frozen at
    log(TINY); the parameter vector is the kernel's.
    """

    def __init__(self, ndim: int, parameter_vector=None, mean=None):
        self.ndim = int(ndim)
        self._p = (
            default_parameter_vector(ndim)
            if parameter_vector is None
            else np.array(parameter_vector, dtype=np.float64)
        )
        self._mean = mean
        self._x = None
        self._diag = None
        self._factor = None
        self._dirty = True
        self.n_factorisations = 0

    # ---- modeling protocol -------------------------------------------------------------
    def __len__(self):
        return self.ndim + 1

    def get_parameter_vector(self):
        return self._p.copy()

    def set_parameter_vector(self, p):
        self._p = np.array(p, dtype=np.float64)
        self._dirty = True

    def get_parameter_names(self):
        return parameter_names(self.ndim)

    # ---- GP.compute / recompute ----------------------------------------------------------
    def compute(self, x, yerr=0.0):
        self._x = np.ascontiguousarray(x, dtype=np.float64)
        if self._x.ndim == 1:
            self._x = self._x[:, None]
        try:
            yerr = float(yerr) * np.ones(len(self._x))
        except TypeError:
            pass
        self._diag = noise_diagonal(yerr)
        self._factorise()

    def _factorise(self):
        k = kernel_value(self._p, self._x)
        k[np.diag_indices_from(k)] += self._diag
        # BasicSolver.compute: cholesky(K, overwrite_a=True, lower=False)
        self._factor = (cholesky(k, overwrite_a=True, lower=False), False)
        self.log_determinant = 2.0 * np.sum(np.log(np.diag(self._factor[0])))
        self._const = -0.5 * (len(self._x) * np.log(2.0 * np.pi) + self.log_determinant)
        self._dirty = False
        self.n_factorisations += 1

    def recompute(self):
        if self._dirty or self._factor is None:
            self._factorise()

    def _call_mean(self, x):
        if self._mean is None:
            return np.zeros(len(x))
        return np.asarray(self._mean(x), dtype=np.float64)

    def _residual(self, y):
        return np.ascontiguousarray(np.asarray(y) - self._call_mean(self._x), dtype=np.float64)

    # ---- likelihood ------------------------------------------------------------------------
    def log_likelihood(self, y):
        """GP.log_likelihood: const - 0.5 r^T K^-1 r; -inf when not finite; LinAlgError propagates."""
        self.recompute()
        r = self._residual(y)
        ll = self._const - 0.5 * np.dot(r, cho_solve(self._factor, r))
        return ll if np.isfinite(ll) else -np.inf

    def grad_log_likelihood(self, y):
        """GP.grad_log_likelihood: 0.5 einsum('ijk,ij', dK/dp, alpha alpha^T - K^-1)."""
        self.recompute()
        r = self._residual(y)
        alpha = cho_solve(self._factor, r)
        k_inv = cho_solve(self._factor, np.eye(len(self._x)))
        a = np.einsum("i,j", alpha, alpha) - k_inv
        kg = kernel_gradient(self._p, self._x)
        return 0.5 * np.einsum("ijk,ij", kg, a)

    # ---- prediction --------------------------------------------------------------------------
    def predict(self, y, t, return_cov=True, return_var=False):
        """GP.predict restricted to return_cov=False (all reference call sites)."""
        if return_cov and not return_var:
            raise NotImplementedError("the reference never requests the full covariance")
        self.recompute()
        alpha = cho_solve(self._factor, self._residual(y))
        xs = np.ascontiguousarray(t, dtype=np.float64)
        if xs.ndim == 1:
            xs = xs[:, None]
        kxs = kernel_value(self._p, xs, self._x)
        mu = np.dot(kxs, alpha) + self._call_mean(xs)
        if not return_var:
            return mu
        kinv_kxs = cho_solve(self._factor, kxs.T)
        var = np.full(len(xs), amplitude(self._p, self.ndim))  # kernel.get_value(xs, diag=True)
        var -= np.sum(kxs.T * kinv_kxs, axis=0)
        return mu, var


# ---- function-style helpers used by the parity tests -------------------------------------------

def lml_and_grad(p, x, diag, r):
    """
    LML and gradient at p for design x [N, d], diagonal term diag [N] (= noise_diagonal(yerr))
    and residual r [N]; george-faithful algorithm (full inverse, materialised dK tensor).
    Also returns logdet and the quadratic form.
    """
    x = np.ascontiguousarray(x, dtype=np.float64)
    p = np.asarray(p, dtype=np.float64)
    n = len(x)
    k = kernel_value(p, x)
    k[np.diag_indices_from(k)] += diag
    factor = (cholesky(k, overwrite_a=True, lower=False), False)
    logdet = 2.0 * np.sum(np.log(np.diag(factor[0])))
    alpha = cho_solve(factor, r)
    quad = np.dot(r, alpha)
    lml = -0.5 * (n * np.log(2.0 * np.pi) + logdet) - 0.5 * quad
    k_inv = cho_solve(factor, np.eye(n))
    a = np.einsum("i,j", alpha, alpha) - k_inv
    kg = kernel_gradient(p, x)
    grad = 0.5 * np.einsum("ijk,ij", kg, a)
    return lml, grad, logdet, quad


def lml_and_grad_lean(p, x, diag, r):
    """
    Same quantities with the work the GPU path does: one factorisation, dpotri-style inverse,
    traces accumulated per parameter without the N x N x P tensor.  This is the "optimised
    CPU" comparator of BASELINE.md section 3, not the george-faithful baseline.
    """
    from scipy.linalg import lapack

    x = np.ascontiguousarray(x, dtype=np.float64)
    p = np.asarray(p, dtype=np.float64)
    n, ndim = x.shape
    k0 = kernel_value(p, x)
    k = k0.copy()
    k[np.diag_indices_from(k)] += diag
    c, info = lapack.dpotrf(k, lower=1, overwrite_a=1)
    if info != 0:
        raise np.linalg.LinAlgError(f"{info}-th leading minor not positive definite")
    logdet = 2.0 * np.sum(np.log(np.diag(c)))
    alpha = cho_solve((c, True), r)
    quad = np.dot(r, alpha)
    lml = -0.5 * (n * np.log(2.0 * np.pi) + logdet) - 0.5 * quad
    kinv, info = lapack.dpotri(c, lower=1, overwrite_c=1)
    kinv = np.tril(kinv) + np.tril(kinv, -1).T
    a = np.outer(alpha, alpha) - kinv
    ak = a * k0
    metric = np.exp(p[1:])
    grad = np.empty(ndim + 1)
    grad[0] = 0.5 * ak.sum()
    for i in range(ndim):
        diff = x[:, i][:, None] - x[:, i][None, :]
        grad[i + 1] = 0.5 * np.sum(ak * (0.5 * diff * diff / metric[i]))
    return lml, grad, logdet, quad


def predict(p, x, diag, r, t, return_var=True):
    """Predictive mean (no mean model) and variance at t [M, d], george-faithful."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.float64)
    k = kernel_value(p, x)
    k[np.diag_indices_from(k)] += diag
    factor = (cholesky(k, overwrite_a=True, lower=False), False)
    alpha = cho_solve(factor, r)
    kxs = kernel_value(p, t, x)
    mu = np.dot(kxs, alpha)
    if not return_var:
        return mu
    kinv_kxs = cho_solve(factor, kxs.T)
    var = np.full(len(t), amplitude(p, x.shape[1])) - np.sum(kxs.T * kinv_kxs, axis=0)
    return mu, var
