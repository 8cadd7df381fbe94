"""
``george.GP``-shaped object backed by the sm_100a CUDA library: the drop-in boundary of the
hot path.  It exposes exactly what swiftemulator's emulators call:

    compute(x, yerr), set_parameter_vector(p), get_parameter_vector(), __len__,
    get_parameter_names(), log_likelihood(y), grad_log_likelihood(y),
    predict(y, t, return_cov=False, return_var=bool)

(/root/reference/swiftemulator/emulators/gaussian_process.py:191-226, :285-287, :344-346 and
the same call sequence in gaussian_process_bins.py, gaussian_process_mcmc.py,
gaussian_process_one_dim.py).

Semantics follow george 0.4 (restated in oracle/gp_oracle.py): zero or callable mean (frozen),
white noise frozen at log(1.25e-12), BasicSolver maths, `set_parameter_vector` only marks the
object dirty, LinAlgError on a failed Cholesky, -inf for a non-finite likelihood.  One device
pass produces value and gradient; results are cached per parameter vector because
scipy.optimize calls fun(p) and jac(p) separately (gaussian_process.py:208-221).
"""

from __future__ import annotations

import numpy as np

from ..device import DeviceProblem
from . import kernels as _kernels
from .modeling import CallableModel, ConstantModel, Model

TINY = 1.25e-12  # george.gp.TINY


class GP:
    def __init__(
        self,
        kernel=None,
        fit_kernel=True,
        mean=None,
        fit_mean=None,
        white_noise=None,
        fit_white_noise=None,
        solver=None,
        device: int = 0,
        **kwargs,
    ):
        if kernel is None:
            raise NotImplementedError("a kernel is required on the accelerated path")
        if solver is not None:
            raise NotImplementedError("only the basic (dense Cholesky) solver exists on the device")
        if fit_mean:
            raise NotImplementedError("fit_mean=True is not used by swiftemulator and has no device path")
        if fit_white_noise or (white_noise is not None and white_noise != np.log(TINY)):
            raise NotImplementedError("the white-noise term is frozen at george's default log(1.25e-12)")
        if not fit_kernel:
            raise NotImplementedError("fit_kernel=False is not used by swiftemulator")
        self.ndim = _kernels.device_form(kernel)
        self.kernel = kernel
        if mean is None:
            self.mean = ConstantModel(0.0)
        elif isinstance(mean, Model):
            self.mean = mean
        elif callable(mean):
            self.mean = CallableModel(mean)
        else:
            self.mean = ConstantModel(float(mean))
        self.device = device
        #: evaluate the gradient together with every likelihood value (what BFGS wants);
        #: the ensemble sampler switches this off: its evaluations then go through the Cholesky
        #: factor alone (EB200_EVAL_VALUE_ONLY, a third of the flops).
        self.eager_gradient = True
        self._problem = None
        self._x = None
        self._yerr2 = None
        self._residual_key = None
        self._cache_key = None      # (p bytes) of the cached evaluation
        self._cache = None          # (lml, grad or None, info)
        self._device_p = None       # p bytes the resident factorisation belongs to
        self.computed = False
        self.n_device_evaluations = 0

    # ---- modeling protocol ----------------------------------------------------------------
    def __len__(self):
        return len(self.kernel)

    def get_parameter_vector(self):
        return self.kernel.get_parameter_vector()

    def set_parameter_vector(self, vector):
        self.kernel.set_parameter_vector(np.asarray(vector, dtype=np.float64))

    def get_parameter_names(self):
        return tuple(f"kernel:{n}" for n in self.kernel.get_parameter_names())

    # ---- compute ------------------------------------------------------------------------------
    def compute(self, x, yerr=0.0, **kwargs):
        x = np.asarray(x)
        if x.ndim == 1:
            x = x[:, None]
        if x.ndim != 2 or x.shape[1] != self.ndim:
            raise ValueError("Dimension mismatch")
        self._x = np.ascontiguousarray(x, dtype=np.float64)
        try:
            yerr2 = float(yerr) ** 2 * np.ones(len(x))
        except TypeError:
            yerr = np.atleast_1d(yerr)
            if yerr.ndim > 1:
                raise ValueError("The predicted dimension must be 1-D")
            if len(yerr) != len(x):
                raise ValueError("Dimension mismatch")
            yerr2 = yerr ** 2  # in the caller's dtype, as george does
        self._yerr2 = np.ascontiguousarray(yerr2, dtype=np.float64)
        # GP.compute -> sqrt(yerr2 + white noise); BasicSolver.compute squares it again
        noise = np.sqrt(self._yerr2 + TINY) ** 2
        if self._problem is not None:
            self._problem.close()
        self._problem = DeviceProblem(self._x, noise, device=self.device)
        self._residual_key = None
        self._cache_key = self._cache = self._device_p = None
        self.computed = True
        # george factorises inside compute(); here the factorisation is deferred to the first
        # likelihood / predict call (which every reference call site makes next), so a matrix
        # that is not positive definite raises LinAlgError there instead.

    # ---- internals ----------------------------------------------------------------------------
    def _call_mean(self, x):
        return np.asarray(self.mean.get_value(x), dtype=np.float64)

    def _check_dimensions(self, y):
        y = np.atleast_1d(y)
        if y.ndim > 1:
            raise ValueError("The predicted dimension must be 1-D")
        if len(y) != len(self._x):
            raise ValueError("Dimension mismatch")
        return y

    def _set_residual(self, r):
        key = r.tobytes()
        if key != self._residual_key:
            self._problem.set_residual(r)
            self._residual_key = key
            self._cache_key = self._cache = self._device_p = None

    def _residual(self, y):
        if not self.computed:
            raise RuntimeError("You need to compute the model first")
        r = np.ascontiguousarray(self._check_dimensions(y) - self._call_mean(self._x), dtype=np.float64)
        self._set_residual(r)

    def _evaluate(self, want_grad: bool, value_only: bool = False):
        p = np.ascontiguousarray(self.get_parameter_vector(), dtype=np.float64)
        key = p.tobytes()
        if self._cache_key == key and (self._cache[1] is not None or not want_grad):
            return self._cache
        lml, grad, info, _logdet, _quad = self._problem.evaluate(p, want_grad, value_only=value_only)
        self.n_device_evaluations += 1
        # a value-only evaluation (Cholesky factor alone) leaves no prediction state behind
        self._device_p = None if value_only else key
        self._cache_key = key
        self._cache = (lml, grad if want_grad else None, info)
        if info != 0:
            raise np.linalg.LinAlgError(f"{info}-th leading minor of the array is not positive definite")
        return self._cache

    # ---- likelihood ------------------------------------------------------------------------------
    def log_likelihood(self, y, quiet=False):
        self._residual(y)
        try:
            lml, _, _ = self._evaluate(want_grad=self.eager_gradient, value_only=not self.eager_gradient)
        except np.linalg.LinAlgError:
            if quiet:
                return -np.inf
            raise
        return lml if np.isfinite(lml) else -np.inf

    def log_likelihood_batch(self, y, vectors, quiet=True):
        """log_likelihood at each row of `vectors` (value only), evaluated together on the device: the
        call an ensemble sampler makes per move (emcee's ``vectorize=True`` protocol).  Does not change
        the GP's own parameter vector.  Failed factorisations / non-finite values give -inf (quiet)."""
        self._residual(y)
        vectors = np.atleast_2d(np.asarray(vectors, dtype=np.float64))
        lml, info, _, _ = self._problem.evaluate_value_batch(vectors)
        self.n_device_evaluations += len(vectors)
        self._device_p = None  # no prediction state survives
        self._cache_key = self._cache = None
        if np.any(info != 0) and not quiet:
            raise np.linalg.LinAlgError("a leading minor of the array is not positive definite")
        return np.where((info == 0) & np.isfinite(lml), lml, -np.inf)

    def grad_log_likelihood(self, y, quiet=False):
        self._residual(y)
        try:
            _, grad, _ = self._evaluate(want_grad=True)
        except np.linalg.LinAlgError:
            if quiet:
                return np.zeros(len(self), dtype=np.float64)
            raise
        return grad.copy()

    def nll(self, vector, y, quiet=True):
        self.set_parameter_vector(vector)
        return -self.log_likelihood(y, quiet=quiet)

    def grad_nll(self, vector, y, quiet=True):
        self.set_parameter_vector(vector)
        return -self.grad_log_likelihood(y, quiet=quiet)

    # ---- prediction -----------------------------------------------------------------------------------
    def predict(self, y, t, return_cov=True, return_var=False, cache=True, kernel=None):
        if kernel is not None:
            raise NotImplementedError("predicting with a different kernel has no device path")
        if return_cov and not return_var:
            raise NotImplementedError(
                "the full predictive covariance is never requested by swiftemulator and has no device path"
            )
        self._residual(y)
        p = np.ascontiguousarray(self.get_parameter_vector(), dtype=np.float64)
        if self._device_p != p.tobytes():
            # factor + alpha at the current parameters must be resident
            self._cache_key = None
            self._evaluate(want_grad=False)
        xs = np.asarray(t)
        if xs.ndim == 1:
            xs = xs[:, None]
        if xs.ndim != 2 or xs.shape[1] != self.ndim:
            raise ValueError("Dimension mismatch")
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        mean_t = self._call_mean(xs)
        if not return_var:
            return self._problem.predict(xs, want_var=False) + mean_t
        mu, var = self._problem.predict(xs, want_var=True)
        return mu + mean_t, var

    # ---- access for batched drivers ---------------------------------------------------------
    @property
    def problem(self) -> DeviceProblem:
        return self._problem

    def close(self):
        if self._problem is not None:
            self._problem.close()
            self._problem = None
            self.computed = False

    def __deepcopy__(self, memo):
        raise TypeError("a device-resident GP cannot be deep-copied; copy the kernel and build a new GP")
