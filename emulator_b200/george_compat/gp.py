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

import collections
import contextlib
import ctypes
import threading
import weakref
from typing import Optional

import numpy as np

from .. import _lib
from .. import combine as _combine
from .. import dist as _dist
from ..device import DeviceProblem, evaluate_many
from . import kernels as _kernels
from .modeling import CallableModel, ConstantModel, Model

TINY = 1.25e-12  # george.gp.TINY


class _Residency:
    """Keeps the device handles of all GP objects of this process within the GPU's memory.

    A handle costs about 2 N^2 doubles (1.07 GB at the N = 8176 of a leave-one-out refit), and the
    reference's drivers keep every fitted GP: CrossCheck holds one per simulation (512 at BASELINE
    configuration C4).  A GP therefore only *needs* its handle while one of its methods runs; handles
    are created lazily and the least recently used unpinned ones are released when a new one would
    not fit.  A released GP keeps its data and parameters and gets a new handle on its next use
    (one re-factorisation)."""

    RESERVE = 1 << 30  # bytes kept free for everything else (CUDA context growth, torch, NCCL)

    def __init__(self):
        self.lock = threading.RLock()
        self.lru = collections.OrderedDict()  # id(gp) -> weakref
        self.releases = 0  # handles released to make room (diagnostics)

    @staticmethod
    def needed_bytes(n: int) -> int:
        npad = -(-n // 128) * 128
        return int(2.2 * npad * npad * 8) + (64 << 20)

    @staticmethod
    def free_bytes(device: int) -> int:
        lib = _lib.require_gpu()
        free, total = ctypes.c_ulonglong(), ctypes.c_ulonglong()
        _lib.check(lib.eb200_device_mem_info(device, ctypes.byref(free), ctypes.byref(total)), "device_mem_info")
        return int(free.value)

    def touch(self, gp):
        with self.lock:
            self.lru[id(gp)] = weakref.ref(gp)
            self.lru.move_to_end(id(gp))

    def forget(self, gp):
        with self.lock:
            self.lru.pop(id(gp), None)

    def make_room(self, device: int, need: int, keep):
        """Release least-recently-used, unpinned handles on `device` until `need` bytes fit."""
        with self.lock:
            if self.free_bytes(device) >= need + self.RESERVE:
                return
            for key in list(self.lru.keys()):
                other = self.lru[key]()
                if other is None:
                    self.lru.pop(key, None)
                    continue
                if other is keep or other.device != device or other._pins > 0 or other._problem is None:
                    continue
                other._release_locked()
                self.releases += 1
                if self.free_bytes(device) >= need + self.RESERVE:
                    return
            # nothing (more) to release: let a request that cannot possibly fit fail here, clearly, instead of
            # inside cudaMalloc (RESERVE is head-room, not part of the need)
            if self.free_bytes(device) < need:
                raise MemoryError(
                    f"GPU {device}: {need >> 20} MiB needed for a GP handle, {self.free_bytes(device) >> 20} MiB free and "
                    "every other handle is in use"
                )


_residency = _Residency()


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
        device: Optional[int] = None,
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
        self.device = _dist.resolve_device(device)
        #: evaluate the gradient together with every likelihood value (what BFGS wants);
        #: the ensemble sampler switches this off: its evaluations then go through the Cholesky
        #: factor alone (EB200_EVAL_VALUE_ONLY, a third of the flops).
        self.eager_gradient = True
        self._problem = None
        self._x = None
        self._yerr2 = None
        self._residual_key = None       # residual held by the device handle
        self._host_residual_key = None  # residual the cached values belong to
        self._r = None
        self._cache_key = None      # (p bytes) of the cached evaluation
        self._cache = None          # (lml, grad or None, info)
        self._device_p = None       # p bytes the resident factorisation belongs to
        self.computed = False
        self.n_device_evaluations = 0
        self._noise = None
        self._pins = 0              # > 0 while a method is using the device handle

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
        self._noise = np.sqrt(self._yerr2 + TINY) ** 2
        if self._x.shape[1] > _lib.MAX_DIM:
            raise ValueError(f"kernel dimension {self._x.shape[1]} exceeds the supported maximum {_lib.MAX_DIM}")
        _lib.require_gpu()  # fail here, loudly, when the CUDA library or a GPU is missing
        self.release_device()
        self._cache_key = self._cache = None
        self._host_residual_key = self._r = None
        self.computed = True
        # george factorises inside compute(); here the device handle and the factorisation are deferred
        # to the first likelihood / predict call (which every reference call site makes next), so a
        # matrix that is not positive definite raises LinAlgError there instead.

    # ---- device residency -----------------------------------------------------------------------
    @contextlib.contextmanager
    def _using_device(self):
        """Pins this GP's handle (creating it if needed) for the duration of one method."""
        with _residency.lock:
            self._pins += 1
        try:
            if self._problem is None:
                if not self.computed:
                    raise RuntimeError("You need to compute the model first")
                _residency.make_room(self.device, _Residency.needed_bytes(len(self._x)), keep=self)
                self._problem = DeviceProblem(self._x, self._noise, device=self.device)
                self._residual_key = None
                self._device_p = None
            _residency.touch(self)
            yield self._problem
        finally:
            with _residency.lock:
                self._pins -= 1

    def _release_locked(self):
        if self._problem is not None:
            self._problem.close()
            self._problem = None
        self._residual_key = None
        self._device_p = None

    def release_device(self):
        """Free the device handle (data, parameters and cached values stay); the next use re-creates it."""
        with _residency.lock:
            if self._pins > 0:
                raise RuntimeError("the device handle is in use")
            self._release_locked()

    @property
    def resident(self) -> bool:
        return self._problem is not None

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

    def _residual(self, y):
        """Host part: r = y - mean(x); a new residual invalidates cached values and the device state."""
        if not self.computed:
            raise RuntimeError("You need to compute the model first")
        r = np.ascontiguousarray(self._check_dimensions(y) - self._call_mean(self._x), dtype=np.float64)
        key = r.tobytes()
        if key != self._host_residual_key:
            self._host_residual_key = key
            self._r = r
            self._cache_key = self._cache = self._device_p = None

    def _sync_residual(self):
        """Device part (handle pinned): upload the residual if the handle holds another one."""
        if self._residual_key != self._host_residual_key:
            self._problem.set_residual(self._r)
            self._residual_key = self._host_residual_key
            self._device_p = None

    def _evaluate(self, want_grad: bool, value_only: bool = False, need_state: bool = False):
        p = np.ascontiguousarray(self.get_parameter_vector(), dtype=np.float64)
        key = p.tobytes()
        if self._cache_key == key and self._cache[2] != 0:
            # a failed factorisation fails again, every time it is asked for (as george does); nothing on the
            # device is valid at these parameters
            raise np.linalg.LinAlgError(self._failure_text(self._cache[2]))
        cached = self._cache_key == key and (self._cache[1] is not None or not want_grad)
        if cached and not (need_state and self._device_p != key):
            return self._cache
        with self._using_device():
            self._sync_residual()
            combiner = None if value_only else _combine.current()
            if combiner is not None:
                # several optimiser threads share launches (emulator_b200/combine.py); same numbers, bit for bit
                lml, grad, info = combiner.evaluate(self._problem, p, want_grad)
            else:
                lml, grad, info, _logdet, _quad = self._problem.evaluate(p, want_grad, value_only=value_only)
        self.n_device_evaluations += 1
        if info < 0:
            # the device watchdog fired (a dependency was not published within its budget): not a property of
            # the matrix -- never reported as "not positive definite", never cached
            self._cache_key = self._cache = self._device_p = None
            raise RuntimeError("device watchdog: the factorisation kernel gave up waiting for a tile (device oversubscribed?)")
        if info != 0:
            self._cache_key = key
            self._cache = (-np.inf, None, info)
            self._device_p = None  # L^-1 / alpha of an aborted factorisation must never be predicted from
            raise np.linalg.LinAlgError(self._failure_text(info))
        # a value-only evaluation (Cholesky factor alone) leaves no prediction state behind
        self._device_p = None if value_only else key
        if not (cached and not want_grad):  # keep a cached gradient when only the device state was refreshed
            self._cache_key = key
            self._cache = (lml, grad if want_grad else None, info)
        return self._cache

    @staticmethod
    def _failure_text(info):
        return f"{info}-th leading minor of the array is not positive definite"

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
        with self._using_device():
            self._sync_residual()
            # the batch factorises up to eight vectors at once, each in a matrix of its own (allocated on first use
            # and kept by the handle): make room for those too
            extra = max(0, min(len(vectors), 8) - 1)
            if extra and not getattr(self._problem, "_batch_slots_reserved", False):
                npad = -(-len(self._x) // 128) * 128
                _residency.make_room(self.device, extra * (npad + 64) * npad * 8, keep=self)
                self._problem._batch_slots_reserved = True
            lml, info, _, _ = self._problem.evaluate_value_batch(vectors)
        self.n_device_evaluations += len(vectors)
        self._device_p = None  # no prediction state survives
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
        xs = np.asarray(t)
        if xs.ndim == 1:
            xs = xs[:, None]
        if xs.ndim != 2 or xs.shape[1] != self.ndim:
            raise ValueError("Dimension mismatch")
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        mean_t = self._call_mean(xs)
        with self._using_device():  # pinned across factorisation and prediction
            # factor + alpha at the current parameters must be resident
            self._evaluate(want_grad=False, need_state=True)
            if not return_var:
                return self._problem.predict(xs, want_var=False) + mean_t
            mu, var = self._problem.predict(xs, want_var=True)
        return mu + mean_t, var

    def predict_grid(self, y, theta, independent, return_var=True, round_f32=True):
        """
        Predictions at every row of `theta` [n_theta, ndim-1] for every value of `independent`: the query
        rows (independent_b, theta_r...) that the sweeps of swiftemulator build on the host, one emulator
        call per parameter row (mocking/__init__.py:84-95), are formed on the device instead.  Returns
        mean[n_theta, n_ind] (and var).  `round_f32` rounds the rows to float32 first, as the emulators do.
        """
        self._residual(y)
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        independent = np.ascontiguousarray(independent, dtype=np.float64)
        with self._using_device():
            self._evaluate(want_grad=False, need_state=True)
            out = self._problem.predict_grid(theta, independent, want_var=return_var, round_f32=round_f32)
        mu = out[0] if return_var else out
        if not (isinstance(self.mean, ConstantModel) and self.mean.value == 0.0):
            # a non-trivial mean model is a host callable: it needs the rows
            dtype = np.float32 if round_f32 else np.float64
            rows = np.empty((theta.shape[0] * len(independent), self.ndim), dtype=dtype)
            rows[:, 0] = np.tile(independent, theta.shape[0])
            rows[:, 1:] = np.repeat(theta, len(independent), axis=0)
            mu = mu + self._call_mean(rows.astype(np.float64)).reshape(mu.shape)
        return (mu, out[1]) if return_var else mu

    def predict_mean_batch(self, y, t, vectors, siblings: int = 8):
        """
        Predictive mean at the rows of `t` for EACH hyperparameter vector in `vectors`: returns [len(vectors), len(t)].
        The loop swiftemulator's MCMC emulator runs one sample at a time to estimate the hyperparameter error
        (gaussian_process_mcmc.py:410-431: set_parameter_vector + predict per sample, each a full
        re-factorisation).  Here the samples are factorised in groups on sibling handles holding the same data
        (`eb200_gp_eval_multi_host`: shared launches for chain-bound sizes) and each sibling predicts its own.
        Does not change the GP's parameter vector.
        """
        self._residual(y)
        xs = np.asarray(t)
        if xs.ndim == 1:
            xs = xs[:, None]
        if xs.ndim != 2 or xs.shape[1] != self.ndim:
            raise ValueError("Dimension mismatch")
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        vectors = np.atleast_2d(np.asarray(vectors, dtype=np.float64))
        if vectors.shape[1] != len(self):
            raise ValueError("Dimension mismatch")
        mean_t = self._call_mean(xs)
        out = np.empty((len(vectors), len(xs)))
        count = max(1, min(int(siblings), len(vectors)))
        with self._using_device():
            self._sync_residual()
            _residency.make_room(self.device, (count - 1) * _Residency.needed_bytes(len(self._x)), keep=self)
            extra = []
            try:
                for _ in range(count - 1):
                    q = DeviceProblem(self._x, self._noise, device=self.device)
                    q.set_residual(self._r)
                    extra.append(q)
                group = [self._problem] + extra
                for g0 in range(0, len(vectors), count):
                    part = vectors[g0 : g0 + count]
                    _, _, info = evaluate_many(group[: len(part)], part, want_grad=False)
                    self.n_device_evaluations += len(part)
                    if np.any(info != 0):
                        raise np.linalg.LinAlgError(self._failure_text(int(info[np.nonzero(info)[0][0]])))
                    for j in range(len(part)):
                        out[g0 + j] = group[j].predict(xs, want_var=False) + mean_t
            finally:
                for q in extra:
                    q.close()
                self._device_p = None  # the handle's factorisation now belongs to one of `vectors`
        return out

    def predict_grid_device(self, y, theta_t, independent_t, mean_out, var_out=None, round_f32=True):
        """
        `predict_grid` with everything in device memory: `theta_t` [n_theta, ndim-1], `independent_t` [n_ind] and
        the outputs `mean_out` / `var_out` [n_theta, n_ind] are CUDA float64 tensors on this GP's device; the call
        returns when the results are complete in device memory and nothing is copied to the host (sharded sweeps all-gather the
        results from where they are, emulator_b200/dist.py).  Zero-mean GPs only (a mean model is a host callable).
        """
        import torch

        if not (isinstance(self.mean, ConstantModel) and self.mean.value == 0.0):
            raise NotImplementedError("device-resident grid prediction needs a zero mean (mean models are host callables)")
        self._residual(y)
        n_theta, n_ind = int(theta_t.shape[0]), int(independent_t.shape[0])
        for t in (theta_t, independent_t, mean_out) + ((var_out,) if var_out is not None else ()):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.device.index == self.device):
                raise ValueError("device-resident grid prediction needs contiguous float64 CUDA tensors on the GP's device")
        if theta_t.shape[1] != self.ndim - 1 or mean_out.numel() != n_theta * n_ind or (var_out is not None and var_out.numel() != n_theta * n_ind):
            raise ValueError("Dimension mismatch")
        with self._using_device():
            self._evaluate(want_grad=False, need_state=True)
            # inputs written on torch's current stream must be complete; the work then runs on the handle's own stream
            # (a stream argument of 0 means exactly that in the C ABI -- torch's default stream also has handle 0, so
            # it cannot be named there) and is complete when this returns: the handle may be released once unpinned
            torch.cuda.current_stream(theta_t.device).synchronize()
            self._problem.predict_grid_device(
                n_theta, theta_t.data_ptr(), n_ind, independent_t.data_ptr(), var_out is not None, mean_out.data_ptr(),
                var_out.data_ptr() if var_out is not None else 0, stream=0, round_f32=round_f32,
            )
            self._problem.synchronize()

    # ---- lifetime ---------------------------------------------------------------------------------
    def close(self):
        self.release_device()
        _residency.forget(self)
        self.computed = False

    def __del__(self):  # pragma: no cover
        try:
            _residency.forget(self)
        except Exception:
            pass

    def __deepcopy__(self, memo):
        raise TypeError("a device-resident GP cannot be deep-copied; copy the kernel and build a new GP")


@contextlib.contextmanager
def pinned_evaluator(gps, ys):
    """
    Context for optimisers that advance many GPs together (emulator_b200/optimise.py): pins the device
    handles of `gps`, uploads their residuals (y - mean) once, and yields
    ``evaluate(indices, vectors, want_grad=True) -> (lml, grad, info)``, which evaluates GP ``indices[k]``
    at ``vectors[k]`` in ONE native call (the evaluations overlap on the device).  The GPs' own parameter
    vectors are not touched; their cached values are dropped.
    """
    with contextlib.ExitStack() as stack:
        for gp, y in zip(gps, ys):
            gp._residual(y)
            stack.enter_context(gp._using_device())
            gp._sync_residual()
        problems = [gp._problem for gp in gps]

        def evaluate(indices, vectors, want_grad=True):
            lml, grad, info = evaluate_many([problems[i] for i in indices], vectors, want_grad=want_grad)
            for i in indices:
                g = gps[i]
                g.n_device_evaluations += 1
                g._device_p = g._cache_key = g._cache = None
            return lml, grad, info

        yield evaluate
