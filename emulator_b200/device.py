"""
Thin object wrapper around one native GP handle (``eb200_gp``): owns the device state of
one Gaussian process -- design matrix, K/L/L^-1 matrix, alpha -- and exposes the C ABI calls.
"""

from __future__ import annotations

import ctypes
from ctypes import c_void_p

import numpy as np

from . import _lib


# evaluation modes of eb200_gp_eval_* (include/emulator_b200.h)
EVAL_LML, EVAL_GRAD, EVAL_VALUE_ONLY = 0, 1, 2


def _mode(want_grad, value_only) -> int:
    if value_only:
        if want_grad:
            raise ValueError("a value-only evaluation has no gradient")
        return EVAL_VALUE_ONLY
    return EVAL_GRAD if want_grad else EVAL_LML


class DeviceProblem:
    """One GP problem resident on one GPU."""

    def __init__(self, x: np.ndarray, noise: np.ndarray, device: int = 0):
        lib = _lib.require_gpu()
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim != 2:
            raise ValueError("x must be [n, d]")
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        if noise.shape != (x.shape[0],):
            raise ValueError("noise must be [n]")
        self.n, self.d = x.shape
        if self.d > _lib.MAX_DIM:
            raise ValueError(f"kernel dimension {self.d} exceeds the supported maximum {_lib.MAX_DIM}")
        self._lib = lib
        self._h = c_void_p()
        _lib.check(
            lib.eb200_gp_create(ctypes.byref(self._h), device, self.n, self.d, _lib.as_double_ptr(x), _lib.as_double_ptr(noise)),
            "eb200_gp_create",
        )
        self.device = device
        self._result = np.zeros(_lib.result_len(self.d))

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.eb200_gp_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def padded_n(self) -> int:
        return self._lib.eb200_gp_padded_n(self._h)

    # -- hot path ---------------------------------------------------------------------------
    def set_residual(self, r: np.ndarray):
        r = np.ascontiguousarray(r, dtype=np.float64)
        if r.shape != (self.n,):
            raise ValueError("residual must be [n]")
        _lib.check(self._lib.eb200_gp_set_residual_host(self._h, _lib.as_double_ptr(r)), "set_residual")

    def evaluate(self, p: np.ndarray, want_grad: bool = True, value_only: bool = False):
        """Returns (lml, grad[P], info, logdet, quad) at hyperparameters p.

        ``value_only`` evaluates the likelihood through the Cholesky factor alone (a third of the
        flops); it leaves no prediction state (L^-1, alpha) on the device."""
        p = np.ascontiguousarray(p, dtype=np.float64)
        if p.shape != (self.d + 1,):
            raise ValueError(f"parameter vector must have {self.d + 1} entries")
        out = self._result
        _lib.check(
            self._lib.eb200_gp_eval_host(self._h, _lib.as_double_ptr(p), _mode(want_grad, value_only), _lib.as_double_ptr(out)),
            "eb200_gp_eval_host",
        )
        d = self.d
        return float(out[0]), out[1 : d + 2].copy(), int(out[d + 2]), float(out[d + 3]), float(out[d + 4])

    def evaluate_value_batch(self, ps: np.ndarray):
        """Value-only evaluations at the rows of ps[nb, P]; returns (lml[nb], info[nb], logdet[nb], quad[nb]).

        The factorisations share launches (up to eight at a time): a value-only evaluation is bound by
        its dependency chain, not by throughput.  Leaves no prediction state."""
        ps = np.ascontiguousarray(ps, dtype=np.float64)
        if ps.ndim != 2 or ps.shape[1] != self.d + 1:
            raise ValueError(f"parameter vectors must be [nb, {self.d + 1}]")
        nb, length = ps.shape[0], _lib.result_len(self.d)
        out = np.zeros((nb, length))
        if nb:
            _lib.check(
                self._lib.eb200_gp_eval_value_batch_host(self._h, nb, _lib.as_double_ptr(ps), _lib.as_double_ptr(out)),
                "eb200_gp_eval_value_batch_host",
            )
        d = self.d
        return out[:, 0].copy(), out[:, d + 2].astype(np.int64), out[:, d + 3].copy(), out[:, d + 4].copy()

    def predict(self, t: np.ndarray, want_var: bool = True):
        t = np.ascontiguousarray(t, dtype=np.float64)
        if t.ndim != 2 or t.shape[1] != self.d:
            raise ValueError(f"query points must be [m, {self.d}]")
        m = t.shape[0]
        mu = np.empty(m)
        var = np.empty(m) if want_var else None
        _lib.check(
            self._lib.eb200_gp_predict_host(
                self._h, m, _lib.as_double_ptr(t), int(bool(want_var)), _lib.as_double_ptr(mu),
                _lib.as_double_ptr(var) if want_var else None,
            ),
            "eb200_gp_predict_host",
        )
        return (mu, var) if want_var else mu

    def predict_grid(self, theta: np.ndarray, independent: np.ndarray, want_var: bool = True, round_f32: bool = True):
        """Predict every row of theta[n_theta, d-1] at every independent value; returns arrays
        [n_theta, n_ind].  The query rows (independent, theta...) are formed on the device."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        independent = np.ascontiguousarray(independent, dtype=np.float64)
        if theta.ndim != 2 or theta.shape[1] != self.d - 1 or independent.ndim != 1 or len(independent) < 1:
            raise ValueError(f"theta must be [n_theta, {self.d - 1}] and independent a non-empty vector")
        n_theta, n_ind = theta.shape[0], len(independent)
        mu = np.empty((n_theta, n_ind))
        var = np.empty((n_theta, n_ind)) if want_var else None
        _lib.check(
            self._lib.eb200_gp_predict_grid_host(
                self._h, n_theta, _lib.as_double_ptr(theta), n_ind, _lib.as_double_ptr(independent), int(bool(round_f32)),
                int(bool(want_var)), _lib.as_double_ptr(mu), _lib.as_double_ptr(var) if want_var else None,
            ),
            "eb200_gp_predict_grid_host",
        )
        return (mu, var) if want_var else mu

    # -- device-pointer variants (inputs already in HBM; asynchronous) ----------------------------
    # `stream` is a cudaStream_t handle as an integer.  0 selects the HANDLE'S OWN stream (not the legacy default
    # stream, whose handle is also 0 -- torch.cuda.default_stream().cuda_stream == 0): pass a stream created with
    # torch.cuda.Stream(), or 1 (cudaStreamLegacy) to name the legacy default stream explicitly.
    def evaluate_device(self, p_ptr: int, want_grad: bool, result_ptr: int, stream: int = 0, value_only: bool = False):
        _lib.check(
            self._lib.eb200_gp_eval_device(self._h, c_void_p(p_ptr), _mode(want_grad, value_only), c_void_p(result_ptr), c_void_p(stream)),
            "eb200_gp_eval_device",
        )

    def set_residual_device(self, r_ptr: int, stream: int = 0):
        _lib.check(self._lib.eb200_gp_set_residual_device(self._h, c_void_p(r_ptr), c_void_p(stream)), "set_residual_device")

    def predict_device(self, m: int, t_ptr: int, want_var: bool, mu_ptr: int, var_ptr: int, stream: int = 0):
        _lib.check(
            self._lib.eb200_gp_predict_device(
                self._h, m, c_void_p(t_ptr), int(bool(want_var)), c_void_p(mu_ptr), c_void_p(var_ptr), c_void_p(stream)
            ),
            "eb200_gp_predict_device",
        )

    def predict_grid_device(self, n_theta: int, theta_ptr: int, n_ind: int, ind_ptr: int, want_var: bool, mu_ptr: int,
                            var_ptr: int, stream: int = 0, round_f32: bool = True):
        """Grid prediction with theta[n_theta, d-1], independent[n_ind] and the results [n_theta * n_ind] in HBM."""
        _lib.check(
            self._lib.eb200_gp_predict_grid_device(
                self._h, n_theta, c_void_p(theta_ptr), n_ind, c_void_p(ind_ptr), int(bool(round_f32)), int(bool(want_var)),
                c_void_p(mu_ptr), c_void_p(var_ptr), c_void_p(stream),
            ),
            "eb200_gp_predict_grid_device",
        )

    def synchronize(self):
        _lib.check(self._lib.eb200_gp_synchronize(self._h), "synchronize")

    def launches_per_eval(self, want_grad: bool = True, value_only: bool = False) -> int:
        return self._lib.eb200_gp_eval_launches(self._h, _mode(want_grad, value_only))

    def profile_stages(self, p: np.ndarray, reps: int = 3) -> np.ndarray:
        """Device milliseconds per stage [params, factorisation (K + Cholesky + inverse), alpha, kinv+traces, finalize]."""
        p = np.ascontiguousarray(p, dtype=np.float64)
        out = (ctypes.c_float * 5)()
        _lib.check(self._lib.eb200_gp_profile_stages(self._h, _lib.as_double_ptr(p), reps, out), "profile_stages")
        return np.array(list(out), dtype=np.float64)

    # -- diagnostics (tests only) -------------------------------------------------------------
    def debug_stage(self, p: np.ndarray, stage: int):
        p = np.ascontiguousarray(p, dtype=np.float64)
        _lib.check(self._lib.eb200_gp_debug_stage(self._h, _lib.as_double_ptr(p), stage), "debug_stage")

    def debug_matrix(self, which: int = 0) -> np.ndarray:
        npad = self.padded_n
        out = np.empty((npad, npad))
        _lib.check(self._lib.eb200_gp_debug_matrix(self._h, which, _lib.as_double_ptr(out)), "debug_matrix")
        return out

    def debug_alpha(self) -> np.ndarray:
        out = np.empty(self.n)
        _lib.check(self._lib.eb200_gp_debug_alpha(self._h, _lib.as_double_ptr(out)), "debug_alpha")
        return out


def evaluate_many(problems, ps, want_grad: bool = True, value_only: bool = False):
    """
    One evaluation on each problem of `problems` (DeviceProblem, equal kernel dimension) at the rows of
    ps[nb, P], in ONE native call: the evaluations are enqueued on the problems' own streams and overlap
    on the device.  Returns (lml[nb], grad[nb, P], info[nb]).
    """
    nb = len(problems)
    ps = np.ascontiguousarray(ps, dtype=np.float64)
    if nb == 0:
        return np.zeros(0), np.zeros((0, ps.shape[1] if ps.ndim == 2 else 0)), np.zeros(0, dtype=np.int64)
    d = problems[0].d
    if ps.shape != (nb, d + 1) or any(q.d != d for q in problems):
        raise ValueError(f"need one parameter vector of {d + 1} entries per problem, all of kernel dimension {d}")
    if len({id(q) for q in problems}) != nb:
        raise ValueError("a problem appears twice: its matrices would be factorised twice at once")
    lib = problems[0]._lib
    length = _lib.result_len(d)
    out = np.zeros((nb, length))
    handles = (c_void_p * nb)(*[q._h for q in problems])
    _lib.check(
        lib.eb200_gp_eval_multi_host(handles, nb, _lib.as_double_ptr(ps), d + 1, _mode(want_grad, value_only),
                                     _lib.as_double_ptr(out), length),
        "eb200_gp_eval_multi_host",
    )
    return out[:, 0].copy(), out[:, 1 : d + 2].copy(), out[:, d + 2].astype(np.int64)


def evaluate_many_device(problems, p_ptr: int, p_stride: int, want_grad: bool, result_ptr: int, result_stride: int,
                         stream: int = 0, value_only: bool = False):
    """Device-pointer sibling of `evaluate_many`: parameter vectors [nb][p_stride] and result vectors
    [nb][result_stride] live in HBM; everything is enqueued on `stream`, nothing synchronises."""
    nb = len(problems)
    if nb == 0:
        return
    if len({id(q) for q in problems}) != nb:
        raise ValueError("a problem appears twice: its matrices would be factorised twice at once")
    lib = problems[0]._lib
    handles = (c_void_p * nb)(*[q._h for q in problems])
    _lib.check(
        lib.eb200_gp_eval_multi_device(handles, nb, c_void_p(p_ptr), p_stride, _mode(want_grad, value_only),
                                       c_void_p(result_ptr), result_stride, c_void_p(stream)),
        "eb200_gp_eval_multi_device",
    )


def profile_group_stages(problems, ps, reps: int = 3) -> np.ndarray:
    """Device milliseconds per stage of ONE grouped LML+grad evaluation of `problems` (up to 8, one shape) at the
    rows of ps: [params, factorisations (shared launch), alpha, K^-1 + traces, final sums]."""
    nb = len(problems)
    ps = np.ascontiguousarray(ps, dtype=np.float64)
    handles = (c_void_p * nb)(*[q._h for q in problems])
    out = (ctypes.c_float * 5)()
    _lib.check(problems[0]._lib.eb200_gp_profile_group_stages(handles, nb, _lib.as_double_ptr(ps), ps.shape[1], reps, out),
               "profile_group_stages")
    return np.array(list(out), dtype=np.float64)
