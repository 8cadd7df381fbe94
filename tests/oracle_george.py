"""
TEST INFRASTRUCTURE: a `george`-shaped module whose GP does its arithmetic in
oracle/gp_oracle.py (numpy/scipy).  Used (a) to run the reference's own emulator code in this
container when generating golden vectors (tests/golden/make_golden.py) and (b) as a stand-in
for the device GP in CPU-only host-logic tests.  Never imported by the product.
"""

from __future__ import annotations

import contextlib
import sys
import types

import numpy as np

from emulator_b200.george_compat import kernels, modeling
from oracle import gp_oracle


class OracleBackedGP:
    """george.GP signature subset; parameters live in the (george_compat) kernel object."""

    def __init__(self, kernel=None, fit_kernel=True, mean=None, fit_mean=False, device=0, **kwargs):
        self.ndim = kernels.device_form(kernel)
        self.kernel = kernel
        if mean is None:
            self.mean = modeling.ConstantModel(0.0)
        elif isinstance(mean, modeling.Model):
            self.mean = mean
        else:
            self.mean = modeling.CallableModel(mean)
        self._gp = gp_oracle.OracleGP(self.ndim, self.kernel.get_parameter_vector(), mean=self.mean.get_value)
        self.eager_gradient = True
        self.computed = False

    def __len__(self):
        return len(self.kernel)

    def get_parameter_vector(self):
        return self.kernel.get_parameter_vector()

    def set_parameter_vector(self, p):
        self.kernel.set_parameter_vector(np.asarray(p, dtype=np.float64))
        self._gp.set_parameter_vector(p)

    def get_parameter_names(self):
        return tuple(f"kernel:{n}" for n in self.kernel.get_parameter_names())

    def compute(self, x, yerr=0.0, **kwargs):
        self._gp.compute(x, yerr)
        self.computed = True

    def log_likelihood(self, y, quiet=False):
        try:
            return self._gp.log_likelihood(y)
        except np.linalg.LinAlgError:
            if quiet:
                return -np.inf
            raise

    def log_likelihood_batch(self, y, vectors, quiet=True):
        keep = self.get_parameter_vector().copy()
        out = []
        for v in np.atleast_2d(vectors):
            self.set_parameter_vector(v)
            out.append(self.log_likelihood(y, quiet=quiet))
        self.set_parameter_vector(keep)
        return np.array(out, dtype=np.float64)

    def grad_log_likelihood(self, y, quiet=False):
        try:
            return self._gp.grad_log_likelihood(y)
        except np.linalg.LinAlgError:
            if quiet:
                return np.zeros(len(self))
            raise

    def predict(self, y, t, return_cov=True, return_var=False, **kwargs):
        return self._gp.predict(y, t, return_cov=return_cov, return_var=return_var)

    def predict_grid(self, y, theta, independent, return_var=True, round_f32=True):
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        independent = np.asarray(independent, dtype=np.float64)
        rows = np.empty((len(theta) * len(independent), self.ndim), dtype=np.float32 if round_f32 else np.float64)
        rows[:, 0] = np.tile(independent, len(theta))
        rows[:, 1:] = np.repeat(theta, len(independent), axis=0)
        out = self.predict(y, rows.astype(np.float64), return_cov=False, return_var=return_var)
        shape = (len(theta), len(independent))
        return (out[0].reshape(shape), out[1].reshape(shape)) if return_var else out.reshape(shape)

    def close(self):
        pass


@contextlib.contextmanager
def pinned_evaluator(gps, ys):
    """Oracle counterpart of george_compat.pinned_evaluator (lock-step optimisers)."""

    def evaluate(indices, vectors, want_grad=True):
        lml, grad, info = [], [], []
        for i, v in zip(indices, vectors):
            g = gps[i]
            keep = g.get_parameter_vector().copy()
            g.set_parameter_vector(v)
            try:
                lml.append(g._gp.log_likelihood(ys[i]))
                grad.append(g._gp.grad_log_likelihood(ys[i]) if want_grad else np.zeros(len(g)))
                info.append(0)
            except np.linalg.LinAlgError:
                lml.append(np.nan)
                grad.append(np.zeros(len(g)))
                info.append(1)
            g.set_parameter_vector(keep)
        return np.array(lml), np.array(grad), np.array(info)

    yield evaluate


def fake_george_module():
    mod = types.ModuleType("george")
    mod.kernels = kernels
    mod.modeling = modeling
    mod.GP = OracleBackedGP
    return mod


def install_reference_import_stubs():
    """
    Make /root/reference/swiftemulator importable in this container: `george` is the
    oracle-backed shim, the plotting / design packages that are absent get empty stubs.
    """
    george = fake_george_module()
    sys.modules["george"] = george
    sys.modules["george.kernels"] = kernels
    sys.modules["george.modeling"] = modeling
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.widgets", "matplotlib.colors", "corner", "pyDOE", "emcee", "SALib",
                 "SALib.analyze", "SALib.analyze.rbd_fast", "unyt", "velociraptor", "velociraptor.observations"):
        if name not in sys.modules:
            stub = types.ModuleType(name)
            sys.modules[name] = stub
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["pyDOE"].lhs = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
