"""
Emulator protocol (same attributes and methods as
/root/reference/swiftemulator/emulators/base.py:21-172) plus the shared GP fit routine that
every GP emulator of the reference spells out inline
(gaussian_process.py:178-226, gaussian_process_bins.py:211-260,
gaussian_process_mcmc.py:200-256, gaussian_process_one_dim.py:177-225).
"""

from __future__ import annotations

import copy
from typing import Dict, Hashable, List, Optional

import attr
import numpy as np
from scipy.optimize import minimize

from .. import george_compat as george
from ..backend import ModelParameters, ModelSpecification, ModelValues


def default_kernel(ndim: int):
    """`1**2 * ExpSquaredKernel(np.ones(ndim), ndim=ndim)` -- the reference's default."""
    return 1**2 * george.kernels.ExpSquaredKernel(np.ones(ndim), ndim=ndim)


def build_gp(kernel, mean_model, x, y, yerr, device: Optional[int] = None):
    """Create the GP the way the reference does: deep copy of the kernel, frozen mean."""
    if mean_model is not None:
        mean_model.train(independent=x, dependent=y)
        gp = george.GP(copy.deepcopy(kernel), fit_kernel=True, mean=mean_model.george_model, fit_mean=False, device=device)
    else:
        gp = george.GP(copy.deepcopy(kernel), device=device)
    gp.compute(x=x, yerr=yerr)
    return gp


def optimise_gp(gp, y):
    """
    scipy.optimize.minimize (BFGS, default options) on the negative log likelihood with its
    analytic gradient, from the kernel's current parameter vector; the result is used whatever
    `success` says, exactly as the reference does (gaussian_process.py:217-224).
    """

    def negative_log_likelihood(p):
        gp.set_parameter_vector(p)
        return -gp.log_likelihood(y)

    def grad_negative_log_likelihood(p):
        gp.set_parameter_vector(p)
        return -gp.grad_log_likelihood(y)

    result = minimize(fun=negative_log_likelihood, x0=gp.get_parameter_vector(), jac=grad_negative_log_likelihood)
    gp.set_parameter_vector(result.x)
    return result


def optimise_gp_with_restarts(gp, kernel, mean_model, x, y, yerr, restarts: int, seed: int = 0, spread: float = 1.0,
                              workers: int = 4):
    """
    Optimiser restarts (an extension: the reference runs BFGS once, from the kernel's initial vector,
    gaussian_process.py:217-221; SURVEY.md section 8e lists restarts among the independent units that shard).
    Start 0 is that initial vector -- so `restarts=0` IS the reference's fit -- and starts 1..restarts are drawn
    N(x0, spread^2) from `seed`.  The starts are independent problems: handed out over ranks and host threads
    (dist.WorkClaimer), each with scipy's BFGS on its own GP over the same data, their likelihood evaluations
    combined into shared launches (emulator_b200/combine.py).  Every rank ends with `gp` at the best optimum found
    (lowest negative log-likelihood; ties go to the lowest start index).  Returns that start's OptimizeResult on the
    rank that ran it and a stub carrying x and fun elsewhere.
    """
    import threading

    from scipy.optimize import OptimizeResult

    from .. import dist
    from ..combine import EvaluationCombiner

    x0 = np.array(gp.get_parameter_vector(), dtype=np.float64)
    rng = np.random.default_rng(seed)
    starts = np.vstack([x0, x0 + spread * rng.standard_normal((restarts, len(x0)))])
    count = len(starts)
    claimer = dist.WorkClaimer(count)
    threads = max(1, min(workers, count)) if mean_model is None else 1  # a mean model object is shared state
    combiner = EvaluationCombiner(threads)
    results, errors = {}, []

    def work():
        try:
            with combiner:
                while True:
                    i = claimer.next()
                    if i is None:
                        break
                    if i == 0:
                        own = gp
                    else:
                        own = build_gp(kernel, mean_model, x, y, yerr, device=getattr(gp, "device", None))
                    own.set_parameter_vector(starts[i])
                    try:
                        results[i] = optimise_gp(own, y)
                    except np.linalg.LinAlgError:  # a start whose matrix cannot be factorised simply loses
                        results[i] = OptimizeResult(x=starts[i], fun=np.inf, success=False, status=-1, nit=0, nfev=0)
                    if own is not gp:
                        own.close()
        except BaseException as exc:
            errors.append(exc)

    if threads == 1:
        work()
    else:
        pool = [threading.Thread(target=work, daemon=True) for _ in range(threads)]
        for t in pool:
            t.start()
        for t in pool:
            t.join()
    if errors:
        raise errors[0]
    table = np.zeros((count, len(x0) + 1))
    for i, res in results.items():
        table[i, 0] = res.fun if np.isfinite(res.fun) else 1e300
        table[i, 1:] = res.x
    table = dist.all_reduce_sum(table)  # every start was run by exactly one rank
    best = int(np.argmin(table[:, 0]))
    gp.set_parameter_vector(table[best, 1:])
    chosen = results.get(best, OptimizeResult(x=table[best, 1:].copy(), fun=float(table[best, 0]), success=None, status=None))
    chosen.restart_index = best
    chosen.restart_funs = table[:, 0].copy()
    return chosen


def optimise_gps_lockstep(gps, ys):
    """
    The same optimisation for many equally shaped GPs at once (per-bin emulators): all problems advance
    together, one batched device call per round (emulator_b200/optimise.py).  Starts from each kernel's
    current parameter vector and leaves every GP at its optimum, whatever the status says -- as
    optimise_gp does.  Opt-in: the optimiser is not scipy's, so iterates (not optima) differ.
    """
    from ..optimise import minimise_lockstep

    backend = george
    if not hasattr(backend, "pinned_evaluator"):
        raise NotImplementedError("lock-step fitting needs the device backend (george_compat)")
    with backend.pinned_evaluator(gps, ys) as evaluate:

        def fun_and_grad(indices, points):
            lml, grad, info = evaluate(list(indices), points)
            bad = (info != 0) | ~np.isfinite(lml)
            return np.where(bad, np.inf, -lml), -grad

        result = minimise_lockstep(fun_and_grad, np.array([gp.get_parameter_vector() for gp in gps]))
    for gp, x in zip(gps, result.x):
        gp.set_parameter_vector(x)
    return result


@attr.s
class BaseEmulator(object):
    ordering: Optional[List[Hashable]] = None
    parameter_order: Optional[List[str]] = None

    model_specification: Optional[ModelSpecification] = None
    model_parameters: Optional[ModelParameters] = None
    model_values: Optional[ModelValues] = None

    independent_variables: Optional[np.array] = None
    dependent_variables: Optional[np.array] = None
    dependent_variable_errors: Optional[np.array] = None

    emulator = None
    #: scipy OptimizeResult of the last hyperparameter fit (not in the reference; diagnostics)
    fit_result = None

    def _build_arrays(self, model_specification, model_parameters, model_values):
        raise NotImplementedError

    def fit_model(self, model_specification, model_parameters, model_values):
        raise NotImplementedError

    def predict_values(self, independent: np.array, model_parameters: Dict[str, float]):
        raise NotImplementedError

    def predict_values_no_error(self, independent: np.array, model_parameters: Dict[str, float]):
        raise NotImplementedError

    def interactive_plot(self, *args, **kwargs):
        raise NotImplementedError("plotting (matplotlib) is outside the accelerated hot path")


def flatten_model_values(model_specification, model_parameters, model_values, with_independent: bool):
    """
    Flatten {uid: {independent, dependent[, dependent_error]}} into the float32 design matrix
    the reference builds row by row (gaussian_process.py:89-131; 1-D variant without the
    independent column gaussian_process_one_dim.py:88-121).  Row order = dict order.
    Returns (X, y, yerr, ordering).
    """
    names = model_specification.parameter_names
    n_par = model_specification.number_of_parameters
    total = model_values.number_of_variables
    width = n_par + (1 if with_independent else 0)
    x = np.empty((total, width), dtype=np.float32)
    y = np.empty(total, dtype=np.float32)
    yerr = np.empty(total, dtype=np.float32)
    ordering = []
    row = 0
    for uid in model_values.model_values.keys():
        ordering.append(uid)
        theta = np.array([model_parameters.model_parameters[uid][name] for name in names])
        relation = model_values.model_values[uid]
        dependent = relation["dependent"]
        count = len(relation["independent"]) if with_independent else len(dependent)
        errors = relation.get("dependent_error", np.zeros(count))
        if np.ndim(errors) != 1:
            raise AttributeError("Multiple dimensional errors are not currently supported in GPE mode")
        stop = row + count
        if with_independent:
            x[row:stop, 0] = relation["independent"]
            x[row:stop, 1:] = theta
        else:
            x[row:stop, :] = theta
        y[row:stop] = dependent[:count]
        yerr[row:stop] = errors[:count]
        row = stop
    assert row == total
    return x, y, yerr, ordering


def query_matrix(independent, model_parameters: Dict[str, float], parameter_order) -> np.ndarray:
    """float32 query rows [independent_i, theta...] (gaussian_process.py:273-283)."""
    theta = np.array([model_parameters[name] for name in parameter_order])
    t = np.empty((len(independent), len(theta) + 1), dtype=np.float32)
    t[:, 0] = independent
    t[:, 1:] = theta
    return t
