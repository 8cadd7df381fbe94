"""
Generates tests/golden/*.npz by running the REFERENCE's own emulator code
(/root/reference/swiftemulator, imported in this container) on seeded inputs, with `george`
replaced by the oracle-backed shim (tests/oracle_george.py) because george itself is not
installable here.  The vectors pin (a) the reference's data flattening / float32 rounding /
optimiser flow and (b) the oracle's numbers, and travel to the GPU box where the CUDA path is
checked against them.

    python tests/golden/make_golden.py        (needs /root/reference; CPU only)
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import oracle_george  # noqa: E402

oracle_george.install_reference_import_stubs()

import swiftemulator as se  # noqa: E402  (the reference)
from swiftemulator.emulators.gaussian_process import GaussianProcessEmulator as RefGPE  # noqa: E402
from swiftemulator.emulators.gaussian_process_bins import GaussianProcessEmulatorBins as RefBins  # noqa: E402
from swiftemulator.emulators.gaussian_process_one_dim import GaussianProcessEmulator1D as Ref1D  # noqa: E402
from swiftemulator.emulators.multi_gaussian_process import MultipleGaussianProcessEmulator as RefMulti  # noqa: E402
from swiftemulator.mean_models.linear import LinearMeanModel as RefLinear  # noqa: E402
from swiftemulator.mean_models.offset import OffsetMeanModel as RefOffset  # noqa: E402
from swiftemulator.sensitivity.cross_check import CrossCheck as RefCrossCheck  # noqa: E402

from emulator_b200 import synthetic  # noqa: E402
from oracle import gp_oracle  # noqa: E402


def to_reference(spec, params, values):
    rspec = se.ModelSpecification(
        number_of_parameters=spec.number_of_parameters, parameter_names=spec.parameter_names,
        parameter_limits=spec.parameter_limits,
    )
    return rspec, se.ModelParameters(model_parameters=params.model_parameters), se.ModelValues(
        model_values={k: dict(v) for k, v in values.model_values.items()}
    )


def reference_test_fixture(seed):
    """The data shape of the reference's own smoke tests (tests/test_emulator_gpe.py:21-58),
    seeded: 2 parameters, 3 sims, independent = arange(10), random dependents, one sim without
    dependent_error."""
    from emulator_b200.backend import ModelParameters, ModelSpecification, ModelValues

    rng = np.random.default_rng(seed)
    spec = ModelSpecification(number_of_parameters=2, parameter_names=["x", "y"], parameter_limits=[[0.0, 1.0], [8.0, 9.0]])
    params = ModelParameters(model_parameters={0: {"x": 0.1, "y": 8.1}, 1: {"x": 0.5, "y": 8.7}, 2: {"x": 0.9, "y": 8.4}})
    values = {}
    for uid in range(3):
        entry = {"independent": np.arange(10, dtype=np.float64), "dependent": rng.random(10)}
        if uid != 1:
            entry["dependent_error"] = 0.1 * rng.random(10)
        values[uid] = entry
    return spec, params, ModelValues(model_values=values)


def ragged_fixture(seed):
    """Simulations with DIFFERENT numbers of points at different independent values (the reference flattens whatever
    each simulation holds, emulators/gaussian_process.py:89-131), one of them without errors."""
    from emulator_b200.backend import ModelParameters, ModelSpecification, ModelValues

    rng = np.random.default_rng(seed)
    names = ["a", "b", "c"]
    spec = ModelSpecification(number_of_parameters=3, parameter_names=names, parameter_limits=[[0.0, 1.0], [-1.0, 1.0], [2.0, 4.0]])
    lengths = [4, 9, 6, 11, 7, 1, 8, 5, 10, 3, 12, 6, 9, 2, 7, 8]
    params, values = {}, {}
    for uid, n in enumerate(lengths):
        theta = {"a": float(rng.random()), "b": float(2 * rng.random() - 1), "c": float(2 + 2 * rng.random())}
        params[f"run{uid}"] = theta
        x = np.sort(2.0 * rng.random(n))
        y = np.sin(1.5 * x + 2.0 * theta["a"]) * theta["c"] + theta["b"] * x * x + 0.01 * rng.standard_normal(n)
        entry = {"independent": x, "dependent": y}
        if uid != 5:  # (the single-point simulation: an error-less one with close points would put cond(K) at 1e13)
            entry["dependent_error"] = 0.02 + 0.02 * rng.random(n)
        values[f"run{uid}"] = entry
    return spec, ModelParameters(model_parameters=params), ModelValues(model_values=values)


def golden_gpe(name, n_par, sims, bins, seed, mean_model=None, fixture=False):
    if fixture == 2:
        spec, params, values = ragged_fixture(seed)
    elif fixture:
        spec, params, values = reference_test_fixture(seed)
    else:
        spec, params, values = synthetic.make_model(n_par, sims, bins, seed=seed)
    rspec, rparams, rvalues = to_reference(spec, params, values)
    emu = RefGPE(mean_model=mean_model)
    emu._build_arrays(rspec, rparams, rvalues)
    x, y, yerr = emu.independent_variables.copy(), emu.dependent_variables.copy(), emu.dependent_variable_errors.copy()
    d = n_par + 1
    p0 = gp_oracle.default_parameter_vector(d)
    # likelihood + gradient at x0 and at perturbed vectors, through the oracle GP the reference drives
    probes = synthetic.perturbed_parameter_vectors(d, count=4, seed=seed + 5, scale=0.3)
    if mean_model is not None:
        mm = mean_model.copy(); mm.train(x, y); mean = mm.predict
    else:
        mean = None
    gp = gp_oracle.OracleGP(d, p0, mean=mean)
    gp.compute(x, yerr)
    lml = []; grad = []
    for p in probes:
        gp.set_parameter_vector(p)
        lml.append(gp.log_likelihood(y)); grad.append(gp.grad_log_likelihood(y))
    emu.fit_model(rspec, rparams, rvalues)
    p_fit = emu.emulator.get_parameter_vector()
    query_x = np.linspace(0.05, 0.95, 7) if not fixture else (np.linspace(0.0, 2.0, 9) if fixture == 2 else np.arange(10, dtype=np.float64))
    lo = np.array(spec.parameter_limits)[:, 0]
    theta = {nm: float(lo[i]) + 0.37 + 0.05 * i for i, nm in enumerate(spec.parameter_names)}
    mu, var = emu.predict_values(query_x, theta)
    mu_only = emu.predict_values_no_error(query_x, theta)
    np.savez(
        os.path.join(HERE, name + ".npz"), n_par=n_par, sims=sims, bins=bins, seed=seed, fixture=int(fixture),
        x=x, y=y, yerr=yerr, probes=probes, lml=np.array(lml), grad=np.array(grad), p_fit=p_fit,
        query_x=query_x, theta=np.array([theta[nm] for nm in spec.parameter_names]), mu=mu, var=var, mu_only=mu_only,
    )
    print(name, "N =", len(y), "p_fit =", np.round(p_fit, 4))


def golden_bins(name, n_par, sims, bins, seed):
    spec, params, values = synthetic.make_model(n_par, sims, bins, seed=seed)
    rspec, rparams, rvalues = to_reference(spec, params, values)
    emu = RefBins()
    emu.fit_model(rspec, rparams, rvalues)
    p_fit = np.array([gp.get_parameter_vector() for gp in emu.bin_gaussian_process])
    theta = {nm: 0.41 + 0.07 * i for i, nm in enumerate(spec.parameter_names)}
    centers = np.array(emu.bin_centers)
    mu, var = emu.predict_values(centers, theta)
    np.savez(
        os.path.join(HERE, name + ".npz"), n_par=n_par, sims=sims, bins=bins, seed=seed, p_fit=p_fit, centers=centers,
        theta=np.array([theta[nm] for nm in spec.parameter_names]), mu=mu, var=var,
        bin0_x=emu.bin_model_values[0]["independent"], bin0_y=emu.bin_model_values[0]["dependent"],
        bin0_err=emu.bin_model_values[0]["dependent_errors"],
    )
    print(name, "bins =", len(centers))


def golden_one_dim(name, n_par, sims, seed):
    spec, params, values = synthetic.make_model(n_par, sims, 1, seed=seed)
    rspec, rparams, rvalues = to_reference(spec, params, values)
    emu = Ref1D()
    emu.fit_model(rspec, rparams, rvalues)
    theta = {nm: 0.3 + 0.1 * i for i, nm in enumerate(spec.parameter_names)}
    mu, var = emu.predict_values(theta)
    np.savez(
        os.path.join(HERE, name + ".npz"), n_par=n_par, sims=sims, seed=seed, p_fit=emu.emulator.get_parameter_vector(),
        theta=np.array([theta[nm] for nm in spec.parameter_names]), mu=mu, var=var, x=emu.independent_variables,
    )
    print(name)


def golden_cross_check(name, n_par, sims, bins, seed):
    spec, params, values = synthetic.make_model(n_par, sims, bins, seed=seed)
    rspec, rparams, rvalues = to_reference(spec, params, values)
    cc = RefCrossCheck()
    cc.build_emulators(rspec, rparams, rvalues)
    total, per_model = cc.get_mean_squared()
    p_fit = np.array([cc.cross_emulators[uid].emulator.get_parameter_vector() for uid in cc.leave_out_order])
    np.savez(os.path.join(HERE, name + ".npz"), n_par=n_par, sims=sims, bins=bins, seed=seed, total=total, p_fit=p_fit,
             per_model=np.array([per_model[uid] for uid in cc.leave_out_order]))
    print(name, "mean squared =", total)


def golden_multi(name, n_par, sims, bins, seed):
    """The reference's multi-region emulator (emulators/multi_gaussian_process.py:163-330): one GP per region of the
    independent axis, fitted one after the other, blended linearly where consecutive regions overlap.  Two layouts:
    regions that overlap (means are pinned everywhere; inside the overlap the reference takes the right-hand ERROR
    from the left emulator's output list, :312-314 -- recorded, and next to it the per-region outputs the documented
    formula needs) and regions that only touch."""
    spec, params, values = synthetic.make_model(n_par, sims, bins, seed=seed)
    rspec, rparams, rvalues = to_reference(spec, params, values)
    lo = np.array(spec.parameter_limits)[:, 0]
    theta = {nm: float(lo[i]) + 0.33 + 0.06 * i for i, nm in enumerate(spec.parameter_names)}
    xs = values.model_values[next(iter(values.model_values))]["independent"]
    query = np.sort(np.concatenate([xs, 0.5 * (xs[1:] + xs[:-1])]))
    out = dict(n_par=n_par, sims=sims, bins=bins, seed=seed, query=query,
               theta=np.array([theta[nm] for nm in spec.parameter_names]))
    for label, regions in (("overlap", [[None, 0.62], [0.38, None]]), ("touch", [[None, 0.5], [0.5, None]]),
                           ("three", [[None, 0.45], [0.3, 0.8], [0.65, None]])):
        emu = RefMulti(independent_regions=regions)
        emu.fit_model(rspec, rparams, rvalues)
        # (points that fall in no region -- exactly on a touching boundary -- are left out: the reference fails there)
        inside = np.zeros(len(query), dtype=bool)
        for low, high in regions:
            inside |= np.logical_and(query > low if low is not None else True, query < high if high is not None else True)
        q = query[inside]
        mu, err = emu.predict_values(q, theta)
        out[label + "_regions"] = np.array([[np.nan if v is None else v for v in r] for r in regions])
        out[label + "_query"] = q
        out[label + "_mu"] = mu
        out[label + "_err"] = err
        out[label + "_p_fit"] = np.array([e.emulator.get_parameter_vector() for e in emu.emulators])
        out[label + "_rows"] = np.array([len(e.dependent_variables) for e in emu.emulators])
        for index, e in enumerate(emu.emulators):  # each region's own prediction over the points it covers
            low, high = regions[index]
            mask = np.logical_and(q > low if low is not None else True, q < high if high is not None else True)
            m, v = e.predict_values(q[mask], theta)
            full_m, full_v = np.full(len(q), np.nan), np.full(len(q), np.nan)
            full_m[mask], full_v[mask] = m, v
            out[f"{label}_region{index}_mu"] = full_m
            out[f"{label}_region{index}_err"] = full_v
    np.savez(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: out[k] for k in ("overlap_rows", "touch_rows", "three_rows")})


def golden_mocking(name, n_par, sims, bins, seed):
    """The reference's prediction sweeps (mocking/__init__.py:17-97 mock_hypercube, :100-200 mock_sweep): one
    `predict_values` call per parameter row, at every independent value of the training data."""
    from swiftemulator.mocking import mock_hypercube as ref_hypercube, mock_sweep as ref_sweep

    spec, params, values = synthetic.make_model(n_par, sims, bins, seed=seed)
    rspec, rparams, rvalues = to_reference(spec, params, values)
    emu = RefGPE()
    emu.fit_model(rspec, rparams, rvalues)
    np.random.seed(seed)  # the reference draws its cube from numpy's global generator (design/random.py:53)
    cube_values, cube_parameters = ref_hypercube(emu, rspec, 9)
    lo, hi = np.array(spec.parameter_limits)[:, 0], np.array(spec.parameter_limits)[:, 1]
    center = {nm: float(lo[i] + (0.4 + 0.1 * i) * (hi[i] - lo[i])) for i, nm in enumerate(spec.parameter_names)}
    swept = spec.parameter_names[1]
    sweep_values, sweep_parameters = ref_sweep(emu, rspec, 7, swept, center)

    def pack(vals, pars):
        uids = list(pars.model_parameters.keys())
        return (np.array([[pars.model_parameters[u][nm] for nm in spec.parameter_names] for u in uids]),
                np.array([vals.model_values[u]["dependent"] for u in uids]),
                np.array([vals.model_values[u]["dependent_error"] for u in uids]),
                vals.model_values[uids[0]]["independent"], np.array(uids))

    cube = pack(cube_values, cube_parameters)
    sweep = pack(sweep_values, sweep_parameters)
    np.savez(os.path.join(HERE, name + ".npz"), n_par=n_par, sims=sims, bins=bins, seed=seed,
             p_fit=emu.emulator.get_parameter_vector(), center=np.array([center[nm] for nm in spec.parameter_names]),
             swept=swept, cube_theta=cube[0], cube_mu=cube[1], cube_err=cube[2], independent=cube[3], cube_uids=cube[4],
             sweep_theta=sweep[0], sweep_mu=sweep[1], sweep_err=sweep[2], sweep_uids=sweep[4])
    print(name, cube[1].shape, sweep[1].shape)


def golden_cross_check_bins(name, n_par, sims, bins, seed):
    """The reference's leave-one-out loop over per-bin emulators (sensitivity/cross_check_bins.py:94-109): one
    GaussianProcessEmulatorBins per left-out simulation.  (Its post-processing is broken in the reference --
    :142-148, :224-229 unpack three values from a two-value predict -- so only the fits are pinned.)"""
    from swiftemulator.sensitivity.cross_check_bins import CrossCheckBins as RefCrossCheckBins

    spec, params, values = synthetic.make_model(n_par, sims, bins, seed=seed)
    rspec, rparams, rvalues = to_reference(spec, params, values)
    cc = RefCrossCheckBins()
    cc.build_emulators(rspec, rparams, rvalues)
    p_fit = np.array([[gp.get_parameter_vector() for gp in cc.cross_emulators[uid].bin_gaussian_process]
                      for uid in cc.leave_out_order])
    theta = {nm: 0.37 + 0.09 * i for i, nm in enumerate(spec.parameter_names)}
    centers = np.array(cc.cross_emulators[cc.leave_out_order[0]].bin_centers)
    pred = np.array([cc.cross_emulators[uid].predict_values(centers, theta) for uid in cc.leave_out_order])
    np.savez(os.path.join(HERE, name + ".npz"), n_par=n_par, sims=sims, bins=bins, seed=seed, p_fit=p_fit,
             order=np.array(cc.leave_out_order), centers=centers, theta=np.array([theta[nm] for nm in spec.parameter_names]),
             mu=pred[:, 0], var=pred[:, 1])
    print(name, p_fit.shape)


class MinimizeRecorder:
    """Stands in for `scipy.optimize.minimize` inside a reference module (they do `from scipy.optimize import minimize`):
    the same call plus a callback that records the iterates x_k -- a callback does not change what BFGS does."""

    def __init__(self):
        self.results, self.iterates = [], []

    def __call__(self, fun, x0, jac=None, **kwargs):
        from scipy.optimize import minimize

        its = [np.array(x0, dtype=np.float64)]
        res = minimize(fun, x0, jac=jac, callback=lambda xk: its.append(np.array(xk, dtype=np.float64)), **kwargs)
        self.results.append(res)
        self.iterates.append(np.array(its))
        return res


def golden_trajectory(name, n_par, sims, bins, seed):
    """The BFGS run of the reference's GaussianProcessEmulator.fit_model (gaussian_process.py:217-224), iterate by
    iterate: what "identical hyperparameter optima for the same optimiser" is checked against on the device."""
    import swiftemulator.emulators.gaussian_process as ref_module

    spec, params, values = synthetic.make_model(n_par, sims, bins, seed=seed)
    rspec, rparams, rvalues = to_reference(spec, params, values)
    recorder = MinimizeRecorder()
    keep = ref_module.minimize
    ref_module.minimize = recorder
    try:
        emu = RefGPE()
        emu.fit_model(rspec, rparams, rvalues)
    finally:
        ref_module.minimize = keep
    res = recorder.results[0]
    np.savez(
        os.path.join(HERE, name + ".npz"), n_par=n_par, sims=sims, bins=bins, seed=seed, iterates=recorder.iterates[0],
        nit=res.nit, nfev=res.nfev, njev=res.njev, status=res.status, fun=res.fun, p_fit=res.x,
    )
    print(name, "N =", len(emu.dependent_variables), "nit", res.nit, "nfev", res.nfev, "njev", res.njev, "status", res.status)


def golden_bins_trajectories(name, n_par, sims, bins, seed):
    """A binned emulator at BASELINE configuration C3's shape: 64 per-bin GPs of N = 256, d = 8."""
    import swiftemulator.emulators.gaussian_process_bins as ref_module

    spec, params, values = synthetic.make_model(n_par, sims, bins, seed=seed)
    rspec, rparams, rvalues = to_reference(spec, params, values)
    recorder = MinimizeRecorder()
    keep = ref_module.minimize
    ref_module.minimize = recorder
    try:
        emu = RefBins()
        emu.fit_model(rspec, rparams, rvalues)
    finally:
        ref_module.minimize = keep
    p_fit = np.array([gp.get_parameter_vector() for gp in emu.bin_gaussian_process])
    theta = {nm: 0.41 + 0.05 * i for i, nm in enumerate(spec.parameter_names)}
    centers = np.array(emu.bin_centers)
    mu, var = emu.predict_values(centers, theta)
    np.savez(
        os.path.join(HERE, name + ".npz"), n_par=n_par, sims=sims, bins=bins, seed=seed, p_fit=p_fit, centers=centers,
        theta=np.array([theta[nm] for nm in spec.parameter_names]), mu=mu, var=var,
        nit=np.array([r.nit for r in recorder.results]), nfev=np.array([r.nfev for r in recorder.results]),
        fun=np.array([r.fun for r in recorder.results]),
    )
    print(name, "bins =", len(centers), "nit", [r.nit for r in recorder.results][:8], "...")


class _Quantity(np.ndarray):
    """The two unyt methods the reference's penalty module calls, on a plain array (unyt is absent here)."""

    def to(self, units):
        return self

    @property
    def value(self):
        return np.asarray(self)


class _Scalar(float):
    def convert_to_units(self, units):
        return None


def _quantity(a):
    return np.asarray(a, dtype=np.float64).view(_Quantity)


PENALTY_CASES = [
    ("L1PenaltyCalculator", dict(offset=5.0, lower=1.0, upper=4.0)),
    ("L1PenaltyCalculatorOneSided", dict(offset=5.0, maximum_penalty="above", lower=1.0, upper=4.0)),
    ("L1PenaltyCalculatorOneSided", dict(offset=4.0, maximum_penalty="below", lower=1.5, upper=4.2)),
    ("L2PenaltyCalculator", dict(offset=5.0, lower=1.0, upper=4.0)),
    ("L1VariablePenaltyCalculator", dict(offset_lower=2.0, offset_upper=6.0, offset_transition=2.5, transition_width=0.5, lower=1.0,
                                         upper=4.0, offset_below_above_ratio=0.5)),
    ("L1SqueezePenaltyCalculator", dict(offset_squeeze=1.0, offset_normal=5.0, offset_transition=2.5, transition_width=0.5, lower=1.0,
                                        upper=4.0, offset_below_above_ratio=2.0)),
    ("GaussianDataErrorsPenaltyCalculator", dict(sigma_max=3.0, lower=1.0, upper=4.0)),
    ("GaussianPercentErrorsPenaltyCalculator", dict(percent_error=10.0, sigma_max=3.0, lower=1.0, upper=4.0)),
    ("GaussianDataErrorsPercentFloorPenaltyCalculator", dict(percent_error=10.0, sigma_max=2.0, lower=1.0, upper=4.0)),
    ("GaussianWeightedDataErrorsPercentFloorPenaltyCalculator", dict(percent_error=10.0, sigma_max=3.0, weight=2.0, lower=1.0, upper=4.0)),
]


def golden_penalties(name):
    """The reference's penalty calculators (comparison/penalty.py) on seeded sweeps, in linear and in log space."""
    # (absent third-party names the reference's module imports at the top; nothing of them is executed)
    sys.modules["velociraptor.observations"].ObservationalData = object
    sys.modules["unyt"].unyt_quantity = float
    import scipy.interpolate  # noqa: F401
    import swiftemulator.comparison.penalty as ref

    rng = np.random.default_rng(77)
    obs_x = np.array([3.0, 1.0, 2.0, 4.0, 2.5])
    obs_y = 10.0 * obs_x + rng.normal(0, 1.0, 5)
    scatter = np.abs(rng.normal(2.0, 0.5, (2, 5)))  # unsymmetric errors
    independent = np.linspace(0.5, 4.5, 11)
    dependent = rng.normal(10.0 * independent, 4.0, size=(16, 11))
    out = {"obs_x": obs_x, "obs_y": obs_y, "scatter": scatter, "independent": independent, "dependent": dependent}

    class Obs:
        x, y, y_scatter = _quantity(obs_x), _quantity(obs_y), _quantity(scatter)

    for index, (cls, kwargs) in enumerate(PENALTY_CASES):
        for log_space in (False, True):
            if log_space:  # limits are plain numbers in log space; the model is compared in log10 of both axes
                kw = {k: (float(np.log10(v)) if k in ("lower", "upper", "offset_transition") else v) for k, v in kwargs.items()}
                kw = {k: (0.3 * v / 5.0 if k.startswith("offset") and k not in ("offset_transition", "offset_below_above_ratio") else v)
                      for k, v in kw.items()}
                ind, dep = np.log10(independent), np.log10(np.abs(dependent) + 1.0)
            else:
                kw = {k: (_Scalar(v) if isinstance(v, float) and k not in ("percent_error", "sigma_max", "weight", "offset_below_above_ratio") else v)
                      for k, v in kwargs.items()}
                ind, dep = independent, dependent
            calc = getattr(ref, cls)(**kw)
            calc.register_observation(Obs, log_space, log_space, None, None)
            rows = []
            for r in range(len(dep)):
                pen = calc.penalty(ind, dep[r].copy())
                rows.append(np.full(ind.shape, np.nan) if np.isscalar(pen) else np.where(
                    np.logical_and(ind >= float(kw["lower"]), ind < float(kw["upper"])),
                    np.zeros_like(ind), np.nan))
                if not np.isscalar(pen):
                    rows[-1][~np.isnan(rows[-1])] = pen
            out[f"case{index}_{'log' if log_space else 'lin'}"] = np.array(rows)
    np.savez(os.path.join(HERE, name + ".npz"), **out)
    print(name, len(PENALTY_CASES), "calculators x 2 spaces")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "penalties":
        golden_penalties("penalties")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ragged":
        golden_gpe("gpe_ragged", 3, 6, 0, seed=112, fixture=2)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "polynomial":
        from swiftemulator.mean_models.polynomial import PolynomialMeanModel as RefPolynomial
        golden_gpe("gpe_polynomial_mean", 3, 18, 8, seed=111, mean_model=RefPolynomial(degree=2))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cross_check_bins":
        golden_cross_check_bins("cross_check_bins_small", 2, 9, 4, seed=110)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "mocking":
        golden_mocking("mocking_small", 3, 16, 9, seed=109)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "multi":
        golden_multi("multi_regions", 2, 14, 12, seed=108)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "round2":  # the vectors added in round 2 only
        golden_trajectory("trajectory_c1", 2, 50, 20, seed=1235)       # BASELINE configuration C1: N = 1000, d = 3
        golden_trajectory("trajectory_c2_shape", 8, 64, 16, seed=201)  # C2's shape at a quarter of its rows: N = 1024, d = 9
        golden_bins_trajectories("bins_c3", 8, 256, 64, seed=1237)     # BASELINE configuration C3 in full
        sys.exit(0)
    golden_gpe("gpe_small", 2, 12, 10, seed=101)
    golden_gpe("gpe_noerr", 2, 3, 10, seed=102, fixture=True)
    golden_gpe("gpe_linear_mean", 3, 20, 8, seed=103, mean_model=RefLinear())
    golden_gpe("gpe_offset_mean", 2, 10, 6, seed=104, mean_model=RefOffset())
    golden_bins("bins_small", 2, 24, 5, seed=105)
    golden_one_dim("one_dim_small", 3, 30, seed=106)
    golden_cross_check("cross_check_small", 2, 8, 6, seed=107)
