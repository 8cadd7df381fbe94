"""
The emulators on the GPU against the golden vectors produced by the reference's own emulator
code (tests/golden/make_golden.py) and against the oracle-backed flow on the same inputs.
"""
import os

import numpy as np
import pytest

from emulator_b200 import mocking, synthetic
from emulator_b200.emulators import (
    GaussianProcessEmulator, GaussianProcessEmulator1D, GaussianProcessEmulatorBins, GaussianProcessEmulatorMCMC,
    MultipleGaussianProcessEmulator,
)
from emulator_b200.mean_models import LinearMeanModel, OffsetMeanModel, PolynomialMeanModel
from emulator_b200.sensitivity import CrossCheck, CrossCheckBins, sobol_indices
from test_host_logic import _regions, check_mocking_against_reference, check_multi_against_reference, fixture_data, golden_inputs, load, theta_dict

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize(
    "name,mean_model",
    [("gpe_small", None), ("gpe_noerr", None), ("gpe_ragged", None), ("gpe_linear_mean", LinearMeanModel()), ("gpe_offset_mean", OffsetMeanModel()),
     ("gpe_polynomial_mean", PolynomialMeanModel(degree=2))],
)
def test_gpe_fit_and_predict_match_reference_golden(name, mean_model):
    g = load(name)
    spec, params, values = golden_inputs(g)
    emu = GaussianProcessEmulator(mean_model=mean_model)
    emu.fit_model(spec, params, values)
    assert np.array_equal(emu.independent_variables, g["x"])
    gp = emu.emulator
    # likelihood / gradient at the golden probe vectors
    if mean_model is None:
        for p, lml, grad in zip(g["probes"], g["lml"], g["grad"]):
            gp.set_parameter_vector(p)
            assert gp.log_likelihood(emu.dependent_variables) == pytest.approx(lml, rel=1e-9)
            got = gp.grad_log_likelihood(emu.dependent_variables)
            assert np.max(np.abs(got - grad)) <= 1e-9 * np.max(np.abs(grad))
        gp.set_parameter_vector(emu.fit_result.x)
    # "identical hyperparameter optima": the same scipy BFGS runs on values that agree to ~1e-12,
    # but it stops on "precision loss" (as in the reference, whose result.success is ignored), i.e.
    # in a flat valley where last-bit differences move the end point by ~1e-5.  The optimum is
    # therefore compared at the optimiser's resolution, and the likelihood reached must be equal.
    assert np.allclose(emu.fit_result.x, g["p_fit"], rtol=1e-3, atol=1e-3)
    y = emu.dependent_variables
    gp.set_parameter_vector(g["p_fit"])
    lml_at_golden = gp.log_likelihood(y)
    assert abs(-emu.fit_result.fun - lml_at_golden) <= 1e-7 * abs(lml_at_golden)
    # predictions at the golden optimum (so that the comparison is at identical parameters)
    gp.set_parameter_vector(g["p_fit"])
    mu, var = emu.predict_values(g["query_x"], theta_dict(spec, g))
    scale = max(np.max(np.abs(g["mu"])), 1e-12)
    assert np.max(np.abs(mu - g["mu"])) <= 1e-9 * scale
    amp = len(g["p_fit"][1:]) * np.exp(g["p_fit"][0])
    assert np.max(np.abs(var - g["var"])) <= 1e-9 * amp
    assert np.array_equal(emu.predict_values_no_error(g["query_x"], theta_dict(spec, g)), mu)


@pytest.mark.parametrize("label", ["overlap", "three"])
def test_multi_region_emulator_matches_reference_golden(label):
    """The reference's per-region loop (multi_gaussian_process.py:167-177) on the device: every region's GP reaches
    the reference's optimum at the optimiser's resolution; predictions are compared at the reference's optima."""
    g = load("multi_regions")
    spec, params, values = golden_inputs(g)
    emu = MultipleGaussianProcessEmulator(independent_regions=_regions(g, label))
    emu.fit_model(spec, params, values)
    for region, p_ref in zip(emu.emulators, g[label + "_p_fit"]):
        assert np.allclose(region.fit_result.x, p_ref, rtol=1e-3, atol=1e-3)
        region.emulator.set_parameter_vector(p_ref)
        lml_at_golden = region.emulator.log_likelihood(region.dependent_variables)
        assert abs(-region.fit_result.fun - lml_at_golden) <= 1e-7 * abs(lml_at_golden)
    amp = max(len(p) - 1 for p in g[label + "_p_fit"]) * np.exp(np.max(g[label + "_p_fit"][:, 0]))
    assert check_multi_against_reference(g, label, emu, theta_dict(spec, g), 1e-9, 1e-9 * amp) > 0


def test_mocking_matches_reference_golden():
    """The reference's prediction sweeps on the device: one grid prediction (query rows never formed, K* rebuilt inside
    the variance GEMM) against the reference's one-`predict_values`-call-per-row loops, at the reference's optimum."""
    g = load("mocking_small")
    spec, params, values = golden_inputs(g)
    emu = GaussianProcessEmulator()
    emu.fit_model(spec, params, values)
    assert np.allclose(emu.fit_result.x, g["p_fit"], rtol=1e-3, atol=1e-3)
    emu.emulator.set_parameter_vector(g["p_fit"])
    amp = (len(g["p_fit"]) - 1) * np.exp(g["p_fit"][0])
    check_mocking_against_reference(g, emu, spec, 1e-9, 1e-9 * amp)


def test_bins_emulator_matches_reference_golden():
    g = load("bins_small")
    spec, params, values = golden_inputs(g)
    emu = GaussianProcessEmulatorBins()
    emu.fit_model(spec, params, values)
    assert np.allclose(emu.bin_fit_parameters, g["p_fit"], rtol=1e-3, atol=1e-3)
    for gp, p in zip(emu.bin_gaussian_process, g["p_fit"]):
        gp.set_parameter_vector(p)
    mu, var = emu.predict_values(g["centers"], theta_dict(spec, g))
    assert np.max(np.abs(mu - g["mu"])) <= 1e-9 * np.max(np.abs(g["mu"]))
    assert np.max(np.abs(var - g["var"])) <= 1e-9 * np.max(2 * np.exp(g["p_fit"][:, 0]))


def test_one_dim_emulator_matches_reference_golden():
    g = load("one_dim_small")
    spec, params, values = synthetic.make_model(int(g["n_par"]), int(g["sims"]), 1, seed=int(g["seed"]))
    emu = GaussianProcessEmulator1D()
    emu.fit_model(spec, params, values)
    assert np.allclose(emu.emulator.get_parameter_vector(), g["p_fit"], rtol=1e-3, atol=1e-3)
    emu.emulator.set_parameter_vector(g["p_fit"])
    mu, var = emu.predict_values(theta_dict(spec, g))
    assert mu == pytest.approx(float(g["mu"]), rel=1e-9)
    assert abs(var - float(g["var"])) <= 1e-9 * 3 * np.exp(g["p_fit"][0])
    assert emu.predict_values_no_error(theta_dict(spec, g)) == mu


def test_cross_check_matches_reference_golden():
    g = load("cross_check_small")
    spec, params, values = golden_inputs(g)
    cc = CrossCheck()
    cc.build_emulators(spec, params, values)
    total, per_model = cc.get_mean_squared()
    assert np.allclose(cc.cross_fit_parameters, g["p_fit"], rtol=2e-3, atol=2e-3)
    assert total == pytest.approx(float(g["total"]), rel=5e-3)


def test_cross_check_bins_matches_reference_golden():
    """Leave-one-out over per-bin emulators on the device (cross_check_bins.py:94-109): 9 x 4 small GPs, optima at the
    optimiser's resolution, predictions compared at the reference's optima."""
    g = load("cross_check_bins_small")
    spec, params, values = golden_inputs(g)
    cc = CrossCheckBins()
    cc.build_emulators(spec, params, values)
    assert list(cc.leave_out_order) == list(g["order"])
    theta = theta_dict(spec, g)
    for k, uid in enumerate(cc.leave_out_order):
        emulator = cc.cross_emulators[uid]
        got = np.array([gp.get_parameter_vector() for gp in emulator.bin_gaussian_process])
        assert np.allclose(got, g["p_fit"][k], rtol=2e-3, atol=2e-3)
        for gp, p_ref in zip(emulator.bin_gaussian_process, g["p_fit"][k]):
            gp.set_parameter_vector(p_ref)
        mu, var = emulator.predict_values(g["centers"], theta)
        amp = np.max((g["p_fit"].shape[2] - 1) * np.exp(g["p_fit"][k][:, 0]))
        assert np.max(np.abs(mu - g["mu"][k])) <= 1e-9 * np.max(np.abs(g["mu"]))
        assert np.max(np.abs(var - g["var"][k])) <= 1e-9 * amp


def test_reference_smoke_tests_shape():
    """The reference's own test bodies (tests/test_emulator_*.py): fit and predict must run."""
    spec, params, values = fixture_data(21)
    theta = {"x": 0.3, "y": 8.4}
    for cls in (GaussianProcessEmulator,):
        emu = cls(); emu.fit_model(spec, params, values)
        mu, var = emu.predict_values(np.arange(10.0), theta)
        assert mu.shape == (10,) and var.shape == (10,)
    bins = GaussianProcessEmulatorBins(); bins.fit_model(spec, params, values)
    assert bins.predict_values(np.arange(10.0), theta)[0].shape == (10,)
    multi = MultipleGaussianProcessEmulator(independent_regions=[[None, 6], [4, None]])
    multi.fit_model(spec, params, values)
    assert np.all(np.isfinite(multi.predict_values(np.arange(10.0), theta)[0]))
    mcmc = GaussianProcessEmulatorMCMC(mcmc_steps=1, burn_in_steps=1, walkers=10, seed=0)
    mcmc.fit_model(spec, params, values)
    assert mcmc.predict_values(np.arange(10.0), theta)[0].shape == (10,)
    emu = GaussianProcessEmulator(); emu.fit_model(spec, params, values)
    mocked, new_params = mocking.mock_hypercube(emu, spec, samples=256)
    assert len(mocked) == 256


def test_device_gp_caches_value_and_gradient_per_parameter_vector():
    spec, params, values = synthetic.make_model(2, 12, 10, seed=3)
    emu = GaussianProcessEmulator()
    emu.fit_model(spec, params, values)
    gp = emu.emulator
    res = emu.fit_result
    # one device evaluation per distinct parameter vector although scipy calls fun and jac separately
    assert gp.n_device_evaluations <= max(res.nfev, res.njev) + 1
    before = gp.n_device_evaluations
    gp.set_parameter_vector(res.x)
    gp.log_likelihood(emu.dependent_variables)
    gp.grad_log_likelihood(emu.dependent_variables)
    emu.predict_values(np.linspace(0, 1, 5), {"p0": 0.5, "p1": 0.5})
    assert gp.n_device_evaluations - before <= 1


def test_sobol_sweep_on_device():
    spec, params, values = synthetic.make_model(3, 40, 8, seed=12)
    emu = GaussianProcessEmulator()
    emu.fit_model(spec, params, values)
    out = sobol_indices(emu, spec, np.linspace(0.1, 0.9, 4), base_samples=256)
    assert out["S1"].shape == (3, 4) and np.all(np.isfinite(out["ST"]))
    assert np.all(out["ST"] > -0.05)


def test_device_handles_are_released_and_recreated_within_the_memory_budget():
    """CrossCheck keeps one fitted GP per simulation (512 at BASELINE C4, 1.07 GB of device state
    each): handles are created lazily and the least recently used ones released when memory runs
    short.  Shrink the budget so that two handles do not fit and check results survive eviction."""
    from emulator_b200 import george_compat as george
    from emulator_b200.george_compat import gp as gp_module
    from emulator_b200 import synthetic

    res = gp_module._residency
    n, d = 300, 3
    gps, ys, outs = [], [], []
    old_reserve = res.RESERVE
    try:
        for k in range(3):
            x, y, yerr = synthetic.design_arrays(n, d, seed=k)
            g = george.GP(1**2 * george.kernels.ExpSquaredKernel(np.ones(d), ndim=d))
            g.compute(x, yerr)
            assert not g.resident  # lazy: no device memory until first use
            gps.append(g); ys.append(y)
        # budget: nothing fits next to what is resident, so every new handle evicts the others
        res.RESERVE = gp_module._Residency.free_bytes(0)
        for g, y in zip(gps, ys):
            outs.append((g.log_likelihood(y), g.grad_log_likelihood(y), g.predict(y, g._x[:5] + 0.01, return_cov=False, return_var=True)))
        assert sum(g.resident for g in gps) == 1 and gps[2].resident
        # an evicted GP answers from its cache without touching the device ...
        evals = gps[0].n_device_evaluations
        assert gps[0].log_likelihood(ys[0]) == outs[0][0] and not gps[0].resident
        assert gps[0].n_device_evaluations == evals
        # ... and re-creates its handle (one re-factorisation) for a prediction, bit-identical
        mu, var = gps[0].predict(ys[0], gps[0]._x[:5] + 0.01, return_cov=False, return_var=True)
        assert gps[0].resident and not gps[2].resident
        assert np.array_equal(mu, outs[0][2][0]) and np.array_equal(var, outs[0][2][1])
        assert np.array_equal(gps[0].grad_log_likelihood(ys[0]), outs[0][1])
    finally:
        res.RESERVE = old_reserve
        for g in gps:
            g.close()


def test_bins_lockstep_fit_reaches_the_optima_of_the_scipy_fit():
    """fit_mode="lockstep": all bins advance together (one batched device call per round).  The line
    search is not scipy's, so iterates differ; the optima and the predictions must agree."""
    from emulator_b200 import synthetic

    spec, params, values = synthetic.make_config("C3", seed_offset=3)
    # a small slice of C3: 6 bins, 48 simulations
    keep = list(values.model_values.keys())[:48]
    from emulator_b200.backend import ModelParameters, ModelValues
    vals = ModelValues({k: {kk: np.asarray(vv)[:6] for kk, vv in values.model_values[k].items()} for k in keep})
    pars = ModelParameters({k: params.model_parameters[k] for k in keep})
    a = GaussianProcessEmulatorBins(fit_mode="scipy", workers=1)
    a.fit_model(spec, pars, vals)
    b = GaussianProcessEmulatorBins(fit_mode="lockstep")
    b.fit_model(spec, pars, vals)
    assert b.lockstep_result is not None and b.lockstep_result.rounds < 400
    theta = pars.model_parameters[keep[5]]
    ind = np.asarray(vals.model_values[keep[0]]["independent"])
    for gp_a, gp_b, data in zip(a.bin_gaussian_process, b.bin_gaussian_process, a.bin_model_values):
        la, lb = gp_a.log_likelihood(data["dependent"]), gp_b.log_likelihood(data["dependent"])
        assert abs(la - lb) <= 1e-5 * max(1.0, abs(la))  # same optimum of the likelihood
    mu_a, var_a = a.predict_values(ind, theta)
    mu_b, var_b = b.predict_values(ind, theta)
    assert np.allclose(mu_a, mu_b, rtol=1e-3, atol=1e-5) and np.allclose(var_a, var_b, rtol=5e-2, atol=1e-8)
    with pytest.raises(ValueError):
        GaussianProcessEmulatorBins(fit_mode="other").fit_model(spec, pars, vals)


def test_cross_check_lockstep_on_the_device():
    from emulator_b200 import synthetic
    from emulator_b200.sensitivity import CrossCheck

    spec, params, values = synthetic.make_model(2, 10, 6, seed=4)
    a = CrossCheck(fit_mode="scipy", workers=1)
    a.build_emulators(spec, params, values)
    b = CrossCheck(fit_mode="lockstep")
    b.build_emulators(spec, params, values)
    for uid in a.cross_emulators:
        ea, eb = a.cross_emulators[uid], b.cross_emulators[uid]
        la = ea.emulator.log_likelihood(ea.dependent_variables)
        lb = eb.emulator.log_likelihood(eb.dependent_variables)
        assert abs(la - lb) <= 1e-5 * max(1.0, abs(la))
    assert abs(a.get_mean_squared()[0] - b.get_mean_squared()[0]) <= 1e-3 * abs(a.get_mean_squared()[0])
