"""
Host-side logic on CPU: the emulators of emulator_b200 driven through the oracle-backed GP
(test-only substitution, see conftest.oracle_backend) must reproduce what the REFERENCE's
emulator code produced on the same seeded inputs (tests/golden, made by
tests/golden/make_golden.py from /root/reference).
"""
import os

import numpy as np
import pytest

from emulator_b200 import mocking, synthetic
from emulator_b200.backend import ModelParameters, ModelSpecification, ModelValues
from emulator_b200.emulators import (
    GaussianProcessEmulator, GaussianProcessEmulator1D, GaussianProcessEmulatorBins, GaussianProcessEmulatorMCMC,
    MultipleGaussianProcessEmulator,
)
from emulator_b200.mean_models import LinearMeanModel, OffsetMeanModel, PolynomialMeanModel, FixedMeanModel
from emulator_b200.sensitivity import CrossCheck, CrossCheckBins, saltelli_design, sobol_indices

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def fixture_data(seed):
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


def ragged_data(seed):
    """Simulations with different numbers of points at different independent values, one without errors (the same
    construction as tests/golden/make_golden.py::ragged_fixture)."""
    rng = np.random.default_rng(seed)
    spec = ModelSpecification(number_of_parameters=3, parameter_names=["a", "b", "c"],
                              parameter_limits=[[0.0, 1.0], [-1.0, 1.0], [2.0, 4.0]])
    params, values = {}, {}
    for uid, n in enumerate([4, 9, 6, 11, 7, 1, 8, 5, 10, 3, 12, 6, 9, 2, 7, 8]):
        theta = {"a": float(rng.random()), "b": float(2 * rng.random() - 1), "c": float(2 + 2 * rng.random())}
        params[f"run{uid}"] = theta
        x = np.sort(2.0 * rng.random(n))
        y = np.sin(1.5 * x + 2.0 * theta["a"]) * theta["c"] + theta["b"] * x * x + 0.01 * rng.standard_normal(n)
        entry = {"independent": x, "dependent": y}
        if uid != 5:  # (the single-point simulation: an error-less one with close points would put cond(K) at 1e13)
            entry["dependent_error"] = 0.02 + 0.02 * rng.random(n)
        values[f"run{uid}"] = entry
    return spec, ModelParameters(model_parameters=params), ModelValues(model_values=values)


def golden_inputs(g):
    if "fixture" in g.files and int(g["fixture"]) == 2:
        return ragged_data(int(g["seed"]))
    if "fixture" in g.files and int(g["fixture"]):
        return fixture_data(int(g["seed"]))
    return synthetic.make_model(int(g["n_par"]), int(g["sims"]), int(g["bins"]), seed=int(g["seed"]))


def theta_dict(spec, g):
    return {name: float(v) for name, v in zip(spec.parameter_names, g["theta"])}


@pytest.mark.parametrize(
    "name,mean_model",
    [("gpe_small", None), ("gpe_noerr", None), ("gpe_ragged", None), ("gpe_linear_mean", LinearMeanModel()), ("gpe_offset_mean", OffsetMeanModel()),
     ("gpe_polynomial_mean", PolynomialMeanModel(degree=2))],
)
def test_gpe_matches_reference_flow(oracle_backend, name, mean_model):
    g = load(name)
    spec, params, values = golden_inputs(g)
    emu = GaussianProcessEmulator(mean_model=mean_model)
    emu.fit_model(spec, params, values)
    # data flattening + float32 rounding: bit exact
    assert emu.independent_variables.dtype == np.float32
    assert np.array_equal(emu.independent_variables, g["x"])
    assert np.array_equal(emu.dependent_variables, g["y"])
    assert np.array_equal(emu.dependent_variable_errors, g["yerr"])
    # same oracle arithmetic + same scipy BFGS => same optimum
    assert np.allclose(emu.emulator.get_parameter_vector(), g["p_fit"], rtol=1e-9, atol=1e-9)
    mu, var = emu.predict_values(g["query_x"], theta_dict(spec, g))
    assert np.allclose(mu, g["mu"], rtol=1e-8, atol=1e-10)
    assert np.allclose(var, g["var"], rtol=1e-6, atol=1e-10)
    assert np.allclose(emu.predict_values_no_error(g["query_x"], theta_dict(spec, g)), g["mu_only"], rtol=1e-8, atol=1e-10)


def test_bins_matches_reference_flow(oracle_backend):
    g = load("bins_small")
    spec, params, values = golden_inputs(g)
    emu = GaussianProcessEmulatorBins()
    emu.fit_model(spec, params, values)
    assert np.array_equal(np.array(emu.bin_centers), g["centers"])
    assert emu.bin_model_values[0]["independent"].dtype == np.float64
    assert np.array_equal(emu.bin_model_values[0]["independent"], g["bin0_x"])
    assert np.array_equal(emu.bin_model_values[0]["dependent"], g["bin0_y"])
    assert np.array_equal(emu.bin_model_values[0]["dependent_errors"], g["bin0_err"])
    assert np.allclose(emu.bin_fit_parameters, g["p_fit"], rtol=1e-8, atol=1e-8)
    mu, var = emu.predict_values(g["centers"], theta_dict(spec, g))
    assert np.allclose(mu, g["mu"], rtol=1e-8, atol=1e-10)
    assert np.allclose(var, g["var"], rtol=1e-6, atol=1e-10)
    with pytest.raises(AttributeError):
        emu.predict_values(np.array([0.123456]), theta_dict(spec, g))


def test_one_dim_matches_reference_flow(oracle_backend):
    g = load("one_dim_small")
    spec, params, values = synthetic.make_model(int(g["n_par"]), int(g["sims"]), 1, seed=int(g["seed"]))
    emu = GaussianProcessEmulator1D()
    emu.fit_model(spec, params, values)
    assert np.array_equal(emu.independent_variables, g["x"])
    assert np.allclose(emu.emulator.get_parameter_vector(), g["p_fit"], rtol=1e-9, atol=1e-9)
    mu, var = emu.predict_values(theta_dict(spec, g))
    assert mu == pytest.approx(float(g["mu"]), rel=1e-8)
    assert var == pytest.approx(float(g["var"]), rel=1e-6, abs=1e-10)


def test_cross_check_matches_reference_flow(oracle_backend):
    g = load("cross_check_small")
    spec, params, values = golden_inputs(g)
    keys_before = list(values.model_values.keys())
    cc = CrossCheck()
    cc.build_emulators(spec, params, values)
    assert list(values.model_values.keys()) == keys_before  # caller's container untouched
    total, per_model = cc.get_mean_squared()
    # the reference rotates the row order per refit (pop / re-insert); same maths, different
    # rounding path, so optima agree to optimiser tolerance rather than to the last bit
    assert np.allclose(cc.cross_fit_parameters, g["p_fit"], rtol=2e-3, atol=2e-3)
    assert total == pytest.approx(float(g["total"]), rel=5e-3)
    mocked = cc.build_mocked_model_values_original_independent()
    assert set(mocked.keys()) == set(values.keys())
    assert np.all(mocked[0]["dependent_error"] >= 0)
    at = cc.build_mocked_model_values(np.linspace(0, 1, 4))
    assert len(at[0]["dependent"]) == 4


def test_untrained_emulators_raise(oracle_backend):
    for emu in (GaussianProcessEmulator(), GaussianProcessEmulatorBins(), GaussianProcessEmulatorMCMC()):
        with pytest.raises(AttributeError):
            emu.predict_values(np.array([0.0]), {"x": 0.0, "y": 8.0})
    with pytest.raises(AttributeError):
        GaussianProcessEmulator1D().predict_values({"x": 0.0})
    with pytest.raises(AttributeError):
        CrossCheck().get_mean_squared()


def test_multi_region_blends_linearly(oracle_backend):
    spec, params, values = fixture_data(5)
    emu = MultipleGaussianProcessEmulator(independent_regions=[[None, 6], [4, None]])
    emu.fit_model(spec, params, values)
    x = np.arange(10, dtype=np.float64)
    theta = {"x": 0.4, "y": 8.5}
    mu, err = emu.predict_values(x, theta)
    mu_only = emu.predict_values_no_error(x, theta)
    assert np.allclose(mu, mu_only)
    left, _ = emu.emulators[0].predict_values(np.array([5.0]), theta)
    right, _ = emu.emulators[1].predict_values(np.array([5.0]), theta)
    assert mu[5] == pytest.approx(0.5 * left[0] + 0.5 * right[0])
    only_left, _ = emu.emulators[0].predict_values(np.array([2.0]), theta)
    assert mu[2] == pytest.approx(only_left[0])
    assert np.all(np.isfinite(err))


def _regions(g, label):
    return [[None if np.isnan(v) else float(v) for v in row] for row in g[label + "_regions"]]


def check_multi_against_reference(g, label, emu, theta, mean_tol, err_tol):
    """Predictions of a fitted multi-region emulator against what the reference's own code produced (golden
    `multi_regions`): means everywhere; errors wherever one region answers; inside an overlap the documented
    weighted square sum of BOTH regions' errors, formed here from the reference's per-region outputs (the
    reference itself reads the right-hand error from the left emulator there, multi_gaussian_process.py:312-314)."""
    q = g[label + "_query"]
    mu, err = emu.predict_values(q, theta)
    scale = np.max(np.abs(g[label + "_mu"]))
    assert np.max(np.abs(mu - g[label + "_mu"])) <= mean_tol * scale
    assert np.max(np.abs(emu.predict_values_no_error(q, theta) - g[label + "_mu"])) <= mean_tol * scale
    regions = _regions(g, label)
    covered = np.array([~np.isnan(g[f"{label}_region{i}_mu"]) for i in range(len(regions))])
    single = covered.sum(axis=0) == 1
    assert single.any()
    assert np.max(np.abs(err[single] - g[label + "_err"][single])) <= err_tol
    for i in np.where(~single)[0]:
        left, right = np.where(covered[:, i])[0]
        low, high = regions[right][0], regions[left][1]
        wl, wr = (high - q[i]) / (high - low), (q[i] - low) / (high - low)
        el, er = g[f"{label}_region{left}_err"][i], g[f"{label}_region{right}_err"][i]
        assert err[i] == pytest.approx(np.sqrt(wl * el * el + wr * er * er), abs=err_tol)
    return int((~single).sum())


@pytest.mark.parametrize("label", ["overlap", "touch", "three"])
def test_multi_region_matches_reference_flow(oracle_backend, label):
    g = load("multi_regions")
    spec, params, values = golden_inputs(g)
    emu = MultipleGaussianProcessEmulator(independent_regions=_regions(g, label))
    emu.fit_model(spec, params, values)
    assert [len(e.dependent_variables) for e in emu.emulators] == list(g[label + "_rows"])
    got = np.array([e.emulator.get_parameter_vector() for e in emu.emulators])
    assert np.allclose(got, g[label + "_p_fit"], rtol=1e-9, atol=1e-9)
    blended = check_multi_against_reference(g, label, emu, theta_dict(spec, g), 1e-9, 1e-10)
    assert (blended > 0) == (label != "touch")


def check_mocking_against_reference(g, emu, spec, mean_tol, err_tol):
    """`mock_hypercube` / `mock_sweep` over a fitted emulator against what the reference's own loops produced
    (golden `mocking_small`; mocking/__init__.py:84-95, :187-198): same unique identifiers, same parameter rows (the
    reference's cube comes from numpy's global generator, seeded the same here), same predictions."""
    np.random.seed(int(g["seed"]))
    cube_values, cube_parameters = mocking.mock_hypercube(emu, spec, 9)
    center = {name: float(v) for name, v in zip(spec.parameter_names, g["center"])}
    sweep_values, sweep_parameters = mocking.mock_sweep(emu, spec, 7, str(g["swept"]), center)
    for label, vals, pars in (("cube", cube_values, cube_parameters), ("sweep", sweep_values, sweep_parameters)):
        uids = list(pars.model_parameters.keys())
        assert uids == [str(u) for u in g[label + "_uids"]]
        theta = np.array([[pars.model_parameters[u][name] for name in spec.parameter_names] for u in uids])
        assert np.array_equal(theta, g[label + "_theta"])
        mu = np.array([vals.model_values[u]["dependent"] for u in uids])
        err = np.array([vals.model_values[u]["dependent_error"] for u in uids])
        assert np.array_equal(vals.model_values[uids[0]]["independent"], g["independent"])
        assert np.max(np.abs(mu - g[label + "_mu"])) <= mean_tol * np.max(np.abs(g[label + "_mu"]))
        assert np.max(np.abs(err - g[label + "_err"])) <= err_tol


def test_mocking_matches_reference_flow(oracle_backend):
    g = load("mocking_small")
    spec, params, values = golden_inputs(g)
    emu = GaussianProcessEmulator()
    emu.fit_model(spec, params, values)
    assert np.allclose(emu.emulator.get_parameter_vector(), g["p_fit"], rtol=1e-9, atol=1e-9)
    check_mocking_against_reference(g, emu, spec, 1e-8, 1e-9)


def test_mcmc_emulator_runs_and_is_seeded(oracle_backend):
    spec, params, values = fixture_data(6)
    a = GaussianProcessEmulatorMCMC(burn_in_steps=2, mcmc_steps=3, walkers=10, seed=4)
    a.fit_model(spec, params, values)
    b = GaussianProcessEmulatorMCMC(burn_in_steps=2, mcmc_steps=3, walkers=10, seed=4)
    b.fit_model(spec, params, values)
    assert a.hyperparameter_samples.shape == (3 * 10, 4)
    assert np.array_equal(a.hyperparameter_best_fit, b.hyperparameter_best_fit)
    mu, var = a.predict_values(np.arange(10.0), {"x": 0.3, "y": 8.2})
    assert mu.shape == (10,) and np.all(np.isfinite(var))
    c = GaussianProcessEmulatorMCMC(burn_in_steps=1, mcmc_steps=2, walkers=10, seed=1, use_hyperparameter_error=True,
                                    samples_for_error=5)
    c.fit_model(spec, params, values)
    mu2, var2 = c.predict_values(np.arange(10.0), {"x": 0.3, "y": 8.2})
    assert np.all(var2 >= 0) or np.all(np.isfinite(var2))
    d = GaussianProcessEmulatorMCMC(burn_in_steps=1, mcmc_steps=1, walkers=10, seed=1, use_hyperparameter_error=True,
                                    samples_for_error=500)
    d.fit_model(spec, params, values)
    with pytest.raises(ValueError):
        d.predict_values(np.arange(10.0), {"x": 0.3, "y": 8.2})


def test_mocking_batched_equals_per_point(oracle_backend):
    spec, params, values = synthetic.make_model(2, 10, 6, seed=9)
    emu = GaussianProcessEmulator()
    emu.fit_model(spec, params, values)
    mocked, new_params = mocking.mock_hypercube(emu, spec, samples=5, rng=np.random.default_rng(0))
    assert len(mocked) == 5 and all(k.startswith("emulated_") for k in mocked.keys())
    uid = "emulated_3"
    mu, var = emu.predict_values(mocked[uid]["independent"], new_params[uid])
    assert np.allclose(mocked[uid]["dependent"], mu, rtol=1e-12, atol=1e-13)
    assert np.allclose(mocked[uid]["dependent_error"], var, rtol=1e-9, atol=1e-12)
    sweep, sweep_params = mocking.mock_sweep(emu, spec, 4, "p1", {"p0": 0.5, "p1": 0.5})
    assert [sweep_params[f"emulated_{i}"]["p1"] for i in range(4)] == list(np.linspace(0, 1, 4))
    assert len(sweep["emulated_0"]["dependent"]) == 6


def test_sobol_design_and_indices(oracle_backend):
    spec, params, values = synthetic.make_model(2, 14, 5, seed=12)
    design = saltelli_design(spec, 64, seed=3)
    assert design.shape == (64 * (2 + 2), 2)
    assert np.array_equal(design[2 * 64 : 3 * 64, 1], design[:64, 1])      # AB_0 keeps A's column 1
    assert np.array_equal(design[2 * 64 : 3 * 64, 0], design[64:128, 0])   # ... and takes B's column 0
    emu = GaussianProcessEmulator()
    emu.fit_model(spec, params, values)
    out = sobol_indices(emu, spec, np.array([0.25, 0.75]), base_samples=128)
    assert out["S1"].shape == (2, 2) and np.all(out["ST"] > -0.05)


def test_mean_models_protocol():
    x = np.linspace(0, 1, 20).reshape(-1, 1).astype(np.float32)
    y = (3 * x[:, 0] + 1).astype(np.float32)
    off = OffsetMeanModel(); off.train(x, y)
    assert off.predict(x).dtype == np.float32 and np.allclose(off.predict(x), y.mean())
    lin = LinearMeanModel(); lin.train(x, y)
    frozen = lin.george_model
    lin.train(x, -y)  # retraining must not change the frozen copy
    assert np.allclose(frozen.get_value(x), y, atol=1e-5)
    poly = PolynomialMeanModel(degree=2); poly.train(x, y)
    assert np.allclose(poly.predict(x), y, atol=1e-4)
    assert np.allclose(FixedMeanModel(model=2.0).predict(x), 2.0)


def test_containers_validate_and_roundtrip(tmp_path):
    with pytest.raises(AttributeError):
        ModelSpecification(number_of_parameters=2, parameter_names=["a"], parameter_limits=[[0, 1]])
    with pytest.raises(AttributeError):
        ModelParameters(model_parameters={0: {"a": 1.0}, 1: {"b": 2.0}})
    with pytest.raises(AttributeError):
        ModelValues(model_values={0: {"independent": np.zeros(2), "dependent": np.zeros(2), "bad": np.zeros(2)}})
    with pytest.raises(AttributeError):
        ModelValues(model_values={0: {"independent": np.zeros(2), "dependent": np.zeros(2), "dependent_error": np.zeros(3)}})
    spec, params, values = synthetic.make_model(2, 4, 3, seed=1)
    assert values.number_of_variables == 12 and spec.salib_problem["num_vars"] == 2
    values.to_json(tmp_path / "v.json"); back = ModelValues.from_json(tmp_path / "v.json")
    assert np.allclose(back["0"]["dependent"], values[0]["dependent"])
    params.to_yaml(tmp_path / "p.yaml"); pback = ModelParameters.from_yaml(tmp_path / "p.yaml")
    assert pback[0]["p0"] == pytest.approx(params[0]["p0"])
    closest, _ = params.find_closest_model(params[2], 1)
    assert closest == [2]


def test_cross_check_bins_builds_and_scores(oracle_backend):
    """Reference test_cross_check_bins.py only builds; the broken post-processing there
    (cross_check_bins.py:142-148, :224-229) is implemented to its evident intent here."""
    spec, params, values = synthetic.make_model(2, 7, 4, seed=77)
    cc = CrossCheckBins()
    cc.build_emulators(spec, params, values)
    assert list(cc.cross_emulators.keys()) == list(values.keys())
    assert all(e.n_bins == 4 for e in cc.cross_emulators.values())
    total, per_model = cc.get_mean_squared()
    assert np.isfinite(total) and set(per_model) == set(values.keys())
    with pytest.raises(AttributeError):
        CrossCheckBins().get_mean_squared()


def test_cross_check_bins_matches_reference_flow(oracle_backend):
    """The fits of the reference's CrossCheckBins.build_emulators (cross_check_bins.py:94-109; golden
    `cross_check_bins_small`) and each leave-one-out emulator's predictions at the bin centres."""
    g = load("cross_check_bins_small")
    spec, params, values = golden_inputs(g)
    cc = CrossCheckBins()
    cc.build_emulators(spec, params, values)
    assert list(cc.leave_out_order) == list(g["order"])
    got = np.array([[gp.get_parameter_vector() for gp in cc.cross_emulators[uid].bin_gaussian_process]
                    for uid in cc.leave_out_order])
    assert np.allclose(got, g["p_fit"], rtol=1e-8, atol=1e-8)
    theta = theta_dict(spec, g)
    pred = np.array([cc.cross_emulators[uid].predict_values(g["centers"], theta) for uid in cc.leave_out_order])
    assert np.allclose(pred[:, 0], g["mu"], rtol=1e-9, atol=1e-11)
    assert np.allclose(pred[:, 1], g["var"], rtol=1e-7, atol=1e-12)


def test_ensemble_sampler_recovers_a_gaussian():
    from emulator_b200.emulators.ensemble import EnsembleSampler

    mean, sigma = np.array([1.0, -2.0]), np.array([0.5, 2.0])
    sampler = EnsembleSampler(20, 2, lambda p: -0.5 * np.sum(((p - mean) / sigma) ** 2), seed=3)
    p0 = mean + 0.1 * np.random.default_rng(0).standard_normal((20, 2))
    p0, _, _ = sampler.run_mcmc(p0, 300)
    sampler.run_mcmc(p0, 700)
    chain = sampler.get_chain()
    assert chain.shape == (1000, 20, 2)
    flat = sampler.get_chain(flat=True, discard=300)
    assert np.allclose(flat.mean(axis=0), mean, atol=0.25)
    assert np.allclose(flat.std(axis=0), sigma, rtol=0.25)
    with pytest.raises(ValueError):
        EnsembleSampler(7, 2, lambda p: 0.0)


def test_vectorized_ensemble_sampler_walks_the_same_chain():
    """`vectorize=True` (one call per half-ensemble, what the batched device path uses) must not
    change the chain: same proposals, same acceptances."""
    from emulator_b200.emulators.ensemble import EnsembleSampler

    mean = np.array([0.3, -1.0, 2.0])
    scalar = EnsembleSampler(12, 3, lambda p: -0.5 * np.sum((p - mean) ** 2), seed=5)
    blocks = []

    def block_fn(ps):
        blocks.append(len(ps))
        return -0.5 * np.sum((ps - mean) ** 2, axis=1)

    vector = EnsembleSampler(12, 3, block_fn, seed=5, vectorize=True)
    p0 = mean + 0.1 * np.random.default_rng(1).standard_normal((12, 3))
    a = scalar.run_mcmc(p0, 25)
    b = vector.run_mcmc(p0, 25)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(scalar.get_chain(), vector.get_chain())
    assert blocks[0] == 12 and set(blocks[1:]) == {6}  # initial ensemble, then half-ensembles
    with pytest.raises(ValueError):
        EnsembleSampler(4, 2, lambda ps: np.zeros(3), vectorize=True).run_mcmc(np.zeros((4, 2)), 1)


def test_gp_handle_residency_logic_with_a_fake_device(monkeypatch):
    """Pin / evict / re-create logic of george_compat.GP, exercised without a GPU: the device problem is
    replaced by a fake that counts its live instances and answers with the oracle."""
    from emulator_b200 import george_compat as george
    from emulator_b200.george_compat import gp as gp_module
    from oracle import gp_oracle

    live = {"count": 0, "created": 0}

    class FakeProblem:
        def __init__(self, x, noise, device=0):
            self.x, self.noise, self.r = x, noise, None
            live["count"] += 1
            live["created"] += 1

        def close(self):
            live["count"] -= 1

        def set_residual(self, r):
            self.r = r.copy()

        def evaluate(self, p, want_grad=True, value_only=False):
            lml, grad, logdet, quad = gp_oracle.lml_and_grad(p, self.x, self.noise, self.r)
            return lml, (grad if want_grad else np.zeros(len(p))), 0, logdet, quad

        def predict(self, t, want_var=True):
            raise AssertionError("not used here")

    monkeypatch.setattr(gp_module, "DeviceProblem", FakeProblem)
    monkeypatch.setattr(gp_module._lib, "require_gpu", lambda: None)
    budget = {"free": 10}
    monkeypatch.setattr(gp_module._Residency, "free_bytes", staticmethod(lambda device: budget["free"] - live["count"] * 4))
    monkeypatch.setattr(gp_module._Residency, "needed_bytes", staticmethod(lambda n: 4))
    monkeypatch.setattr(gp_module._residency, "RESERVE", 0)

    gps, ys = [], []
    for k in range(4):
        x, y, yerr = synthetic.design_arrays(30, 2, seed=k)
        g = george.GP(1**2 * george.kernels.ExpSquaredKernel(np.ones(2), ndim=2))
        g.compute(x, yerr)
        gps.append(g); ys.append(y)
    assert live["count"] == 0  # lazy
    values = [g.log_likelihood(y) for g, y in zip(gps, ys)]
    # the budget (10) holds two handles of 4 "bytes": the two least recently used ones were released
    assert live["count"] == 2 and [g.resident for g in gps] == [False, False, True, True]
    # cached values are served without a handle
    assert gps[0].log_likelihood(ys[0]) == values[0] and not gps[0].resident and live["created"] == 4
    # a new parameter vector needs the device again: the LRU handle (gps[2]) goes
    gps[0].set_parameter_vector(gps[0].get_parameter_vector() + 0.1)
    assert np.isfinite(gps[0].log_likelihood(ys[0]))
    assert [g.resident for g in gps] == [True, False, False, True] and live["created"] == 5
    # a pinned handle is never evicted, and cannot be released by hand
    with gps[3]._using_device():
        gps[1].set_parameter_vector(gps[1].get_parameter_vector() - 0.1)
        gps[1].log_likelihood(ys[1])
        assert gps[3].resident and not gps[0].resident
        with pytest.raises(RuntimeError):
            gps[3].release_device()
    for g in gps:
        g.close()
    assert live["count"] == 0


def test_lockstep_bfgs_minimises_many_problems_together():
    """The vectorised BFGS of emulator_b200.optimise against known minima and scipy's BFGS."""
    from scipy.optimize import minimize
    from emulator_b200.optimise import minimise_lockstep

    rng = np.random.default_rng(0)
    nb, P = 12, 5
    mats = []
    for _ in range(nb):
        m = rng.standard_normal((P, P))
        mats.append(m @ m.T + np.eye(P))
    centres = rng.standard_normal((nb, P))
    calls = []

    def quartic(indices, points):
        calls.append(len(indices))
        f, g = [], []
        for i, x in zip(indices, points):
            z = x - centres[i]
            f.append(0.5 * z @ mats[i] @ z + 0.1 * np.sum(z ** 4))
            g.append(mats[i] @ z + 0.4 * z ** 3)
        return np.array(f), np.array(g)

    res = minimise_lockstep(quartic, np.zeros((nb, P)))
    assert np.all(res.status == 0) and np.max(np.abs(res.x - centres)) < 1e-4
    assert res.rounds == len(calls) and calls[0] == nb and min(calls) >= 1  # finished problems drop out
    assert np.all(np.max(np.abs(res.jac), axis=1) < 1e-5)

    def rosenbrock(indices, points):
        f, g = [], []
        for x in points:
            f.append(np.sum(100.0 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2))
            gg = np.zeros_like(x)
            gg[:-1] += -400.0 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2.0 * (1 - x[:-1])
            gg[1:] += 200.0 * (x[1:] - x[:-1] ** 2)
            g.append(gg)
        return np.array(f), np.array(g)

    x0 = 0.5 * rng.standard_normal((6, 4))
    res = minimise_lockstep(rosenbrock, x0)
    assert np.all(res.status == 0) and np.max(np.abs(res.x - 1.0)) < 1e-5
    ref = minimize(lambda x: rosenbrock([0], [x])[0][0], x0[0], jac=lambda x: rosenbrock([0], [x])[1][0], method="BFGS")
    # (the 4-d Rosenbrock function has a second, local minimum that scipy's line search may fall into)
    assert res.fun[0] <= ref.fun + 1e-10 and np.all(res.nfev < 200)

    # a failed evaluation (non-finite value) is a step that was too long, not an error
    def walled(indices, points):
        f, g = quartic(indices, points)
        return np.where(np.max(np.abs(points), axis=1) > 6.0, np.inf, f), g

    res = minimise_lockstep(walled, np.zeros((nb, P)))
    assert np.all(res.status == 0) and np.max(np.abs(res.x - centres)) < 1e-4
    with pytest.raises(ValueError):
        minimise_lockstep(walled, np.full((nb, P), 7.0))


def test_bins_lockstep_fit_matches_the_scipy_fit(oracle_backend):
    """Host logic of fit_mode="lockstep" (all bins advance together) against the per-bin scipy fits."""
    spec, params, values = synthetic.make_model(2, 14, 4, seed=6)
    a = GaussianProcessEmulatorBins(fit_mode="scipy", workers=1)
    a.fit_model(spec, params, values)
    b = GaussianProcessEmulatorBins(fit_mode="lockstep")
    b.fit_model(spec, params, values)
    assert b.lockstep_result.x.shape == (4, 3) and np.all(b.lockstep_result.status >= 0)
    for ga, gb, data in zip(a.bin_gaussian_process, b.bin_gaussian_process, a.bin_model_values):
        la, lb = ga.log_likelihood(data["dependent"]), gb.log_likelihood(data["dependent"])
        assert abs(la - lb) <= 1e-5 * max(1.0, abs(la))
    theta = params.model_parameters[list(values.model_values.keys())[3]]
    ind = np.asarray(values.model_values[list(values.model_values.keys())[0]]["independent"])
    assert np.allclose(a.predict_values(ind, theta)[0], b.predict_values(ind, theta)[0], rtol=1e-3, atol=1e-6)


def test_cross_check_lockstep_matches_the_scipy_refits(oracle_backend):
    """CrossCheck(fit_mode="lockstep"): the leave-one-out refits advance together; same optima."""
    spec, params, values = synthetic.make_model(2, 8, 5, seed=9)
    a = CrossCheck(fit_mode="scipy", workers=1)
    a.build_emulators(spec, params, values)
    b = CrossCheck(fit_mode="lockstep")
    b.build_emulators(spec, params, values)
    assert b.lockstep_result is not None and b.cross_fit_parameters.shape == a.cross_fit_parameters.shape
    for uid in a.cross_emulators:
        ea, eb = a.cross_emulators[uid], b.cross_emulators[uid]
        la = ea.emulator.log_likelihood(ea.dependent_variables)
        lb = eb.emulator.log_likelihood(eb.dependent_variables)
        assert abs(la - lb) <= 1e-5 * max(1.0, abs(la))
    ms_a, _ = a.get_mean_squared()
    ms_b, _ = b.get_mean_squared()
    assert abs(ms_a - ms_b) <= 1e-3 * abs(ms_a)
    with pytest.raises(ValueError):
        CrossCheck(fit_mode="nope").build_emulators(spec, params, values)


def test_mocking_a_binned_emulator_goes_through_its_grid_prediction(oracle_backend):
    """Sweeps over a GaussianProcessEmulatorBins: one batched call per bin (predict_grid), same values as
    the per-point predict_values path."""
    spec, params, values = synthetic.make_model(2, 12, 5, seed=2)
    emu = GaussianProcessEmulatorBins(workers=1)
    emu.fit_model(spec, params, values)
    mocked, new_params = mocking.mock_hypercube(emu, spec, samples=7, rng=np.random.default_rng(1))
    assert len(mocked) == 7
    for uid in ("emulated_0", "emulated_6"):
        mu, var = emu.predict_values(mocked[uid]["independent"], new_params[uid])
        assert np.allclose(mocked[uid]["dependent"], mu, rtol=1e-12, atol=1e-13)
        assert np.allclose(mocked[uid]["dependent_error"], var, rtol=1e-9, atol=1e-12)
    theta = np.array([[new_params[f"emulated_{k}"][name] for name in emu.parameter_order] for k in range(7)])
    centers = np.array(emu.bin_centers)
    assert np.array_equal(emu.predict_grid(theta, centers[:2], want_var=False), emu.predict_grid(theta, centers[:2])[0])
    with pytest.raises(AttributeError):
        emu.predict_grid(theta, np.array([0.123456]))
