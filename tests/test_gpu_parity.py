"""
Parity of the CUDA path against the oracle, called through the C ABI (ctypes).

Tolerances (fp64; BASELINE.json north_star asks for 1e-9 relative):
  * LML, logdet, quadratic form: 1e-9 relative
  * gradient: 1e-9 of max|grad| (the survey measured 5e-10 for the CPU path's own
    row-permutation noise at cond(K) ~ 1e9)
  * predictive mean: 1e-9 of max|mean|; predictive variance: 1e-9 of the prior variance
    `amp` (the variance is amp minus a quadratic form of size ~amp; relative-to-itself
    accuracy of a near-zero variance is limited by cond(K) for george as well)
"""
import os

import numpy as np
import pytest

from emulator_b200 import synthetic
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-9


def problem(n, d, seed=0, err=0.02):
    x, y, yerr = synthetic.design_arrays(n, d, seed=seed, error_fraction=err)
    return x.astype(np.float64), y.astype(np.float64), orc.noise_diagonal(yerr)


def device_problem(x, diag, y):
    from emulator_b200.device import DeviceProblem

    gp = DeviceProblem(x, diag)
    gp.set_residual(y)
    return gp


def relmax(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("n,d", [(1, 1), (2, 2), (63, 3), (64, 2), (65, 4), (127, 5), (128, 6), (129, 3), (300, 8), (640, 9),
                                 (1000, 3), (1536, 10), (777, 12), (200, 16), (96, 7), (500, 1)])
def test_lml_and_gradient_match_oracle(n, d):
    x, y, diag = problem(n, d, seed=n + d)
    gp = device_problem(x, diag, y)
    rng = np.random.default_rng(n)
    for trial in range(3):
        p = orc.default_parameter_vector(d) + (0.0 if trial == 0 else 0.4 * rng.standard_normal(d + 1))
        lml, grad, info, logdet, quad = gp.evaluate(p, True)
        ref_lml, ref_grad, ref_logdet, ref_quad = orc.lml_and_grad(p, x, diag, y)
        assert info == 0
        assert abs(lml - ref_lml) <= TOL * abs(ref_lml)
        assert abs(logdet - ref_logdet) <= TOL * max(abs(ref_logdet), 1.0)
        assert abs(quad - ref_quad) <= TOL * abs(ref_quad)
        assert np.max(np.abs(grad - ref_grad)) <= TOL * np.max(np.abs(ref_grad))
        lml_only, _, info2, _, _ = gp.evaluate(p, False)
        assert info2 == 0 and lml_only == lml  # same kernels, deterministic
        # value-only mode: Cholesky factor alone, the residual carried as an extra matrix row
        lml_v, grad_v, info3, logdet_v, quad_v = gp.evaluate(p, False, value_only=True)
        assert info3 == 0 and np.all(grad_v == 0.0)
        assert logdet_v == logdet  # same Cholesky tiles
        assert abs(quad_v - ref_quad) <= TOL * abs(ref_quad)
        assert abs(lml_v - ref_lml) <= TOL * abs(ref_lml)
    gp.close()


@pytest.mark.parametrize("n,d,m", [(50, 2, 1), (200, 3, 37), (640, 9, 300), (1000, 3, 1500), (129, 4, 128), (300, 8, 129)])
def test_prediction_matches_oracle(n, d, m):
    x, y, diag = problem(n, d, seed=7 * n)
    gp = device_problem(x, diag, y)
    p = orc.default_parameter_vector(d) + 0.2
    gp.evaluate(p, False)
    t = np.random.default_rng(m).random((m, d)).astype(np.float32).astype(np.float64)
    t[: min(m, 3)] = x[: min(m, 3)]  # include training points (variance close to the noise level)
    mu, var = gp.predict(t, True)
    ref_mu, ref_var = orc.predict(p, x, diag, y, t)
    assert relmax(mu, ref_mu) <= TOL
    assert np.max(np.abs(var - ref_var)) <= TOL * orc.amplitude(p, d)
    assert np.array_equal(gp.predict(t, False), mu)
    gp.close()


@pytest.mark.parametrize("n,d,n_theta,n_ind", [(300, 4, 7, 5), (640, 9, 300, 32), (200, 2, 1, 1), (1000, 3, 4097, 3)])
def test_grid_prediction_equals_row_by_row_prediction(n, d, n_theta, n_ind):
    """The grid form (float32-rounded like the emulators' query points) must give what predict gives on the same
    rows built on the host.  The grid evaluates the kernel in its factorised form, exp(a) * exp(b) instead of
    exp(a + b) -- one exponential per (parameter row, training point) instead of one per entry -- so the two agree to
    a few rounding units of the kernel values, not bit for bit."""
    x, y, diag = problem(n, d, seed=n + n_theta)
    gp = device_problem(x, diag, y)
    p = orc.default_parameter_vector(d) + 0.15
    gp.evaluate(p, False)
    rng = np.random.default_rng(n_theta)
    theta = rng.random((n_theta, d - 1)) * 1.7 - 0.3
    ind = np.linspace(0.0, 1.0, n_ind) + 1e-9
    rows = np.empty((n_theta * n_ind, d), dtype=np.float32)
    rows[:, 0] = np.tile(ind, n_theta)
    rows[:, 1:] = np.repeat(theta, n_ind, axis=0)
    mu_ref, var_ref = gp.predict(rows.astype(np.float64), True)
    mu, var = gp.predict_grid(theta, ind, True)
    assert mu.shape == (n_theta, n_ind)
    assert relmax(mu.ravel(), mu_ref) <= 1e-11 and np.max(np.abs(var.ravel() - var_ref)) <= 1e-12 * orc.amplitude(p, d)
    assert np.array_equal(gp.predict_grid(theta, ind, False), mu)
    # unrounded rows: against the oracle
    k = min(3, n_theta)
    mu64, var64 = gp.predict_grid(theta[:k], ind, True, round_f32=False)
    rows64 = np.column_stack([np.tile(ind, k), np.repeat(theta[:k], n_ind, axis=0)])
    ref = orc.predict(p, x, diag, y, rows64)
    assert relmax(mu64.ravel(), ref[0]) <= TOL and np.max(np.abs(var64.ravel() - ref[1])) <= TOL * orc.amplitude(p, d)
    gp.close()


def test_prediction_requires_factorisation_and_follows_latest_parameters():
    x, y, diag = problem(120, 3)
    gp = device_problem(x, diag, y)
    with pytest.raises(RuntimeError):
        gp.predict(x[:4], True)
    p1 = orc.default_parameter_vector(3)
    p2 = p1 + 0.5
    gp.evaluate(p1, True)
    mu1 = gp.predict(x[:9] + 0.01, False)
    gp.evaluate(p2, False)
    mu2 = gp.predict(x[:9] + 0.01, False)
    assert relmax(mu1, orc.predict(p1, x, diag, y, x[:9] + 0.01, False)) <= TOL
    assert relmax(mu2, orc.predict(p2, x, diag, y, x[:9] + 0.01, False)) <= TOL
    gp.close()


def test_value_only_evaluation_leaves_no_prediction_state():
    x, y, diag = problem(200, 3)
    gp = device_problem(x, diag, y)
    p = orc.default_parameter_vector(3) + 0.1
    gp.evaluate(p, True)
    mu = gp.predict(x[:5] + 0.02, False)
    lml_v, _, info, _, _ = gp.evaluate(p + 0.3, False, value_only=True)
    assert info == 0 and abs(lml_v - orc.lml_and_grad(p + 0.3, x, diag, y)[0]) <= TOL * abs(lml_v)
    with pytest.raises(RuntimeError):
        gp.predict(x[:5] + 0.02, False)  # L^-1 / alpha were not produced at p + 0.3
    with pytest.raises(ValueError):
        gp.evaluate(p, True, value_only=True)
    gp.evaluate(p, False)
    assert np.array_equal(gp.predict(x[:5] + 0.02, False), mu)
    assert gp.launches_per_eval(False, value_only=True) == 3  # parameters, factorisation, reduction
    # a new residual is picked up by the residual row
    gp.set_residual(2.0 * y)
    _, _, _, _, quad2 = gp.evaluate(p, False, value_only=True)
    assert abs(quad2 - 4.0 * orc.lml_and_grad(p, x, diag, y)[3]) <= 4 * TOL * abs(quad2)
    gp.close()


@pytest.mark.parametrize("n,d,nb", [(200, 3, 1), (300, 4, 3), (129, 2, 4), (640, 9, 7), (1536, 5, 4)])
def test_batched_value_only_evaluations_match_single_ones(n, d, nb):
    """Several factorisations share one launch; every one must equal its stand-alone evaluation bit for bit."""
    x, y, diag = problem(n, d, seed=3 * n + nb)
    gp = device_problem(x, diag, y)
    rng = np.random.default_rng(nb)
    ps = orc.default_parameter_vector(d) + 0.3 * rng.standard_normal((nb, d + 1))
    lml, info, logdet, quad = gp.evaluate_value_batch(ps)
    assert lml.shape == (nb,) and np.all(info == 0)
    for b in range(nb):
        single = gp.evaluate(ps[b], False, value_only=True)
        assert (lml[b], logdet[b], quad[b]) == (single[0], single[3], single[4])
        ref = orc.lml_and_grad(ps[b], x, diag, y)
        assert abs(lml[b] - ref[0]) <= TOL * abs(ref[0])
    # repeated batches are bit-identical, and the batch leaves no prediction state
    again = gp.evaluate_value_batch(ps)
    assert np.array_equal(again[0], lml)
    with pytest.raises(RuntimeError):
        gp.predict(x[:3], False)
    gp.close()


def test_batched_value_only_full_size_is_stable_under_repetition():
    """Regression: per-CTA panel flags were keyed by task number, which repeats across the problems of a
    batch; at N = 4096 a CTA helping a second problem could see a stale flag (intermittent NaN)."""
    n, d = 4096, 9
    x, y, diag = problem(n, d, seed=1235)
    gp = device_problem(x, diag, y)
    ps = orc.default_parameter_vector(d) + 0.05 * np.random.default_rng(0).standard_normal((11, d + 1))
    first = gp.evaluate_value_batch(ps)
    assert np.all(first[1] == 0) and np.all(np.isfinite(first[0]))
    for b in (0, 5, 10):
        assert gp.evaluate(ps[b], False, value_only=True)[0] == first[0][b]
    for _ in range(40):
        again = gp.evaluate_value_batch(ps)
        assert np.array_equal(again[1], first[1]) and np.array_equal(again[0], first[0])
    gp.close()


def test_batched_value_only_reports_failed_factorisations_per_problem():
    x = np.zeros((40, 2))
    x[:, 0] = np.repeat(np.arange(20), 2) * 0.05  # duplicated rows, no noise: singular at any p
    gp = device_problem(x, np.full(40, 0.0), np.ones(40))
    gp2 = device_problem(x, np.full(40, 1e-2), np.ones(40))
    ps = np.array([[5.0, 3.0, 3.0], [0.0, 0.0, 0.0], [1.0, -1.0, 2.0]])
    _, info, _, _ = gp.evaluate_value_batch(ps)
    assert np.all(info > 0)
    lml2, info2, _, _ = gp2.evaluate_value_batch(ps)
    assert np.all(info2 == 0) and np.all(np.isfinite(lml2))
    gp.close(); gp2.close()


def test_value_only_not_positive_definite_reports_info():
    x = np.zeros((40, 2))
    x[:, 0] = np.repeat(np.arange(20), 2) * 0.05
    gp = device_problem(x, np.zeros(40), np.ones(40))
    _, _, info, _, _ = gp.evaluate(np.array([5.0, 3.0, 3.0]), False, value_only=True)
    assert info > 0
    gp.close()


def test_intermediate_matrices_match_numpy():
    """K, its Cholesky factor, the triangular inverse, alpha and K^-1, stage by stage."""
    from scipy.linalg import cholesky, solve_triangular

    n, d = 333, 5
    x, y, diag = problem(n, d, seed=11)
    gp = device_problem(x, diag, y)
    npad = gp.padded_n
    assert npad == 384
    p = orc.default_parameter_vector(d) - 0.1
    k = orc.kernel_value(p, x)
    k[np.diag_indices(n)] += diag
    gp.debug_stage(p, 1)
    a = gp.debug_matrix(0)
    assert relmax(np.tril(a[:n, :n]), np.tril(k)) <= 1e-14
    assert np.array_equal(np.tril(a)[n:, :], np.eye(npad)[n:, :])  # identity padding
    lo = cholesky(k, lower=True)
    gp.debug_stage(p, 2)
    a = gp.debug_matrix(0)
    assert relmax(np.tril(a[:n, :n]), lo) <= 1e-10
    xi = solve_triangular(lo, np.eye(n), lower=True)
    w = gp.debug_matrix(1)  # the same launch also produced X = L^-1 in the second matrix
    assert relmax(np.tril(w[:n, :n]), xi) <= 1e-8
    assert np.all(w[np.triu_indices(npad, 1)] == 0.0)  # strict upper triangle of L^-1 is exactly zero
    gp.debug_stage(p, 3)
    assert relmax(gp.debug_alpha(), np.linalg.solve(k, y)) <= 1e-7
    gp.debug_stage(p, 4)
    a = gp.debug_matrix(0)
    assert relmax(np.tril(a[:n, :n]), np.tril(xi.T @ xi)) <= 1e-8
    gp.close()


def test_not_positive_definite_reports_info_and_recovers():
    x = np.zeros((40, 2))
    x[:, 0] = np.repeat(np.arange(20), 2) * 0.05  # duplicated rows
    diag = np.zeros(40)  # no noise at all: the duplicated rows make K exactly singular
    y = np.ones(40)
    gp = device_problem(x, diag, y)
    p = np.array([5.0, 3.0, 3.0])
    lml, grad, info, _, _ = gp.evaluate(p, True)
    with pytest.raises(np.linalg.LinAlgError):
        orc.lml_and_grad(p, x, diag, y)
    assert info > 0
    # the handle stays usable
    diag_ok = np.full(40, 1e-2)
    gp2 = device_problem(x, diag_ok, y)
    lml2, _, info2, _, _ = gp2.evaluate(orc.default_parameter_vector(2), True)
    assert info2 == 0 and abs(lml2 - orc.lml_and_grad(orc.default_parameter_vector(2), x, diag_ok, y)[0]) <= TOL * abs(lml2)
    gp.close(); gp2.close()


def test_many_handles_coexist():
    handles, refs = [], []
    for i in range(12):
        x, y, diag = problem(64 + 16 * i, 3, seed=i)
        handles.append(device_problem(x, diag, y))
        refs.append(orc.lml_and_grad(orc.default_parameter_vector(3), x, diag, y))
    for gp, ref in zip(handles, refs):
        lml, grad, info, _, _ = gp.evaluate(orc.default_parameter_vector(3), True)
        assert info == 0 and abs(lml - ref[0]) <= TOL * abs(ref[0])
        assert np.max(np.abs(grad - ref[1])) <= TOL * np.max(np.abs(ref[1]))
    for gp in handles:
        gp.close()


def test_residual_change_only_affects_quadratic_terms():
    x, y, diag = problem(200, 4)
    gp = device_problem(x, diag, y)
    p = orc.default_parameter_vector(4)
    a = gp.evaluate(p, True)
    gp.set_residual(2.0 * y)
    b = gp.evaluate(p, True)
    assert b[3] == a[3]  # logdet
    assert abs(b[4] - 4.0 * a[4]) <= 1e-12 * abs(b[4])  # quadratic form scales with r^2
    gp.close()


# ---- BASELINE.json full size (C2: N = 4096, kernel d = 9): oracle at lean cost + properties ----

def test_full_size_c2_against_oracle_and_properties():
    n, d = 4096, 9
    x, y, diag = problem(n, d, seed=1235)
    gp = device_problem(x, diag, y)
    ps = synthetic.perturbed_parameter_vectors(d, count=2, seed=7)
    for p in ps[:2]:
        lml, grad, info, logdet, quad = gp.evaluate(p, True)
        ref = orc.lml_and_grad_lean(p, x, diag, y)
        assert info == 0
        assert abs(lml - ref[0]) <= TOL * abs(ref[0])
        assert np.max(np.abs(grad - ref[1])) <= TOL * np.max(np.abs(ref[1]))
        value_only = gp.evaluate(p, False, value_only=True)
        assert value_only[2] == 0 and abs(value_only[0] - ref[0]) <= TOL * abs(ref[0])
    # size-independent properties
    p = ps[0]
    base = gp.evaluate(p, True)
    # (1) gradient = derivative of the value (central differences on the device values)
    for i in (0, 1, d):
        e = np.zeros(d + 1); e[i] = 1e-4
        fd = (gp.evaluate(p + e, False)[0] - gp.evaluate(p - e, False)[0]) / 2e-4
        assert abs(fd - base[1][i]) <= 1e-5 * max(1.0, np.max(np.abs(base[1])))
    gp.evaluate(p, False)
    # (2) prediction at training inputs interpolates within the noise; variance in [0, amp]
    mu, var = gp.predict(x[:512], True)
    assert np.max(np.abs(mu - y[:512])) < 0.25
    assert np.all(var > -1e-9) and np.all(var <= orc.amplitude(p, d) * (1 + 1e-12))
    # (3) far from the data the predictor reverts to the prior
    mu_far, var_far = gp.predict(np.full((4, d), 40.0), True)
    assert np.max(np.abs(mu_far)) < 1e-12 and np.allclose(var_far, orc.amplitude(p, d), rtol=1e-12)
    gp.close()
    # (4) row-permutation invariance of value and gradient (separate handle)
    perm = np.random.default_rng(0).permutation(n)
    gp2 = device_problem(x[perm], diag[perm], y[perm])
    other = gp2.evaluate(p, True)
    assert abs(other[0] - base[0]) <= TOL * abs(base[0])
    assert np.max(np.abs(other[1] - base[1])) <= 1e-8 * np.max(np.abs(base[1]))
    gp2.close()


def test_ragged_full_size_c4_shape():
    """C4's refit size N = 8176 (not a multiple of the tile): value against the lean oracle."""
    n, d = 8176, 9
    x, y, diag = problem(n, d, seed=1238)
    gp = device_problem(x, diag, y)
    p = orc.default_parameter_vector(d)
    lml, grad, info, logdet, quad = gp.evaluate(p, True)
    ref = orc.lml_and_grad_lean(p, x, diag, y)
    assert info == 0 and gp.padded_n == 8192
    assert abs(lml - ref[0]) <= TOL * abs(ref[0])
    assert np.max(np.abs(grad - ref[1])) <= TOL * np.max(np.abs(ref[1]))
    gp.close()


def test_device_pointer_entry_points_with_torch_memory():
    """The *_device C ABI calls on torch-owned device memory and a torch stream."""
    import torch

    n, d = 512, 4
    x, y, diag = problem(n, d, seed=3)
    gp = device_problem(x, diag, y)
    dev = torch.device("cuda:0")
    ps = synthetic.perturbed_parameter_vectors(d, count=3, seed=1)
    p_dev = torch.from_numpy(ps).to(dev)
    out = torch.zeros((len(ps), d + 5), dtype=torch.float64, device=dev)
    r_dev = torch.from_numpy(y).to(dev)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        gp.set_residual_device(r_dev.data_ptr(), stream.cuda_stream)
        for i in range(len(ps)):
            gp.evaluate_device(p_dev[i].data_ptr(), True, out[i].data_ptr(), stream.cuda_stream)
        t = torch.from_numpy(x[:100].copy()).to(dev)
        mu = torch.empty(100, dtype=torch.float64, device=dev)
        var = torch.empty(100, dtype=torch.float64, device=dev)
        gp.predict_device(100, t.data_ptr(), True, mu.data_ptr(), var.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    res = out.cpu().numpy()
    for i, p in enumerate(ps):
        ref = orc.lml_and_grad(p, x, diag, y)
        assert abs(res[i, 0] - ref[0]) <= TOL * abs(ref[0])
        assert np.max(np.abs(res[i, 1 : d + 2] - ref[1])) <= TOL * np.max(np.abs(ref[1]))
    ref_mu, ref_var = orc.predict(ps[-1], x, diag, y, x[:100])
    assert relmax(mu.cpu().numpy(), ref_mu) <= TOL
    assert np.max(np.abs(var.cpu().numpy() - ref_var)) <= TOL * orc.amplitude(ps[-1], d)
    gp.close()


def test_repeated_evaluations_are_bit_identical():
    """The dataflow kernel's task interleaving varies from launch to launch; the arithmetic must not
    (fixed accumulation order per tile, ordered final reductions, no floating-point atomics)."""
    for n, d, reps in ((1000, 3, 60), (4096, 9, 12)):
        x, y, diag = problem(n, d, seed=17)
        gp = device_problem(x, diag, y)
        ps = synthetic.perturbed_parameter_vectors(d, count=2, seed=5, scale=0.3)
        first = [gp.evaluate(p, True) for p in ps]
        t = x[:64] + 0.01
        gp.evaluate(ps[1], True)
        mu0, var0 = gp.predict(t, True)
        for _ in range(reps):
            for p, ref in zip(ps, first):
                got = gp.evaluate(p, True)
                assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and got[2] == 0
        gp.evaluate(ps[1], True)
        mu1, var1 = gp.predict(t, True)
        assert np.array_equal(mu0, mu1) and np.array_equal(var0, var1)
        gp.close()


def test_concurrent_handles_from_threads_agree_with_sequential():
    """Several handles driven from host threads at once (what the per-bin and leave-one-out drivers do)."""
    from concurrent.futures import ThreadPoolExecutor

    problems = [problem(700 + 100 * i, 4, seed=30 + i) for i in range(6)]
    p = orc.default_parameter_vector(4) + 0.1
    sequential = []
    for x, y, diag in problems:
        gp = device_problem(x, diag, y)
        sequential.append(gp.evaluate(p, True))
        gp.close()

    def work(args):
        x, y, diag = args
        gp = device_problem(x, diag, y)
        out = [gp.evaluate(p, True) for _ in range(5)][-1]
        gp.close()
        return out

    with ThreadPoolExecutor(max_workers=6) as pool:
        threaded = list(pool.map(work, problems))
    for a, b in zip(sequential, threaded):
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and b[2] == 0


def test_empty_and_invalid_inputs():
    """Empty query / parameter blocks return empty results; impossible problems are refused with a message."""
    from emulator_b200.device import DeviceProblem

    x, y, diag = problem(70, 3)
    gp = device_problem(x, diag, y)
    p = orc.default_parameter_vector(3)
    gp.evaluate(p, False)
    mu, var = gp.predict(np.zeros((0, 3)), True)
    assert mu.shape == (0,) and var.shape == (0,)
    mu, var = gp.predict_grid(np.zeros((0, 2)), np.array([0.5]), True)
    assert mu.shape == (0, 1) and var.shape == (0, 1)
    lml, info, _, _ = gp.evaluate_value_batch(np.zeros((0, 4)))
    assert lml.shape == (0,) and info.shape == (0,)
    with pytest.raises(ValueError):
        gp.evaluate(np.zeros(3), True)            # wrong parameter count
    with pytest.raises(ValueError):
        gp.predict(np.zeros((4, 2)), True)        # wrong query dimension
    with pytest.raises(ValueError):
        gp.set_residual(np.zeros(69))             # wrong residual length
    with pytest.raises(ValueError):
        gp.predict_grid(np.zeros((2, 3)), np.array([0.5]), True)  # theta must have d - 1 columns
    gp.close()
    with pytest.raises(ValueError):
        DeviceProblem(np.zeros((5, 17)), np.ones(5))   # kernel dimension above the supported maximum
    with pytest.raises(RuntimeError):
        DeviceProblem(np.zeros((0, 2)), np.ones(0))    # no training rows
    # non-finite hyperparameters: the factorisation fails (info != 0) or the value is not finite, never a hang
    gp = device_problem(x, diag, y)
    out = gp.evaluate(np.array([np.nan, 0.0, 0.0, 0.0]), True)
    assert out[2] != 0 or not np.isfinite(out[0])
    out = gp.evaluate(np.array([800.0, 0.0, 0.0, 0.0]), False, value_only=True)   # exp overflow
    assert out[2] != 0 or not np.isfinite(out[0])
    assert gp.evaluate(p, True)[2] == 0  # and the handle recovers
    gp.close()


def test_device_against_extended_precision_on_an_ill_conditioned_problem():
    """Where K is ill conditioned (d = 1, dense design: cond(K) ~ 1e8) the device and the fp64 oracle each
    differ from the exact answer by ~cond(K) * 2^-53; the longdouble oracle is the yardstick: the device
    must be as close to it as the CPU path is (within a small factor), and exact to 1e-11 when K is
    well conditioned."""
    from oracle import gp_oracle_extended as ext

    rng = np.random.default_rng(58)
    for n, d, shift, well in ((90, 4, np.zeros(5), True), (144, 1, np.array([0.3, -0.4]), False)):
        x = rng.random((n, d)).astype(np.float32).astype(np.float64)
        y = np.sin(3.0 * x.sum(axis=1)) + 0.1 * rng.standard_normal(n)
        noise = orc.noise_diagonal((0.02 * np.abs(y) + 1e-3).astype(np.float32))
        p = orc.default_parameter_vector(d) + shift
        exact = ext.lml_and_grad(p, x, noise, y)
        e_lml, e_grad = float(exact[0]), np.asarray(exact[1], dtype=np.float64)
        cpu = orc.lml_and_grad(p, x, noise, y)
        gp = device_problem(x, noise, y)
        lml, grad, info, _, _ = gp.evaluate(p, True)
        assert info == 0
        dev_l, dev_g = abs(lml - e_lml) / abs(e_lml), np.max(np.abs(grad - e_grad)) / np.max(np.abs(e_grad))
        cpu_l, cpu_g = abs(cpu[0] - e_lml) / abs(e_lml), np.max(np.abs(cpu[1] - e_grad)) / np.max(np.abs(e_grad))
        if well:
            assert dev_l <= 1e-11 and dev_g <= 1e-11
        else:
            k = orc.kernel_value(p, x)
            k[np.diag_indices(n)] += noise
            floor = np.linalg.cond(k) * 2.0 ** -53
            assert dev_l <= 50 * floor and dev_g <= 50 * floor
            assert dev_g <= 10 * max(cpu_g, floor) and dev_l <= 10 * max(cpu_l, floor)
        gp.close()


@pytest.mark.parametrize("n,d,nb", [(256, 8, 11), (1000, 3, 9), (2048, 9, 3), (130, 2, 2)])
def test_multi_handle_evaluation_equals_single_evaluations(n, d, nb):
    """evaluate_many: equally shaped chain-bound problems share factorisation launches (chains interleaved);
    every result must equal the handle's own stand-alone evaluation bit for bit, in every mode."""
    from emulator_b200.device import evaluate_many

    probs, ps, singles = [], [], []
    for k in range(nb):
        x, y, diag = problem(n, d, seed=100 * n + k)
        gp = device_problem(x, diag, y)
        p = orc.default_parameter_vector(d) + 0.2 * np.random.default_rng(k).standard_normal(d + 1)
        probs.append(gp); ps.append(p)
        singles.append(gp.evaluate(p, True))
    ps = np.array(ps)
    for rep in range(3):
        lml, grad, info = evaluate_many(probs, ps, want_grad=True)
        assert np.all(info == 0)
        for k in range(nb):
            assert lml[k] == singles[k][0] and np.array_equal(grad[k], singles[k][1])
    lml2, grad2, info2 = evaluate_many(probs, ps, want_grad=False)
    assert np.array_equal(lml2, lml) and np.all(grad2 == 0.0) and np.all(info2 == 0)
    # prediction state is valid on every handle afterwards
    t = problem(n, d, seed=7)[0][:5]
    mu = probs[nb - 1].predict(t, False)
    ref = orc.predict(ps[nb - 1], *[problem(n, d, seed=100 * n + nb - 1)[i] for i in (0, 2, 1)], t, False)
    assert relmax(mu, ref) <= TOL
    # value-only requests go through the handles' own graphs
    lml3, _, info3 = evaluate_many(probs, ps, want_grad=False, value_only=True)
    assert np.all(info3 == 0) and np.max(np.abs(lml3 - lml) / np.abs(lml)) <= 1e-12
    for gp in probs:
        gp.close()
