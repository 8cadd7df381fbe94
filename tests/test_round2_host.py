"""
Host-side logic added in round 2, on CPU: the evaluation combiner, dynamic work claiming, the leave_out subset of
CrossCheck, a known-answer test of the Sobol estimators, trajectory goldens through the oracle-backed flow, and
walker sharding over two gloo ranks.
"""
import os
import socket
import sys
import threading

import numpy as np
import pytest
import torch.multiprocessing as mp

from emulator_b200 import combine, dist, synthetic
from emulator_b200.backend import ModelSpecification
from emulator_b200.sensitivity import CrossCheck, saltelli_design, sobol_indices
from test_host_logic import golden_inputs, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- evaluation combiner ------------------------------------------------------------------------------------
class FakeProblem:
    d = 2

    def __init__(self, tag):
        self.tag = tag


def test_combiner_merges_the_requests_of_all_active_threads(monkeypatch):
    calls = []

    def fake_evaluate_many(problems, ps, want_grad=True, value_only=False):
        calls.append([q.tag for q in problems])
        ps = np.asarray(ps)
        return ps.sum(axis=1), 2.0 * ps, np.zeros(len(problems), dtype=np.int64)

    monkeypatch.setattr(combine, "evaluate_many", fake_evaluate_many)
    comb = combine.EvaluationCombiner(3)
    results = {}

    def work(tag, rounds):
        with comb:
            assert combine.current() is comb
            q = FakeProblem(tag)
            for r in range(rounds):
                p = np.array([tag, r, 1.0])
                lml, grad, info = comb.evaluate(q, p, True)
                assert lml == p.sum() and np.array_equal(grad, 2.0 * p) and info == 0
                results[(tag, r)] = lml
        assert combine.current() is None

    threads = [threading.Thread(target=work, args=(t, 3 + t)) for t in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=30)
        assert not t.is_alive(), "combiner deadlocked"
    assert len(results) == 3 + 4 + 5
    # three full rounds are shared by all three threads, then the group shrinks as threads retire
    assert sorted(len(c) for c in calls) == [1, 2, 3, 3, 3]
    assert comb.calls == 5 and comb.evaluations == 12


def test_combiner_hands_failures_to_every_waiting_thread(monkeypatch):
    def broken(problems, ps, want_grad=True, value_only=False):
        raise RuntimeError("device lost")

    monkeypatch.setattr(combine, "evaluate_many", broken)
    comb = combine.EvaluationCombiner(2)
    seen = []

    def work():
        try:
            with comb:
                comb.evaluate(FakeProblem(0), np.zeros(3), True)
        except RuntimeError as exc:
            seen.append(str(exc))

    threads = [threading.Thread(target=work) for _ in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=30)
        assert not t.is_alive()
    assert seen == ["device lost", "device lost"]


def test_work_claimer_without_a_process_group_is_the_round_robin_split():
    claimer = dist.WorkClaimer(5)
    got = []
    while True:
        i = claimer.next()
        if i is None:
            break
        got.append(i)
    assert got == [0, 1, 2, 3, 4] == claimer.taken
    assert dist.resolve_device(None) == 0 and dist.resolve_device(3) == 3


# ---- CrossCheck subset ------------------------------------------------------------------------------------------
def test_cross_check_leave_out_subset_matches_the_full_run(oracle_backend):
    spec, params, values = synthetic.make_model(2, 7, 5, seed=33)
    full = CrossCheck(workers=1)
    full.build_emulators(spec, params, values)
    uids = list(values.model_values.keys())
    subset = CrossCheck(workers=3, leave_out=[uids[4], uids[1]])
    subset.build_emulators(spec, params, values)
    assert subset.leave_out_order == [uids[1], uids[4]]  # the container's order, as in the full run
    assert np.array_equal(subset.cross_fit_parameters, full.cross_fit_parameters[[1, 4]])
    total, per_model = subset.get_mean_squared()
    _, per_full = full.get_mean_squared()
    assert sorted(per_model.keys()) == [uids[1], uids[4]]
    assert np.array_equal(per_model[uids[4]], per_full[uids[4]])
    assert list(subset.build_mocked_model_values(np.linspace(0.1, 0.9, 4)).model_values.keys()) == [uids[1], uids[4]]
    with pytest.raises(KeyError):
        CrossCheck(leave_out=["nope"]).build_emulators(spec, params, values)
    assert subset.build_stats["refits"] == 2


# ---- "identical optima": the goldens through the oracle-backed flow -----------------------------------------------
@pytest.mark.parametrize("name", ["trajectory_c1"])
def test_trajectory_golden_is_reproduced_by_the_host_flow(oracle_backend, name, monkeypatch):
    """Same oracle arithmetic + same scipy => the reference's BFGS run, iterate for iterate (pins the host flow the
    GPU trajectory test relies on: tests/test_gpu_round2.py)."""
    import emulator_b200.emulators.base as base
    from emulator_b200.emulators import GaussianProcessEmulator
    from scipy.optimize import minimize

    g = load(name)
    spec, params, values = golden_inputs(g)
    iterates = []

    def recording(fun, x0, jac=None, **kw):
        iterates.append(np.array(x0, dtype=np.float64))
        return minimize(fun, x0, jac=jac, callback=lambda xk: iterates.append(np.array(xk)), **kw)

    monkeypatch.setattr(base, "minimize", recording)
    emu = GaussianProcessEmulator()
    emu.fit_model(spec, params, values)
    assert emu.fit_result.nit == int(g["nit"]) and emu.fit_result.nfev == int(g["nfev"])
    assert np.allclose(np.array(iterates), g["iterates"], rtol=0, atol=1e-9)


# ---- Sobol estimators: known answer ----------------------------------------------------------------------------------
class AnalyticEmulator:
    """An 'emulator' whose mean is an additive-plus-interaction function with analytic Sobol indices:
    f = a1 x1 + a2 x2 + b x1 x3 (+ c t), x_i ~ U[0, 1]."""

    parameter_order = ["x1", "x2", "x3"]
    a1, a2, b = 2.0, 1.0, 3.0

    def predict_batch(self, rows, want_var=True):
        t, x1, x2, x3 = (rows[:, k].astype(np.float64) for k in range(4))
        f = self.a1 * x1 + self.a2 * x2 + self.b * x1 * x3 + 0.5 * t
        return (f, np.zeros_like(f)) if want_var else f


def test_sobol_indices_of_a_function_with_known_indices():
    spec = ModelSpecification(number_of_parameters=3, parameter_names=["x1", "x2", "x3"], parameter_limits=[[0.0, 1.0]] * 3)
    emu = AnalyticEmulator()
    a1, a2, b = emu.a1, emu.a2, emu.b
    # analytic variance decomposition (x_i independent U[0,1]: var = 1/12, mean = 1/2)
    v1 = (a1 + b / 2) ** 2 / 12
    v2 = a2 ** 2 / 12
    v3 = (b / 2) ** 2 / 12
    v13 = b ** 2 / 144
    total = v1 + v2 + v3 + v13
    s1 = np.array([v1, v2, v3]) / total
    st = np.array([v1 + v13, v2, v3 + v13]) / total
    out = sobol_indices(emu, spec, np.array([0.0, 0.7]), base_samples=1 << 14, seed=11)
    assert out["S1"].shape == (3, 2) and out["names"] == ["x1", "x2", "x3"]
    for column in range(2):  # the independent variable only shifts f: same indices at every value
        assert np.max(np.abs(out["S1"][:, column] - s1)) <= 5e-3
        assert np.max(np.abs(out["ST"][:, column] - st)) <= 5e-3
    design = saltelli_design(spec, 8, seed=1)
    assert design.shape == (8 * 5, 3) and np.array_equal(design[16:24, 0], design[8:16, 0]) and np.array_equal(design[16:24, 1], design[:8, 1])


# ---- walker sharding over two gloo ranks -------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _mcmc_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as td

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_george
    import emulator_b200.george_compat as gc

    gc.GP = oracle_george.OracleBackedGP
    from emulator_b200 import dist as d
    from emulator_b200 import synthetic as syn
    from emulator_b200.emulators import GaussianProcessEmulatorMCMC

    spec, params, values = syn.make_model(2, 8, 6, seed=71)
    emu = GaussianProcessEmulatorMCMC(burn_in_steps=3, mcmc_steps=5, walkers=8, seed=4)
    emu.fit_model(spec, params, values)
    np.save(os.path.join(out_dir, f"chain_{rank}.npy"), emu.hyperparameter_samples)
    # dynamic claiming: every index handed out exactly once over the ranks
    claimer = d.WorkClaimer(9)
    mine = []
    while True:
        i = claimer.next()
        if i is None:
            break
        mine.append(i)
    owner = np.zeros(9)
    owner[mine] = 1.0
    assert np.array_equal(d.all_reduce_sum(owner), np.ones(9))
    td.barrier()
    td.destroy_process_group()


def test_walkers_sharded_over_two_ranks_walk_the_single_rank_chain(tmp_path, oracle_backend):
    from emulator_b200.emulators import GaussianProcessEmulatorMCMC

    mp.spawn(_mcmc_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    c0, c1 = np.load(tmp_path / "chain_0.npy"), np.load(tmp_path / "chain_1.npy")
    assert np.array_equal(c0, c1)
    spec, params, values = synthetic.make_model(2, 8, 6, seed=71)
    emu = GaussianProcessEmulatorMCMC(burn_in_steps=3, mcmc_steps=5, walkers=8, seed=4)
    emu.fit_model(spec, params, values)
    assert np.array_equal(emu.hyperparameter_samples, c0)  # world = 1 and world = 2: the same chain


# ---- optimiser restarts ------------------------------------------------------------------------------------------
def test_restarts_keep_the_reference_run_as_start_zero_and_never_do_worse(oracle_backend):
    from emulator_b200.emulators import GaussianProcessEmulator

    spec, params, values = synthetic.make_model(2, 8, 6, seed=5)
    single = GaussianProcessEmulator()
    single.fit_model(spec, params, values)
    multi = GaussianProcessEmulator(restarts=3, restart_seed=2)
    multi.fit_model(spec, params, values)
    funs = multi.fit_result.restart_funs
    assert len(funs) == 4 and funs[0] == single.fit_result.fun  # start 0 is the reference's own run
    assert multi.fit_result.fun == funs.min() and multi.fit_result.restart_index == int(np.argmin(funs))
    assert np.array_equal(multi.emulator.get_parameter_vector(), multi.fit_result.x)
    again = GaussianProcessEmulator(restarts=3, restart_seed=2)
    again.fit_model(spec, params, values)
    assert np.array_equal(again.fit_result.restart_funs, funs)  # seeded


# ---- recycled host buffers ------------------------------------------------------------------------------------------
def test_pinned_pool_recycles_a_buffer_only_when_nothing_refers_to_it():
    import gc

    pool = dist.PinnedPool()
    a1 = pool.take((4, 3)).numpy()
    a1[:] = 1
    a2 = pool.take((4, 3)).numpy()
    a2[:] = 2
    assert a1.sum() == 12 and a2.sum() == 24 and len(pool._entries) == 2  # two live results, two buffers
    piece = a1[:2]
    del a1
    gc.collect()
    pool.take((4, 3))
    assert len(pool._entries) == 3  # a slice of the first result still lives: not recycled
    assert piece.sum() == 6
    del piece, a2
    gc.collect()
    before = len(pool._entries)
    v = pool.take((2, 3))
    assert len(pool._entries) == before  # everything was released: a buffer is reused
    del v
    pool.reserve((100,), 2)
    assert sum(e.numel() >= 100 for e in pool._entries) >= 2


# ---- leave-one-out post-processing ---------------------------------------------------------------------------------------
def test_lazy_emulators_build_on_first_access_only():
    from emulator_b200.sensitivity.cross_check import _LazyEmulators

    built = []

    def build(uid):
        built.append(uid)
        return f"emulator-{uid}"

    lazy = _LazyEmulators(["a", "b", "c"], {"b": "fitted-b"}, build)
    assert list(lazy.keys()) == ["a", "b", "c"] and len(lazy) == 3 and built == []
    assert lazy["b"] == "fitted-b" and built == []
    assert lazy["c"] == "emulator-c" and lazy["c"] == "emulator-c" and built == ["c"]  # built once
    assert [uid for uid, _ in lazy.built_items()] == ["b", "c"]
    with pytest.raises(KeyError):
        lazy["nope"]
    assert dict(lazy.items())["a"] == "emulator-a" and built == ["c", "a"]


def test_leave_one_out_predictions_with_ragged_independents(oracle_backend):
    """Simulations with different numbers of points: the sharded collection pads per simulation and cuts back."""
    spec, params, values = synthetic.make_model(2, 6, 5, seed=34)
    raw = {uid: dict(v) for uid, v in values.model_values.items()}
    first = next(iter(raw))
    raw[first] = {k: np.asarray(v)[:3] for k, v in raw[first].items()}  # one simulation has only three points
    from emulator_b200.backend import ModelValues

    ragged = ModelValues(model_values=raw)
    cc = CrossCheck(workers=2)
    cc.build_emulators(spec, params, ragged)
    total, per_model = cc.get_mean_squared()
    assert len(per_model[first]) == 3 and all(len(per_model[uid]) == 5 for uid in raw if uid != first)
    mocked = cc.build_mocked_model_values_original_independent()
    for uid in raw:
        direct, var = cc.cross_emulators[uid].predict_values(raw[uid]["independent"], params.model_parameters[uid])
        assert np.array_equal(mocked.model_values[uid]["dependent"], direct)
        assert np.array_equal(mocked.model_values[uid]["dependent_error"], np.sqrt(var))
    assert np.isclose(total, np.mean(np.concatenate([per_model[uid] for uid in raw])))
