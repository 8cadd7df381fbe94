"""
Leave-one-out cross checks -- API of
/root/reference/swiftemulator/sensitivity/cross_check.py:30-345 (plots excluded) and
cross_check_bins.py:26-113.  The S refits are independent: they are assigned round-robin to
ranks (one process per GPU) and only the fitted hyperparameter vectors are all-gathered, so
every rank ends with a usable emulator for every left-out simulation.

The reference pops the left-out entry from the caller's dict and re-inserts it, which rotates
the row order from refit to refit (:99, :114).  Here each refit sees the original order with
the one simulation removed and the caller's container is never mutated.
"""

from __future__ import annotations

import threading
from collections.abc import Mapping
from typing import Dict, Hashable, List, Optional

import attr
import numpy as np

from .. import dist
from ..backend import ModelParameters, ModelSpecification, ModelValues
from ..combine import EvaluationCombiner
from ..emulators.base import build_gp, default_kernel, optimise_gps_lockstep
from ..emulators.gaussian_process import GaussianProcessEmulator
from ..emulators.gaussian_process_bins import GaussianProcessEmulatorBins


class _LazyEmulators(Mapping):
    """{uid: emulator} over `order`: entries this rank fitted are there, the others are built on first access."""

    def __init__(self, order, built, build):
        self._order, self._built, self._build = list(order), dict(built), build

    def __getitem__(self, uid):
        if uid not in self._built:
            if uid not in self._order:
                raise KeyError(uid)
            self._built[uid] = self._build(uid)
        return self._built[uid]

    def __iter__(self):
        return iter(self._order)

    def __len__(self):
        return len(self._order)

    def built_items(self):
        """The entries that exist already (no construction)."""
        return list(self._built.items())


def _without(model_values: ModelValues, uid) -> ModelValues:
    return ModelValues(model_values={k: v for k, v in model_values.model_values.items() if k != uid})


@attr.s
class CrossCheck(object):
    kernel = attr.ib(default=None)
    mean_model = attr.ib(default=None)
    hide_progress: bool = attr.ib(default=True)
    device: Optional[int] = attr.ib(default=None)  # None: this rank's GPU under torchrun, else GPU 0 (dist.resolve_device)
    #: leave-one-out refits of this rank run from this many host threads, each with scipy's BFGS on its own GP;
    #: their likelihood evaluations are combined into shared device launches (emulator_b200/combine.py): the
    #: Cholesky dependency chain of one refit leaves SMs idle that the other refits' tiles use.  Iterates and
    #: optima are those of a sequential run, bit for bit.
    workers: int = attr.ib(default=4)
    #: "scipy": every refit is optimised by scipy's BFGS as in the reference; "lockstep": the refits of a rank
    #: advance together through emulator_b200.optimise (same optima, different iterates; pays for small
    #: problems, where the per-evaluation Python overhead dominates)
    fit_mode: str = attr.ib(default="scipy")
    #: restrict the cross check to these unique identifiers (extension: the reference always leaves out every
    #: simulation in turn, cross_check.py:98)
    leave_out: Optional[List[Hashable]] = attr.ib(default=None)
    #: refits are handed out one at a time to whichever rank / thread asks next (a counter in the process group's
    #: store) instead of i mod world: BFGS takes 60-80 evaluations per refit, a fixed split leaves ranks idle
    dynamic_claim: bool = attr.ib(default=True)
    lockstep_result = None
    _fitted_here = frozenset()
    #: diagnostics of the last build: refits done by this rank, their device evaluations, combined native calls
    build_stats: Optional[Dict[str, float]] = None

    model_specification: ModelSpecification = None
    model_parameters: ModelParameters = None
    model_values: ModelValues = None

    leave_out_order: Optional[List[Hashable]] = None
    cross_emulators: Optional[Dict[Hashable, GaussianProcessEmulator]] = None
    cross_fit_parameters: Optional[np.ndarray] = None

    def build_emulators(self, model_specification, model_parameters, model_values):
        self.model_specification = model_specification
        self.model_parameters = model_parameters
        self.model_values = model_values
        self.leave_out_order = list(model_values.model_values.keys())
        if self.leave_out is not None:
            wanted = set(self.leave_out)
            missing = wanted - set(self.leave_out_order)
            if missing:
                raise KeyError(f"leave_out names unknown simulations: {sorted(missing, key=str)}")
            self.leave_out_order = [uid for uid in self.leave_out_order if uid in wanted]
        count = len(self.leave_out_order)
        n_hyper = model_specification.number_of_parameters + 2 if self.kernel is None else len(self.kernel)

        fitted = np.zeros((count, n_hyper))
        emulators = {}

        def refit(i):
            uid = self.leave_out_order[i]
            emulator = GaussianProcessEmulator(kernel=self.kernel, mean_model=self.mean_model, device=self.device)
            emulator.fit_model(model_specification, model_parameters, _without(model_values, uid))
            return i, uid, emulator

        if self.fit_mode not in ("scipy", "lockstep"):
            raise ValueError("fit_mode must be 'scipy' or 'lockstep'")
        stats = {"combined_calls": 0, "combined_evaluations": 0, "combined_seconds": 0.0}
        if self.fit_mode == "lockstep" and self.kernel is None and self.mean_model is None:
            # all refits of this rank advance together (emulator_b200/optimise.py): one batched device call
            # per optimiser round.  Every refit has the same shape (one simulation left out).
            mine = dist.my_indices(count)
            done = []
            for i in mine:
                uid = self.leave_out_order[i]
                emulator = GaussianProcessEmulator(device=self.device)
                emulator._build_arrays(model_specification, model_parameters, _without(model_values, uid))
                emulator.kernel = default_kernel(model_specification.number_of_parameters + 1)
                emulator.emulator = build_gp(
                    emulator.kernel, None, emulator.independent_variables, emulator.dependent_variables,
                    emulator.dependent_variable_errors, device=self.device,
                )
                done.append((i, uid, emulator))
            if done:
                self.lockstep_result = optimise_gps_lockstep(
                    [e.emulator for _, _, e in done], [e.dependent_variables for _, _, e in done]
                )
        else:
            world = dist.world_size()
            # shared kernel / mean-model objects are mutated by fit_model: those refits stay sequential
            threads = max(1, self.workers) if (self.kernel is None and self.mean_model is None) else 1
            threads = min(threads, max(1, -(-count // world)))  # no more threads than refits per rank
            # Handing refits out one at a time balances BFGS runs of 60-80 evaluations when a rank gets many of them;
            # with only a few per rank the last one claimed starts when the others are nearly done and runs alone
            # (measured, 32 refits on 8 GPUs: 3 .. 6 refits per rank, the slowest rank 7.4 s against 4.7 s for a rank
            # with exactly four), so small jobs keep the i mod world split.
            claimer = dist.WorkClaimer(count, dynamic=self.dynamic_claim and count >= 4 * world * threads)
            done, errors = [], []
            combiner = EvaluationCombiner(threads)

            def work():
                try:
                    with combiner:
                        while True:
                            i = claimer.next()
                            if i is None:
                                break
                            done.append(refit(i))
                except BaseException as exc:  # surfaced on the calling thread below
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
            stats["combined_calls"], stats["combined_evaluations"] = combiner.calls, combiner.evaluations
            stats["combined_seconds"] = combiner.seconds
        owner = np.zeros(count)
        for i, uid, emulator in done:
            fitted[i] = emulator.emulator.get_parameter_vector()
            owner[i] = 1.0
            emulators[uid] = emulator
        stats["refits"] = len(done)
        stats["device_evaluations"] = int(sum(getattr(e.emulator, "n_device_evaluations", 0) for _, _, e in done))
        self.build_stats = stats
        # every refit was done by exactly one rank, the others hold zeros: a sum collects them all
        fitted = dist.all_reduce_sum(fitted)
        if not np.all(dist.all_reduce_sum(owner) == 1.0):
            raise RuntimeError("leave-one-out refits were not each done exactly once")
        self.cross_fit_parameters = fitted
        # refits done on other ranks: the GP at the gathered optimum (no optimisation) is built when somebody asks for
        # it -- constructing 500 emulator objects nobody may look at costs more host time than a refit's share of the GPU
        def rebuild(uid):
            i = self.leave_out_order.index(uid)
            emulator = GaussianProcessEmulator(kernel=self.kernel, mean_model=self.mean_model, device=self.device)
            emulator._build_arrays(model_specification, model_parameters, _without(model_values, uid))
            if emulator.kernel is None:
                emulator.kernel = default_kernel(model_specification.number_of_parameters + 1)
            gp = build_gp(
                emulator.kernel, emulator.mean_model, emulator.independent_variables, emulator.dependent_variables,
                emulator.dependent_variable_errors, device=self.device,
            )
            gp.set_parameter_vector(fitted[i])
            emulator.emulator = gp
            return emulator

        self._fitted_here = set(emulators.keys())
        self.cross_emulators = _LazyEmulators(self.leave_out_order, emulators, rebuild)

    def _require_built(self):
        if self.cross_emulators is None:
            raise AttributeError("You need to build the emulators before the prediction step.")

    def _leave_one_out_predictions(self, independent_of):
        """
        {uid: (emulated, variance)} of every left-out simulation at `independent_of(uid)`, each predicted by the
        emulator that never saw it.  Sharded like the refits: every rank predicts with the emulators it fitted itself
        (their device state is still there), the rows are collected with one all-reduce (each simulation has exactly
        one owner, the others hold zeros), and nobody re-creates -- and re-factorises -- the emulators of other ranks.
        """
        self._require_built()
        order = self.leave_out_order
        points = [np.asarray(independent_of(uid)) for uid in order]
        width = max((len(x) for x in points), default=0)
        table = np.zeros((len(order), 2, max(width, 1)))
        mine = self._fitted_here if dist.world_size() > 1 else set(order)
        for k, uid in enumerate(order):
            if uid in mine and len(points[k]):
                emulated, variance = self.cross_emulators[uid].predict_values(
                    independent=points[k], model_parameters=self.model_parameters.model_parameters[uid]
                )
                table[k, 0, : len(points[k])], table[k, 1, : len(points[k])] = emulated, variance
        table = dist.all_reduce_sum(table)
        return {uid: (table[k, 0, : len(points[k])].copy(), table[k, 1, : len(points[k])].copy()) for k, uid in enumerate(order)}

    def build_mocked_model_values_original_independent(self) -> ModelValues:
        predicted = self._leave_one_out_predictions(lambda uid: self.model_values.model_values[uid]["independent"])
        out = {}
        for uid in self.leave_out_order:
            emulated, variance = predicted[uid]
            out[uid] = {"independent": self.model_values.model_values[uid]["independent"], "dependent": emulated,
                        "dependent_error": np.sqrt(variance)}
        return ModelValues(out)

    def build_mocked_model_values(self, emulate_at: np.array) -> ModelValues:
        predicted = self._leave_one_out_predictions(lambda uid: emulate_at)
        out = {}
        for uid in self.leave_out_order:
            emulated, variance = predicted[uid]
            out[uid] = {"independent": emulate_at, "dependent": emulated, "dependent_error": np.sqrt(variance)}
        return ModelValues(out)

    def get_mean_squared(self, use_dependent_error: bool = False, use_y_as_error: bool = False,
                         use_squared_difference: bool = True):
        predicted = self._leave_one_out_predictions(lambda uid: self.model_values.model_values[uid]["independent"])
        per_model, everything = {}, []
        for uid in self.leave_out_order:
            y_model = self.model_values.model_values[uid]["dependent"]
            emulated, _ = predicted[uid]
            if use_y_as_error:
                weight = y_model
            else:
                weight = self.model_values.model_values[uid].get("dependent_error", None)
            difference = (y_model - emulated) / weight if use_dependent_error else y_model - emulated
            if use_squared_difference:
                difference = difference ** 2
            per_model[uid] = difference
            everything.extend(difference)
        return np.mean(everything), per_model

    def plot_results(self, *args, **kwargs):
        raise NotImplementedError("plotting (matplotlib) is outside the accelerated hot path")


@attr.s
class CrossCheckBins(object):
    kernel = attr.ib(default=None)
    mean_model = attr.ib(default=None)
    hide_progress: bool = attr.ib(default=True)
    device: Optional[int] = attr.ib(default=None)  # None: this rank's GPU under torchrun, else GPU 0 (dist.resolve_device)

    model_specification: ModelSpecification = None
    model_parameters: ModelParameters = None
    model_values: ModelValues = None

    leave_out_order: Optional[List[Hashable]] = None
    cross_emulators: Optional[Dict[Hashable, GaussianProcessEmulatorBins]] = None

    def build_emulators(self, model_specification, model_parameters, model_values):
        self.model_specification = model_specification
        self.model_parameters = model_parameters
        self.model_values = model_values
        self.leave_out_order = list(model_values.model_values.keys())
        emulators = {}
        for uid in self.leave_out_order:
            # each binned refit shards its own bins over the ranks
            emulator = GaussianProcessEmulatorBins(kernel=self.kernel, mean_model=self.mean_model, device=self.device)
            emulator.fit_model(model_specification, model_parameters, _without(model_values, uid))
            emulators[uid] = emulator
        self.cross_emulators = emulators

    def get_mean_squared(self, use_dependent_error: bool = False, use_y_as_error: bool = False,
                         use_squared_difference: bool = True):
        """Evident intent of cross_check_bins.py:190-249 (broken there against the 2-tuple predict)."""
        if self.cross_emulators is None:
            raise AttributeError("You need to build the emulators before the prediction step.")
        per_model, everything = {}, []
        for uid, emulator in self.cross_emulators.items():
            x_model = self.model_values.model_values[uid]["independent"]
            y_model = self.model_values.model_values[uid]["dependent"]
            valid = np.isin(x_model, emulator.bin_centers)
            emulated, _ = emulator.predict_values(x_model[valid], self.model_parameters.model_parameters[uid])
            weight = y_model[valid] if use_y_as_error else self.model_values.model_values[uid].get(
                "dependent_error", np.ones_like(y_model))[valid]
            difference = (y_model[valid] - emulated) / weight if use_dependent_error else y_model[valid] - emulated
            if use_squared_difference:
                difference = difference ** 2
            per_model[uid] = difference
            everything.extend(difference)
        return np.mean(everything), per_model
