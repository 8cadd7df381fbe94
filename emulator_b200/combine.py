"""
Evaluation combiner: several host threads, each driving its own optimiser on its own GP, share device launches.

The reference fits independent GPs one after the other -- per-bin emulators
(/root/reference/swiftemulator/emulators/gaussian_process_bins.py:206-262), leave-one-out refits
(/root/reference/swiftemulator/sensitivity/cross_check.py:98-114) -- each through its own
`scipy.optimize.minimize` call, whose `fun(p)` / `jac(p)` callbacks block on one likelihood evaluation at a
time.  One factorisation cannot keep a B200 busy until it is several thousand rows large (its Cholesky
dependency chain feeds the SMs too slowly), several interleaved ones can (`eb200_gp_eval_multi_host`).

So the optimisers stay scipy's, unmodified and one per problem -- same iterates, same optima as a sequential
run, bit for bit, because a grouped evaluation returns exactly what a stand-alone one does -- and only their
evaluations are combined: a thread that needs a value parks its request here; when every active thread has
one parked, the last to arrive issues ONE native call for all of them and hands the results back.
"""

from __future__ import annotations

import threading
import time
from typing import List, Optional

import numpy as np

from .device import evaluate_many

_tls = threading.local()


def current() -> Optional["EvaluationCombiner"]:
    """The combiner the calling thread works under (None: evaluations go to the device one by one)."""
    return getattr(_tls, "combiner", None)


class EvaluationCombiner:
    def __init__(self, participants: int):
        if participants < 1:
            raise ValueError("a combiner needs at least one participating thread")
        self._cond = threading.Condition()
        self._active = participants
        self._pending: List[dict] = []
        self.calls = 0        # native calls issued
        self.evaluations = 0  # evaluations they carried
        self.seconds = 0.0    # wall time spent inside those calls (device work + launch + synchronisation)

    # ---- thread membership ------------------------------------------------------------------------
    def __enter__(self):
        """Entered by each participating thread; evaluations of GPs used inside go through the combiner."""
        _tls.combiner = self
        return self

    def __exit__(self, *exc):
        _tls.combiner = None
        self.retire()
        return False

    def retire(self):
        """The calling thread will submit no more requests (the others no longer wait for it)."""
        with self._cond:
            self._active -= 1
            batch = self._take_if_complete()
        if batch:
            self._run(batch)

    # ---- evaluation -----------------------------------------------------------------------------------
    def evaluate(self, problem, p: np.ndarray, want_grad: bool):
        """(lml, grad, info) of `problem` (a DeviceProblem, pinned by the caller) at p; blocks until a combined
        call that includes this request has run."""
        req = {"problem": problem, "p": np.array(p, dtype=np.float64), "grad": bool(want_grad), "done": False, "out": None,
               "err": None}
        with self._cond:
            self._pending.append(req)
            batch = self._take_if_complete()
        if batch:
            self._run(batch)
        with self._cond:
            while not req["done"]:
                self._cond.wait()
        if req["err"] is not None:
            raise req["err"]
        return req["out"]

    def _take_if_complete(self):
        # caller holds the lock
        if self._pending and len(self._pending) >= self._active:
            batch, self._pending = self._pending, []
            return batch
        return None

    def _run(self, batch):
        try:
            # one native call per evaluation mode and kernel dimension (in practice: one)
            keys = sorted({(r["grad"], r["problem"].d) for r in batch})
            for want_grad, d in keys:
                part = [r for r in batch if r["grad"] == want_grad and r["problem"].d == d]
                t0 = time.perf_counter()
                lml, grad, info = evaluate_many([r["problem"] for r in part], np.array([r["p"] for r in part]), want_grad=want_grad)
                self.seconds += time.perf_counter() - t0
                for k, r in enumerate(part):
                    r["out"] = (float(lml[k]), grad[k].copy(), int(info[k]))
                self.calls += 1
                self.evaluations += len(part)
        except BaseException as exc:  # hand the failure to every waiting thread instead of leaving them parked
            for r in batch:
                if r["out"] is None:
                    r["err"] = exc
        with self._cond:
            for r in batch:
                r["done"] = True
            self._cond.notify_all()
