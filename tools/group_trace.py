"""CTA-time budget of a GROUPED factorisation launch from in-kernel stamps (development tool):
    python tools/group_trace.py N D NB"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from emulator_b200 import _lib, synthetic
from emulator_b200.device import DeviceProblem
from oracle import gp_oracle as orc

n, d, nb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x, y, yerr = synthetic.design_arrays(n, d, seed=1235)
probs = []
for k in range(nb):
    q = DeviceProblem(x.astype(np.float64), orc.noise_diagonal(yerr), device=0)
    q.set_residual(y.astype(np.float64))
    probs.append(q)
ps = np.ascontiguousarray(synthetic.perturbed_parameter_vectors(d, count=8, seed=7)[1 : 1 + nb])
lib = probs[0]._lib
ntasks = lib.eb200_gp_factor_tasks(probs[0].handle)
out = np.zeros((nb, ntasks, 10), dtype=np.uint64)
tasks = np.zeros((ntasks, 6), dtype=np.int32)
handles = (ctypes.c_void_p * nb)(*[q._h for q in probs])
_lib.check(lib.eb200_gp_debug_group_trace(handles, nb, _lib.as_double_ptr(ps), d + 1, out.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)),
                                          tasks.ctypes.data_as(ctypes.POINTER(ctypes.c_int))), "group_trace")
t0 = int(out[:, :, 0][out[:, :, 0] > 0].min())
t = (out[:, :, :6].astype(np.int64) - t0) / 1e3  # us
end = t[:, :, 5].max()
ctas = len(np.unique(out[:, :, 7]))
print(f"n={n} nb={nb} tasks/problem={ntasks} launch end {end:.1f} us, CTAs seen {ctas}, CTA-time available {ctas * end / 1e3:.1f} ms")
NAMES = {0: "POTRF", 1: "TRTRI", 2: "TRTRI_PART", 3: "POTRF_PART", 4: "XDIAG"}
tot = 0.0
for ty in range(5):
    m = tasks[:, 0] == ty
    if not m.any():
        continue
    i, j, k = tasks[m, 1], tasks[m, 2], tasks[m, 3]
    if ty == 0: prods = j - k + 1
    elif ty == 1: prods = i - np.where(k > j, k, j) + 1
    elif ty == 2: prods = k - tasks[m, 5]
    elif ty == 3: prods = k
    else: prods = np.zeros(m.sum())
    dur = (t[:, m, 5] - t[:, m, 0]).sum()
    wait = out[:, m, 8].sum() / 1e6
    yld = out[:, m, 9].sum() / 1e6
    tot += dur
    line = f"{NAMES[ty]:11s} n={nb * m.sum():6d} CTA-time {dur / 1e3:7.1f} ms products {nb * int(prods.sum()):7d} us/product {dur / max(nb * prods.sum(), 1):6.2f} | flag-wait(thread0) {wait:6.1f} ms yield {yld:5.1f} ms"
    if ty == 1:
        line += f" | accumulate {np.sum(t[:, m, 1] - t[:, m, 0]) / 1e3:6.1f} park+waitX {np.sum(t[:, m, 2] - t[:, m, 1]) / 1e3:6.1f} finalize {np.sum(t[:, m, 4] - t[:, m, 2]) / 1e3:6.1f} publish {np.sum(t[:, m, 5] - t[:, m, 4]) / 1e3:6.1f}"
    if ty == 0:
        far = np.where(m)[0][(i - j > 2)]
        line += f" | far n={nb * len(far)}: kgen {np.sum(t[:, far, 1] - t[:, far, 0]) / 1e3:6.1f} accumulate {np.sum(t[:, far, 2] - t[:, far, 1]) / 1e3:6.1f} waitX {np.sum(t[:, far, 3] - t[:, far, 2]) / 1e3:6.1f} finalize {np.sum(t[:, far, 4] - t[:, far, 3]) / 1e3:6.1f} publish {np.sum(t[:, far, 5] - t[:, far, 4]) / 1e3:6.1f}"
    print(line)
print(f"sum of task time {tot / 1e3:.1f} ms = {tot / (ctas * end):.3f} of the CTA-time available; 64^3 product at DMMA peak, two CTAs per SM: {2 * 64**3 / 125e9 * 1e6:.2f} us")
# per-CTA idle: time between tasks (claiming)
cta_ids = out[:, :, 7].astype(np.int64).ravel()
starts = t[:, :, 0].ravel(); ends = t[:, :, 5].ravel()
gap = 0.0
for c in np.unique(cta_ids):
    sel = cta_ids == c
    o = np.argsort(starts[sel])
    s_, e_ = starts[sel][o], ends[sel][o]
    gap += np.sum(np.maximum(s_[1:] - e_[:-1], 0.0)) + s_[0] + (end - e_[-1])
print(f"CTA-time outside tasks (claiming, start, tail) {gap / 1e3:.1f} ms")
# when each problem's Cholesky / everything finished
for b in range(nb):
    mch = tasks[:, 0] == 0
    print(f"problem {b}: last Cholesky tile published {t[b, mch, 5].max():8.1f} us, last task {t[b, :, 5].max():8.1f} us")
