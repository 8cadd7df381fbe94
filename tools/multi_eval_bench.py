"""Time eb200_gp_eval_multi_host over nb equally shaped small problems (development tool)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from emulator_b200.device import DeviceProblem, evaluate_many
from emulator_b200 import synthetic
from oracle import gp_oracle as orc
n, d = int(sys.argv[1]), int(sys.argv[2])
probs = []
for k in range(64):
    x, y, yerr = synthetic.design_arrays(n, d, seed=k)
    q = DeviceProblem(x.astype(np.float64), orc.noise_diagonal(yerr)); q.set_residual(y.astype(np.float64)); probs.append(q)
p = np.tile(orc.default_parameter_vector(d), (64, 1))
for nb in (1, 8, 16, 32, 64):
    evaluate_many(probs[:nb], p[:nb])
    t0 = time.perf_counter()
    for _ in range(50): evaluate_many(probs[:nb], p[:nb])
    dt = (time.perf_counter() - t0) / 50
    print(json.dumps({"n": n, "nb": nb, "ms_per_call": dt * 1e3, "us_per_eval": dt * 1e6 / nb}))
