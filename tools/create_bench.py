"""Time handle creation / destruction (development tool)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from emulator_b200.device import DeviceProblem
from emulator_b200 import synthetic
from oracle import gp_oracle as orc
for n, d in ((256, 8), (1000, 3), (4096, 9)):
    x, y, yerr = synthetic.design_arrays(n, d, seed=0)
    x = x.astype(np.float64); noise = orc.noise_diagonal(yerr)
    DeviceProblem(x, noise).close()
    t0 = time.perf_counter(); qs = [DeviceProblem(x, noise) for _ in range(16)]; t1 = time.perf_counter()
    for q in qs: q.close()
    t2 = time.perf_counter()
    print(json.dumps({"n": n, "create_ms": (t1 - t0) / 16 * 1e3, "destroy_ms": (t2 - t1) / 16 * 1e3}))
