"""Per-stage device time of one LML+grad evaluation at (n, d) (development tool)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from emulator_b200.device import DeviceProblem
from emulator_b200 import synthetic
from oracle import gp_oracle as orc
for arg in sys.argv[1:]:
    n, d = (int(v) for v in arg.split(":"))
    x, y, yerr = synthetic.design_arrays(n, d, seed=n)
    gp = DeviceProblem(x.astype(np.float64), orc.noise_diagonal(yerr)); gp.set_residual(y.astype(np.float64))
    p = orc.default_parameter_vector(d)
    st = gp.profile_stages(p, reps=20)
    print(json.dumps({"n": n, "d": d, "stage_ms [params, factor, alpha, kinv+traces, finalize]": [round(float(v), 4) for v in st], "sum": round(float(st.sum()), 4)}))
    gp.close()
