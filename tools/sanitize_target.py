"""Small end-to-end run (eval + predict) used as the compute-sanitizer target (development tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from emulator_b200.device import DeviceProblem
from emulator_b200 import synthetic
from oracle import gp_oracle as orc
for n, d in ((200, 3), (1500, 9)):
    x, y, yerr = synthetic.design_arrays(n, d, seed=3)
    x, y = x.astype(np.float64), y.astype(np.float64)
    diag = orc.noise_diagonal(yerr)
    gp = DeviceProblem(x, diag); gp.set_residual(y)
    p = orc.default_parameter_vector(d) + 0.1
    out = gp.evaluate(p, True)
    mu, var = gp.predict(x[:150] + 0.01, True)
    ref = orc.lml_and_grad_lean(p, x, diag, y)
    print(n, d, out[2], abs(out[0] - ref[0]) / abs(ref[0]), float(np.max(np.abs(out[1] - ref[1])) / np.max(np.abs(ref[1]))))
    gp.close()
