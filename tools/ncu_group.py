"""Two grouped LML+grad evaluations of eight C2-size handles and nothing else: the target of `ncu` captures of the
grouped kernels (development tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from emulator_b200 import synthetic
from emulator_b200.device import DeviceProblem, evaluate_many
from oracle import gp_oracle as orc

n, d, nh = 4096, 9, 8
x, y, yerr = synthetic.design_arrays(n, d, seed=1235)
probs = []
for k in range(nh):
    q = DeviceProblem(x.astype(np.float64), orc.noise_diagonal(yerr), device=0)
    q.set_residual(y.astype(np.float64))
    probs.append(q)
ps = synthetic.perturbed_parameter_vectors(d, count=16, seed=7)[1:]
for rep in range(2):
    lml, grad, info = evaluate_many(probs, ps[rep * nh : (rep + 1) * nh], want_grad=True)
print(lml, info)
