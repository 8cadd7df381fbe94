"""Device time of grid predictions (mean + variance) against a C2-size GP for the chunk size / task order given by
EB200_PREDICT_CHUNK / EB200_VAR_GROUP (read once by the library): rows/s and the variance GEMM's TFLOP/s
(N^2 flop per query row).  Development tool.

    python tools/predict_bench.py [N] [D] [N_THETA] [REPS]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from emulator_b200 import synthetic
from emulator_b200.device import DeviceProblem
from oracle import gp_oracle as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 9
n_theta = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
n_ind = 32
dev = torch.device("cuda", 0)
x, y, yerr = synthetic.design_arrays(n, d, seed=1235)
gp = DeviceProblem(x.astype(np.float64), orc.noise_diagonal(yerr), device=0)
gp.set_residual(y.astype(np.float64))
p = synthetic.perturbed_parameter_vectors(d, 1)[1]
gp.evaluate(p, False)
theta = torch.from_numpy(np.random.default_rng(0).random((n_theta, d - 1))).to(dev)
ind = torch.from_numpy(np.linspace(0.0, 1.0, n_ind)).to(dev)
mu = torch.zeros((n_theta, n_ind), dtype=torch.float64, device=dev)
var = torch.zeros_like(mu)
stream = torch.cuda.Stream(device=dev)  # (handle 0 would select the GP handle's own stream)
torch.cuda.set_stream(stream)
out = {"n": n, "d": d, "query_rows": n_theta * n_ind, "chunk": os.environ.get("EB200_PREDICT_CHUNK", "default"),
       "var_group": os.environ.get("EB200_VAR_GROUP", "default")}
for want_var in (True, False):
    gp.predict_grid_device(n_theta, theta.data_ptr(), n_ind, ind.data_ptr(), want_var, mu.data_ptr(), var.data_ptr(), stream=stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        gp.predict_grid_device(n_theta, theta.data_ptr(), n_ind, ind.data_ptr(), want_var, mu.data_ptr(), var.data_ptr(), stream=stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    key = "mean_var" if want_var else "mean_only"
    out[key + "_ms"] = ms
    out[key + "_rows_per_s"] = n_theta * n_ind / ms * 1e3
    if want_var:
        out["var_gemm_tflops_incl_everything"] = float(n) * n * n_theta * n_ind / ms / 1e9
out["checksum"] = float(mu.sum().item() + var.sum().item())
print(json.dumps(out), flush=True)
gp.close()
