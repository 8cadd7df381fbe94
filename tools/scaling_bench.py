"""Multi-GPU measurements of the two sharded workloads of BASELINE.json (development tool; run under
torchrun, one rank per GPU, or plainly on one GPU):

  C5  Sobol sweep: R parameter rows x 32 independent values predicted (mean + variance) against the
      C2 emulator (N = 4096), rows split contiguously over ranks, one all-gather.
  C4  leave-one-out refits at N = 8176: K refits round-robin over ranks, fitted hyperparameters
      all-gathered (what CrossCheck.build_emulators does for all 512).

Prints one JSON line from rank 0; times are wall clock between barriers (max over ranks by construction)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as td

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
refits = int(sys.argv[2]) if len(sys.argv) > 2 else 8
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
from emulator_b200 import dist, synthetic
from emulator_b200.emulators import GaussianProcessEmulator
from emulator_b200.emulators.base import build_gp, default_kernel
from emulator_b200.backend.model_values import ModelValues
from emulator_b200.sensitivity import sobol

out = {"world": world, "theta_rows": rows, "refits": refits}
# ---- C5 ----
spec, params, values = synthetic.make_config("C2")
emu = GaussianProcessEmulator(device=local)
emu._build_arrays(spec, params, values)
emu.kernel = default_kernel(spec.number_of_parameters + 1)
gp = build_gp(emu.kernel, None, emu.independent_variables, emu.dependent_variables, emu.dependent_variable_errors, device=local)
gp.set_parameter_vector(synthetic.perturbed_parameter_vectors(9, 1)[0])  # same hyperparameters on every rank (no optimisation here)
emu.emulator = gp
design = sobol.saltelli_design(spec, 1 << 10, seed=11)
reps = -(-rows // len(design))
theta = np.tile(design, (reps, 1))[:rows]
ind = np.linspace(0.0, 1.0, 32)
sobol.sweep_predict(emu, theta[: 1 << 10], ind, want_var=True)  # warm-up (factorisation, allocations)
dist.barrier(); t0 = time.perf_counter()
mean, var = sobol.sweep_predict(emu, theta, ind, want_var=True)
dist.barrier(); dt = time.perf_counter() - t0
out["C5_seconds"] = dt
out["C5_query_rows_per_s"] = rows * 32 / dt
out["C5_checksum"] = float(mean.sum() + var.sum())
gp.close()
# ---- C4 ----
spec, params, values = synthetic.make_config("C4")
uids = list(values.model_values.keys())[:refits]
mine = dist.my_indices(refits)
fitted = np.zeros((refits, spec.number_of_parameters + 2))
dist.barrier(); t0 = time.perf_counter()
evals = 0
for i in mine:
    e = GaussianProcessEmulator(device=local)
    e.fit_model(spec, params, ModelValues({k: v for k, v in values.model_values.items() if k != uids[i]}))
    fitted[i] = e.emulator.get_parameter_vector()
    evals += e.emulator.n_device_evaluations
    e.emulator.close()
fitted = dist.all_gather_rows(fitted, mine, refits)
dist.barrier(); dt = time.perf_counter() - t0
out["C4_seconds"] = dt
out["C4_refits_per_s"] = refits / dt
out["C4_evals_this_rank"] = evals
out["C4_checksum"] = float(fitted.sum())
if dist.rank() == 0:
    print(json.dumps(out))
if world > 1:
    td.destroy_process_group()
