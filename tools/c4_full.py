"""BASELINE.json configuration C4 for real: CrossCheck over ALL 512 simulations (512 leave-one-out refits at
N = 8176, d = 9) through build_emulators + get_mean_squared, under the handle residency manager; run under torchrun,
one rank per GPU.  Rank 0 prints one JSON line (times are wall clock between barriers = max over ranks)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as td

sims = int(sys.argv[1]) if len(sys.argv) > 1 else 512
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
from emulator_b200 import dist, synthetic
from emulator_b200.george_compat import gp as gp_module
from emulator_b200.sensitivity import CrossCheck

spec, params, values = synthetic.make_config("C4")
uids = list(values.model_values.keys())[:sims]
cc = CrossCheck(leave_out=None if sims >= len(values.model_values) else uids)
dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
cc.build_emulators(spec, params, values)
dist.barrier(); t_build = time.perf_counter() - t0
stats = cc.build_stats
totals = dist.all_reduce_sum(np.array([stats["device_evaluations"], stats["combined_calls"], stats["refits"]], dtype=np.float64))
per_rank = dist.all_reduce_sum(np.eye(world)[dist.rank()] * stats["refits"])
free0 = gp_module._Residency.free_bytes(local)
releases0 = gp_module._residency.releases
dist.barrier(); t0 = time.perf_counter()
total, per_model = cc.get_mean_squared()
dist.barrier(); t_score = time.perf_counter() - t0
out = {
    "world": world, "refits": len(cc.leave_out_order), "n_per_refit": sum(len(v["independent"]) for v in values.model_values.values()) - 16,
    "build_emulators_s": t_build, "refits_per_s": len(cc.leave_out_order) / t_build, "device_evaluations": int(totals[0]),
    "combined_native_calls": int(totals[1]), "refits_per_rank": [int(v) for v in per_rank],
    "get_mean_squared_s": t_score, "mean_squared": float(total),
    "handles_released_during_build_rank0": int(releases0), "handles_released_total_rank0": int(gp_module._residency.releases),
    "free_gib_after_build_rank0": free0 / 2**30, "fit_parameters_checksum": float(cc.cross_fit_parameters.sum()),
    "fit_parameter_spread": [float(v) for v in np.std(cc.cross_fit_parameters, axis=0)],
}
if dist.rank() == 0:
    print(json.dumps(out), flush=True)
for _, e in cc.cross_emulators.built_items():
    e.emulator.close()
if world > 1:
    td.barrier()
    td.destroy_process_group()
