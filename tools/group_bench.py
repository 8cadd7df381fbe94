"""Grouped multi-handle evaluations (shared factorisation launches) against stand-alone evaluations of the
same handles: device time per evaluation and bit-for-bit comparison (development tool).

    python tools/group_bench.py N D HANDLES REPS

Group size and chain CTAs per problem come from EB200_GROUP_SIZE / EB200_GROUP_CHAIN (read once by the library)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from emulator_b200 import synthetic
from emulator_b200.device import DeviceProblem, evaluate_many, evaluate_many_device
from oracle import gp_oracle as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 9
nh = int(sys.argv[3]) if len(sys.argv) > 3 else 8
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dev = torch.device("cuda", 0)
x, y, yerr = synthetic.design_arrays(n, d, seed=1235)
x, y = x.astype(np.float64), y.astype(np.float64)
noise = orc.noise_diagonal(yerr)
probs = []
for k in range(nh):
    q = DeviceProblem(x, noise, device=0)
    q.set_residual(y)
    probs.append(q)
ps = synthetic.perturbed_parameter_vectors(d, count=max(32, nh), seed=7)
P, L = d + 1, d + 5
p_dev = torch.from_numpy(ps).to(dev)
out_single = torch.zeros((len(ps), L), dtype=torch.float64, device=dev)
out_group = torch.zeros((len(ps), L), dtype=torch.float64, device=dev)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
sp = stream.cuda_stream


def run_single(count):
    for i in range(count):
        probs[i % nh].evaluate_device(p_dev[i % len(ps)].data_ptr(), True, out_single[i % len(ps)].data_ptr(), sp)


# BENCH_STREAMS=2: consecutive groups go to alternating streams (the tail of one group's factorisation -- CTAs
# leaving the persistent kernel -- then overlaps the starved start of the next group's)
n_streams = int(os.environ.get("BENCH_STREAMS", "1"))
gsize = int(os.environ.get("EB200_GROUP_SIZE", "4" if n > 2048 else "8"))
extra_streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams - 1)]
streams = [stream] + extra_streams


def run_group(count):
    if n_streams == 1:
        for i0 in range(0, count, nh):
            k = i0 % len(ps)
            evaluate_many_device(probs, p_dev[k].data_ptr(), P, True, out_group[k].data_ptr(), L, sp)
        return
    for s_ in extra_streams:
        s_.wait_stream(stream)
    g = 0
    for i0 in range(0, count, nh):
        k = i0 % len(ps)
        for h0 in range(0, nh, gsize):
            st = streams[g % n_streams]
            g += 1
            evaluate_many_device(probs[h0 : h0 + gsize], p_dev[k + h0].data_ptr(), P, True, out_group[k + h0].data_ptr(), L, st.cuda_stream)
    for s_ in extra_streams:
        stream.wait_stream(s_)


def timed(fn, count):
    fn(count)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn(count)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * count)


count = (len(ps) // nh) * nh
res = {"n": n, "d": d, "handles": nh, "group_size": os.environ.get("EB200_GROUP_SIZE", "8"),
       "chain": os.environ.get("EB200_GROUP_CHAIN", "default"), "streams": n_streams}
res["single_ms_per_eval"] = timed(run_single, count)
res["group_ms_per_eval"] = timed(run_group, count)
a, b = out_single.cpu().numpy()[:count], out_group.cpu().numpy()[:count]
res["bit_identical"] = bool(np.array_equal(a, b))
res["max_abs_diff"] = float(np.max(np.abs(a - b)))
res["info_max"] = float(np.max(np.abs(a[:, d + 2])))
# host path (pinned copies inside): one native call per group
import time

lml, grad, info = evaluate_many(probs, ps[:nh], want_grad=True)
t0 = time.perf_counter()
for r in range(reps):
    lml, grad, info = evaluate_many(probs, ps[:nh], want_grad=True)
res["group_host_ms_per_eval"] = (time.perf_counter() - t0) * 1e3 / (reps * nh)
res["host_equals_device"] = bool(np.array_equal(lml, b[:nh, 0]) and np.array_equal(grad, b[:nh, 1 : P + 1]))
# reference values
ref_lml, ref_grad, _, _ = orc.lml_and_grad_lean(ps[0], x, noise, y) if n <= 4096 else (None, None, None, None)
if ref_lml is not None:
    res["oracle_rel_lml"] = float(abs(b[0, 0] - ref_lml) / abs(ref_lml))
    res["oracle_rel_grad"] = float(np.max(np.abs(b[0, 1 : P + 1] - ref_grad)) / np.max(np.abs(ref_grad)))
print(json.dumps(res), flush=True)
for q in probs:
    q.close()
