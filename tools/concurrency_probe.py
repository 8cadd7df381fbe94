"""Throughput with R replica handles evaluating independent hyperparameter vectors concurrently (development tool)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from emulator_b200 import synthetic
from emulator_b200.device import DeviceProblem
n, d = 4096, 9
x, y, yerr = synthetic.design_arrays(n, d, seed=1235)
noise = np.sqrt(np.ascontiguousarray(yerr ** 2, dtype=np.float64) + 1.25e-12) ** 2
ps = synthetic.perturbed_parameter_vectors(d, 32)
dev = torch.device("cuda:0")
p_dev = torch.from_numpy(ps).to(dev)
for R in (1, 2, 3, 4):
    gps = [DeviceProblem(x.astype(np.float64), noise) for _ in range(R)]
    for g in gps: g.set_residual(y.astype(np.float64))
    streams = [torch.cuda.Stream() for _ in range(R)]
    out = torch.zeros((len(ps), d + 5), dtype=torch.float64, device=dev)
    main = torch.cuda.current_stream()
    def run(nev):
        for i in range(nev):
            r = i % R
            gps[r].evaluate_device(p_dev[i % len(ps)].data_ptr(), True, out[i % len(ps)].data_ptr(), streams[r].cuda_stream)
    run(2 * R); torch.cuda.synchronize()
    nev = 48
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for s in streams: s.wait_stream(main)
    run(nev)
    for s in streams: main.wait_stream(s)
    e1.record(main); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"replicas": R, "evals": nev, "ms_per_eval": ms / nev, "evals_per_s": nev / ms * 1e3, "info_ok": bool((out[:, d + 2] == 0).all())}), flush=True)
    for g in gps: g.close()
