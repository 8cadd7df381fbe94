"""Timeline analysis of the dataflow factorisation kernel from in-kernel globaltimer stamps (development tool)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from emulator_b200 import _lib
from emulator_b200.device import DeviceProblem
from oracle import gp_oracle as orc
n, d = int(sys.argv[1]), int(sys.argv[2])
value_only = len(sys.argv) > 3 and sys.argv[3] == "value"
rng = np.random.default_rng(0)
x = rng.random((n, d)).astype(np.float32).astype(np.float64)
y = np.sin(x.sum(1) * 2.0)
gp = DeviceProblem(x, orc.noise_diagonal(np.full(n, 0.03, dtype=np.float32))); gp.set_residual(y)
p = orc.default_parameter_vector(d)
gp.evaluate(p, True)
nt = gp.padded_n // 64
ntasks = (gp._lib.eb200_gp_value_tasks if value_only else gp._lib.eb200_gp_factor_tasks)(gp.handle)
trace_fn = gp._lib.eb200_gp_debug_value_trace if value_only else gp._lib.eb200_gp_debug_factor_trace
out = np.zeros((ntasks, 10), dtype=np.uint64); tasks = np.zeros((ntasks, 6), dtype=np.int32)
for rep in range(2):
    _lib.check(trace_fn(gp.handle, _lib.as_double_ptr(p), out.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)),
                                                    tasks.ctypes.data_as(ctypes.POINTER(ctypes.c_int))), "trace")
os.makedirs("gpurun_out", exist_ok=True)
np.savez(f"gpurun_out/trace_{n}_{'value' if value_only else 'full'}.npz", out=out, tasks=tasks, nt=nt)
sm_of = out[:, 6].astype(np.int64); cta_of = out[:, 7].astype(np.int64)
t = (out[:, :6].astype(np.int64) - int(out[:, 0].min())) / 1e3  # us
print("ntasks", ntasks, "total us", t[:, 5].max())
po = tasks[:, 0] == 0
diag = {int(r[2]): k for k, r in enumerate(tasks) if r[0] == 0 and r[1] == r[2]}
sub = {int(r[2]): k for k, r in enumerate(tasks) if r[0] == 0 and r[1] == r[2] + 1 and r[1] < nt}
print("j  diag: claimed kgen_done C_done stored published | sub(j+1,j): claimed C_done Ljj_seen stored published")
for j in list(range(0, min(nt, 3))) + list(range(nt // 2, min(nt, nt // 2 + 12))) + list(range(max(nt - 3, 0), nt)):
    kd = diag[j]; row = f"{j:3d} " + " ".join(f"{v:9.2f}" for v in t[kd, [0, 1, 2, 4, 5]])
    if j in sub:
        ks = sub[j]; row += " | " + " ".join(f"{v:9.2f}" for v in t[ks, [0, 2, 3, 4, 5]])
    print(row)
pub = np.array([t[diag[j], 5] for j in range(nt)])
steps = np.diff(pub)
print("step us percentiles 10/50/90/max:", np.percentile(steps, [10, 50, 90, 100]))
# what the diagonal task of step j+1 waited for: its C_done minus the sub-diagonal publish of step j
lag_d = np.array([t[diag[j + 1], 2] - t[sub[j], 5] for j in range(nt - 1)])
lag_s = np.array([t[sub[j], 3] - t[diag[j], 5] for j in range(nt - 1) if j in sub])
print("diag C_done - prev sub published (last-panel latency) 10/50/90/max:", np.percentile(lag_d, [10, 50, 90, 100]))
print("sub Ljj_seen - diag published 10/50/90/max:", np.percentile(lag_s, [10, 50, 90, 100]))
late_s = np.array([t[sub[j], 2] - t[diag[j], 5] for j in range(nt - 1) if j in sub])
print("sub C_done - diag published (positive = sub tile's own accumulation finished late) 10/50/90/max:", np.percentile(late_s, [10, 50, 90, 100]))
print("cholesky finished at us", pub.max(), " mean diag-to-diag step us:", np.diff(pub).mean())
dd = np.array([t[diag[j], 4] - t[diag[j], 2] for j in range(nt)]); print("diag factor+store us: mean", dd.mean())
subk = [sub[j] for j in sorted(sub)]
print("chain sub tiles: Ljj_seen->stored (trsm) 10/50/90:", np.percentile(t[subk, 4] - t[subk, 3], [10, 50, 90]), " publish:", np.median(t[subk, 5] - t[subk, 4]))
dk = [diag[j] for j in sorted(diag)]
print("chain diag tiles: C_done->stored (chol) 10/50/90:", np.percentile(t[dk, 4] - t[dk, 2], [10, 50, 90]), " publish:", np.median(t[dk, 5] - t[dk, 4]))
second = {int(r[2]): k for k, r in enumerate(tasks) if r[0] == 0 and r[1] == r[2] + 2 and r[1] < nt}
lat2 = np.array([t[second[j], 5] - t[diag[j], 5] for j in second])
print("second sub-diagonal tile published - diag published 10/50/90/max:", np.percentile(lat2, [10, 50, 90, 100]))
off = po & (tasks[:, 1] != tasks[:, 2])
print("trsm(load Ljj+subst+store) us mean", (t[off, 4] - t[off, 3]).mean())
tr = ~po
if tr.any():
    print("trtri tasks:", tr.sum(), " accumulate us mean", (t[tr, 1] - t[tr, 0]).mean(), " wait Xrr us mean", (t[tr, 2] - t[tr, 1]).mean(),
          " finalize us mean", (t[tr, 4] - t[tr, 2]).mean(), " first start", t[tr, 0].min(), " last end", t[tr, 5].max())
busy = (t[:, 5] - t[:, 0]).sum(); print("sum task time us", busy, "-> per-SM avg", busy / 148)

chain = [k for k, r in enumerate(tasks) if r[0] == 0 and r[1] - r[2] <= 1 and r[1] < nt]
print("chain tasks ran on SMs", sorted(set(sm_of[chain].tolist())), "CTAs", sorted(set(cta_of[chain].tolist())))
pairs = {}
for k in range(ntasks): pairs.setdefault(int(cta_of[k]), set()).add(int(sm_of[k]))
same = sum(1 for b in range(148) if b in pairs and (b + 148) in pairs and pairs[b] == pairs[b + 148])
print("CTA b and b+148 on the same SM for", same, "of 148 pairs; distinct SMs used", len(set(sm_of.tolist())))
