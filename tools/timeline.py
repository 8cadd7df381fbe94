"""Timeline of a saved factor trace: per time bin, CTA-time inside tasks and the part of it spent
waiting for flags (attributed uniformly over each task's lifetime), per task type (development tool)."""
import sys
import numpy as np
f = np.load(sys.argv[1]); out, tasks, nt = f["out"], f["tasks"], int(f["nt"])
t = (out[:, :6].astype(np.int64) - int(out[:, 0].min())) / 1e3
binw = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
end = t[:, 5].max(); nb = int(end // binw) + 1
busy = np.zeros((6, nb)); wait = np.zeros((6, nb))
for k in range(len(tasks)):
    ty = int(tasks[k, 0]); a, b = t[k, 0], t[k, 5]
    if b <= a: continue
    w = out[k, 8] / 1e3 if out.shape[1] >= 10 else 0.0
    for bi in range(int(a // binw), int(b // binw) + 1):
        lo, hi = max(a, bi * binw), min(b, (bi + 1) * binw)
        if hi > lo:
            busy[ty, bi] += hi - lo
            wait[ty, bi] += w * (hi - lo) / (b - a)
print("bin_us   in-task%  waiting%   | computing CTAs by type: POTRF TRTRI TRTRI_PART POTRF_PART XDIAG KINV")
for bi in range(nb):
    cap = 296 * binw
    comp = (busy[:, bi] - wait[:, bi]) / binw
    print(f"{bi*binw:7.0f} {100*busy[:, bi].sum()/cap:8.1f} {100*wait[:, bi].sum()/cap:8.1f}   | " + " ".join(f"{c:6.1f}" for c in comp[:6]))
