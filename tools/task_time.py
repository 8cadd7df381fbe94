"""CTA-time budget per task type from a saved factor trace (development tool)."""
import sys
import numpy as np
f = np.load(sys.argv[1]); out, tasks, nt = f["out"], f["tasks"], int(f["nt"])
t = (out[:, :6].astype(np.int64) - int(out[:, 0].min())) / 1e3
NAMES = {0: "POTRF", 1: "TRTRI", 2: "TRTRI_PART", 3: "POTRF_PART", 4: "XDIAG", 5: "KINV"}
end = t[:, 5].max()
print("kernel end us", end, " CTA-time available (296 CTAs) ms", 296 * end / 1e3)
tot = 0
for ty in range(6):
    m = tasks[:, 0] == ty
    if not m.any(): continue
    i, j, k = tasks[m, 1], tasks[m, 2], tasks[m, 3]
    if ty == 0: prods = j - k
    elif ty == 1: prods = i - np.where(k > j, k, j) + 1  # + the finalisation product
    elif ty == 2: prods = k - (tasks[m, 5] if tasks.shape[1] > 5 else j)
    elif ty == 5: prods = k - tasks[m, 5]
    elif ty == 3: prods = k
    else: prods = np.zeros(m.sum())
    dur = t[m, 5] - t[m, 0]
    tot += dur.sum()
    line = f"{NAMES[ty]:11s} n={m.sum():5d} CTA-time {dur.sum()/1e3:7.1f} ms  products {int(prods.sum()):7d}  us/product {dur.sum()/max(prods.sum(),1):6.2f}"
    if out.shape[1] >= 10:
        line += f" | flag-wait {out[m, 8].sum()/1e6:6.1f} ms yield {out[m, 9].sum()/1e6:6.1f} ms"
    if ty == 1:
        line += f" | accumulate {np.sum(t[m,1]-t[m,0])/1e3:6.1f} park+waitX {np.sum(t[m,2]-t[m,1])/1e3:6.1f} finalize {np.sum(t[m,4]-t[m,2])/1e3:6.1f} publish {np.sum(t[m,5]-t[m,4])/1e3:6.1f}"
    if ty == 0:
        far = (i - j > 2)
        mm = np.where(m)[0][far]
        line += f" | far tiles n={far.sum()}: kgen {np.sum(t[mm,1]-t[mm,0])/1e3:6.1f} accumulate {np.sum(t[mm,2]-t[mm,1])/1e3:6.1f} waitX {np.sum(t[mm,3]-t[mm,2])/1e3:6.1f} finalize {np.sum(t[mm,4]-t[mm,3])/1e3:6.1f} publish {np.sum(t[mm,5]-t[mm,4])/1e3:6.1f}"
    print(line)
print("sum of task time ms", tot / 1e3, " -> fraction of available", tot / (296 * end))
print("64^3 product at DMMA peak (250 GF/s per SM, two CTAs): ", 2 * 64**3 / 125e9 * 1e6, "us per product per CTA")
