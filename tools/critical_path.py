"""Critical-path walk over a saved factor trace (gpurun_out/trace_*.npz from tools/factor_trace.py).

Starting from the last diagonal tile, repeatedly step to the input that was published last;
prints, per task on the path, when it was claimed, when its last input arrived, and when it
published (development tool)."""
import sys
import numpy as np

f = np.load(sys.argv[1])
out, tasks, nt = f["out"], f["tasks"], int(f["nt"])
t = (out[:, :6].astype(np.int64) - int(out[:, 0].min())) / 1e3
NAMES = {0: "POTRF", 1: "TRTRI", 2: "TRTRI_PART", 3: "POTRF_PART", 4: "XDIAG"}
NEAR_BAND = 2
idx = {}
for k, r in enumerate(tasks):
    idx[(int(r[0]), int(r[1]), int(r[2]))] = k
pub = lambda key: t[idx[key], 5]


def inputs(k):
    ty, i, j, kk = (int(v) for v in tasks[k][:4])
    res = []
    if ty == 0:      # L_ij: needs L_ik, L_jk for k in [kbeg, j), the partial, and the diagonal tile
        for c in range(kk, j):
            res.append((0, i, c))
            if i != j:
                res.append((0, j, c))
        if kk > 0:
            res.append((3, i, j))
        if i != j:
            res.append((0, j, j) if (i - j <= NEAR_BAND and i < nt) else (4, j, j))
    elif ty == 4:
        res.append((0, j, j))
    elif ty == 3:    # partial: columns [0, kk)
        for c in range(kk):
            res.append((0, i, c))
            if i != j:
                res.append((0, j, c))
    return [key for key in res if key in idx]


k = idx[(0, nt - 1, nt - 1)]
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 60
print("task                claimed   last-input  C_done    published   (own = published - max(claimed, last input))")
path = []
while True:
    ins = inputs(k)
    if not ins:
        break
    last = max(ins, key=pub)
    ty, i, j, kk = (int(v) for v in tasks[k][:4])
    own = t[k, 5] - max(t[k, 0], pub(last))
    path.append((NAMES[ty], i, j, t[k, 0], pub(last), t[k, 2], t[k, 5], own, last))
    k = idx[last]
for row in path[:limit]:
    print(f"{row[0]:11s}({row[1]:2d},{row[2]:2d}) {row[3]:9.2f} {row[4]:9.2f} {row[5]:9.2f} {row[6]:9.2f}   own {row[7]:6.2f}   <- {NAMES[row[8][0]]}{row[8][1:]}")
kinds = {}
for row in path:
    key = (row[0], "diag" if row[1] == row[2] else f"off{min(row[1] - row[2], 4)}")
    kinds.setdefault(key, []).append(row[7])
print("critical path length", len(path), "tasks; own-time by kind:")
for key, v in sorted(kinds.items()):
    print(f"  {key}: n={len(v)} sum={sum(v):8.1f} us  mean={np.mean(v):6.2f}")

# optional: explain one tile -- python critical_path.py trace.npz 0 I J
if len(sys.argv) > 4:
    def explain(key, depth, indent=""):
        k = idx[key]
        ins = sorted(inputs(k), key=pub, reverse=True)[:3]
        print(f"{indent}{NAMES[key[0]]}{key[1:]} claimed {t[k,0]:.2f} stamps {t[k,1]:.2f} {t[k,2]:.2f} {t[k,3]:.2f} {t[k,4]:.2f} published {t[k,5]:.2f} sm {int(out[k,6])} cta {int(out[k,7])}")
        if depth > 0:
            for key2 in ins:
                explain(key2, depth - 1, indent + "    ")
    explain((0, int(sys.argv[3]), int(sys.argv[4])), int(sys.argv[5]) if len(sys.argv) > 5 else 2)
