"""From a saved full-mode factor trace: when does each row of the inverse finish relative to the
diagonal tile of the same row (development tool)."""
import sys
import numpy as np
f = np.load(sys.argv[1]); out, tasks, nt = f["out"], f["tasks"], int(f["nt"])
t = (out[:, :6].astype(np.int64) - int(out[:, 0].min())) / 1e3
T = {int(r[1]): t[k, 5] for k, r in enumerate(tasks) if r[0] == 0 and r[1] == r[2]}
print(" r   T_r    Lrow_done-T_r   Xrow_first_start-T_r  Xrow_done-T_r   busy CTAs at T_r")
busy_at = lambda x: int(np.sum((t[:, 0] <= x) & (t[:, 5] >= x)))
for r in range(1, nt):
    xs = [k for k, q in enumerate(tasks) if q[0] == 1 and q[1] == r]
    ls = [k for k, q in enumerate(tasks) if q[0] == 0 and q[1] == r and q[2] < r]
    print(f"{r:3d} {T[r]:8.1f} {max(t[k,5] for k in ls)-T[r]:10.1f} {min(t[k,0] for k in xs)-T[r]:14.1f} {max(t[k,5] for k in xs)-T[r]:14.1f} {busy_at(T[r]):8d}")
print("kernel end", t[:, 5].max())
