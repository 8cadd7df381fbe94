"""Timeline analysis of the dataflow Cholesky from in-kernel globaltimer stamps (development tool)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from emulator_b200 import _lib
from emulator_b200.device import DeviceProblem
from oracle import gp_oracle as orc
n, d = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(0)
x = rng.random((n, d)).astype(np.float32).astype(np.float64)
y = np.sin(x.sum(1) * 2.0)
gp = DeviceProblem(x, orc.noise_diagonal(np.full(n, 0.03, dtype=np.float32))); gp.set_residual(y)
p = orc.default_parameter_vector(d)
gp.evaluate(p, True)
nt = gp.padded_n // 64
ntasks = nt * (nt + 1) // 2
out = np.zeros((ntasks, 6), dtype=np.uint64); tiles = np.zeros((ntasks, 2), dtype=np.int32)
for rep in range(2):
    _lib.check(gp._lib.eb200_gp_debug_potrf_trace(gp.handle, _lib.as_double_ptr(p), out.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)),
                                                   tiles.ctypes.data_as(ctypes.POINTER(ctypes.c_int))), "trace")
t = (out.astype(np.int64) - int(out[:, 0].min())) / 1e3  # us
print("total us", t[:, 5].max())
diag = {int(tj): k for k, (ti, tj) in enumerate(tiles) if ti == tj}
sub = {int(tj): k for k, (ti, tj) in enumerate(tiles) if ti == tj + 1}
print("j  diag: claimed kgen_done C_done stored published | sub(j+1,j): kgen_done Ljj_seen stored published")
for j in list(range(0, min(nt, 6))) + list(range(nt // 2, nt // 2 + 3)) + list(range(max(nt - 3, 0), nt)):
    kd = diag[j]; row = f"{j:3d} " + " ".join(f"{v:9.2f}" for v in t[kd, [0, 1, 2, 4, 5]])
    if j in sub:
        ks = sub[j]; row += " | " + " ".join(f"{v:9.2f}" for v in t[ks, [1, 3, 4, 5]])
    print(row)
pub = np.array([t[diag[j], 5] for j in range(nt)])
print("mean diag-to-diag step us:", np.diff(pub).mean())
dd = np.array([t[diag[j], 4] - t[diag[j], 2] for j in range(nt)]); print("diag factor+store us: mean", dd.mean())
kg = (t[:, 1] - t[:, 0]); print("stage X + kgen us mean", kg.mean())
off = tiles[:, 0] != tiles[:, 1]
ts = (t[off, 4] - t[off, 3]); print("trsm(load Ljj+subst+store) us mean", ts.mean())
pb = (t[:, 5] - t[:, 4]); print("publish (fence+flag) us mean", pb.mean())
busy = (t[:, 5] - t[:, 0]).sum(); print("sum task time us", busy, "-> per-SM avg", busy / 148)
