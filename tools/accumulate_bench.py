"""Dataflow main loop (64x64 tiles) in isolation vs cuBLAS DGEMM (development tool)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctypes import c_void_p
from emulator_b200 import _lib
lib = _lib.require_gpu()
dev = torch.device("cuda:0")
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts)
m = n = 4096; k = 2048
a = torch.randn(m, k, dtype=torch.float64, device=dev)
bk = torch.randn(n, k, dtype=torch.float64, device=dev)   # KCONTIG B: [n][k]
bm = bk.t().contiguous()                                   # MCONTIG B: [k][n]
c = torch.empty(m, n, dtype=torch.float64, device=dev)
ref = a @ bk.t()
st = torch.cuda.current_stream().cuda_stream
for mc in (0, 1):
    for grid in (148, 296):
        def run():
            rc = lib.eb200_bench_accumulate(m, n, k, mc, grid, c_void_p(a.data_ptr()), c_void_p((bm if mc else bk).data_ptr()), c_void_p(c.data_ptr()), c_void_p(st))
            assert rc == 0, _lib.last_error()
        t = timeit(run)
        err = float((c - ref).abs().max() / ref.abs().max())
        print(json.dumps({"b_layout": "MCONTIG" if mc else "KCONTIG", "grid": grid, "ms": t, "tflops": 2.0 * m * n * k / t / 1e9, "rel_err": err}), flush=True)
tc = timeit(lambda: torch.matmul(a, bk.t(), out=c))
print(json.dumps({"cublas_ms": tc, "cublas_tflops": 2.0 * m * n * k / tc / 1e9}))
