"""Engine DMMA throughput vs cuBLAS DGEMM (torch.matmul fp64) on the GPU box (development tool)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctypes import c_void_p
from emulator_b200 import _lib

lib = _lib.require_gpu()
dev = torch.device("cuda:0")


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts), sorted(ts)[len(ts) // 2]


for (m, n, k) in [(4096, 4096, 4096), (8192, 8192, 8192), (4096, 4096, 128), (8192, 8192, 128)]:
    a = torch.randn(m, k, dtype=torch.float64, device=dev)
    b = torch.randn(n, k, dtype=torch.float64, device=dev)
    c = torch.empty(m, n, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def mine():
        rc = lib.eb200_bench_gemm_nt(m, n, k, c_void_p(a.data_ptr()), c_void_p(b.data_ptr()), c_void_p(c.data_ptr()), c_void_p(st))
        assert rc == 0, _lib.last_error()
    def cublas():
        torch.matmul(a, b.t(), out=c)
    tm, _ = timeit(mine)
    ref = a @ b.t()
    mine(); torch.cuda.synchronize()
    err = float((c - ref).abs().max() / ref.abs().max())
    tc, _ = timeit(cublas)
    fl = 2.0 * m * n * k
    print(json.dumps({"mnk": [m, n, k], "engine_ms": tm, "engine_tflops": fl / tm / 1e9, "cublas_ms": tc,
                      "cublas_tflops": fl / tc / 1e9, "rel_err": err}), flush=True)
