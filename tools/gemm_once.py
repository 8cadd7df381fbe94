"""One engine GEMM (profiling target, development tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctypes import c_void_p
from emulator_b200 import _lib
lib = _lib.require_gpu()
dev = torch.device("cuda:0")
m = n = k = 4096
a = torch.randn(m, k, dtype=torch.float64, device=dev); b = torch.randn(n, k, dtype=torch.float64, device=dev)
c = torch.empty(m, n, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    assert lib.eb200_bench_gemm_nt(m, n, k, c_void_p(a.data_ptr()), c_void_p(b.data_ptr()), c_void_p(c.data_ptr()), c_void_p(st)) == 0
torch.cuda.synchronize()
print("ok")
