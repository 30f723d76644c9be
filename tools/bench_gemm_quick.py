import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from ase_b200 import ops, lib as L
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device='cuda')
for (M, N, K) in ((32768, 1024, 64), (32768, 1024, 1024)):
    A = torch.randn(M, K, device='cuda'); B = torch.randn(N, K, device='cuda'); out = torch.zeros(M, N, device='cuda'); bias = torch.randn(N, device='cuda')
    kts = []
    for r in range(5):
        flush.zero_(); L.lib.ase_gemm_tc_profile(1)
        ops.gemm(A, B, False, False, bias, 1, out=out, backend=1); torch.cuda.synchronize()
        ms, n, fl = C.c_double(), C.c_int64(), C.c_double(); L.lib.ase_gemm_tc_profile_read(C.byref(ms), C.byref(n), C.byref(fl)); kts.append(ms.value)
    print(f"debug={os.environ.get('ASE_TC_DEBUG','0')} M={M} N={N} K={K}: main kernel {sorted(kts)[2]*1e3:.1f} us")
