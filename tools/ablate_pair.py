"""Mainloop ablation of the CTA-pair GEMM: time per k-block round from the slope of a K sweep, under ASE_TC_DEBUG bits
(1 no store, 2 no TMEM drain loads, 4 no correction MMAs, 8 no L2 prefetch, 16 loads always hit the same L2-resident boxes,
32 no loads at all (MMAs on stale shared memory), 64 no main MMAs either).  Results are garbage under any bit; only the times count."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from ase_b200 import ops, lib as L
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device='cuda')
res = {}
for (M, N, K) in ((32768, 1024, 1024), (32768, 1024, 2048)):
    A = torch.randn(M, K, device='cuda'); B = torch.randn(N, K, device='cuda'); out = torch.zeros(M, N, device='cuda'); bias = torch.randn(N, device='cuda')
    kts = []
    for r in range(5):
        flush.zero_(); L.lib.ase_gemm_tc_profile(1)
        ops.gemm(A, B, False, False, bias, 1, out=out, backend=2); torch.cuda.synchronize()
        ms, n, fl = C.c_double(), C.c_int64(), C.c_double(); L.lib.ase_gemm_tc_profile_read(C.byref(ms), C.byref(n), C.byref(fl)); kts.append(ms.value)
    res[K] = sorted(kts)[2] * 1e3
# 32768x1024 = 512 pair tiles over 74 pairs: 7 rounds for the busiest pair; (2048 - 1024) / 64 = 16 k-blocks per tile
per_kb = (res[2048] - res[1024]) / (7 * 16)
print(f"debug={os.environ.get('ASE_TC_DEBUG', '0'):>3}  K=1024 {res[1024]:7.1f} us   K=2048 {res[2048]:7.1f} us   per k-block {per_kb:.3f} us")
