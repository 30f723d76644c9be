"""Micro-benchmark of ase_gemm over the learner's shapes (CUDA events, warm, L2 flushed between reps by a 256 MB write).
  python tools/bench_gemm.py [backend]"""
import os
import sys
import ctypes as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ase_b200 import ops, lib as L

backend = int(sys.argv[1]) if len(sys.argv) > 1 else 1
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device='cuda')


def timeit(M, N, K, a_trans, b_trans, accumulate=False, split_k=0, reps=5, **kw):
    A = torch.randn((K, M) if a_trans else (M, K), device='cuda')
    B = torch.randn((K, N) if b_trans else (N, K), device='cuda')
    out = torch.zeros(M, N, device='cuda')
    bias = torch.randn(N, device='cuda') if kw.get('bias') else None
    ts, kts = [], []
    for r in range(reps + 2):
        flush.zero_()
        L.lib.ase_gemm_tc_profile(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(A, B, a_trans, b_trans, bias, 1 if kw.get('bias') else 0, out=out, accumulate=accumulate, split_k=split_k, backend=backend)
        e1.record()
        torch.cuda.synchronize()
        ms, n, fl = C.c_double(), C.c_int64(), C.c_double()
        L.lib.ase_gemm_tc_profile_read(C.byref(ms), C.byref(n), C.byref(fl))
        if r >= 2:
            ts.append(e0.elapsed_time(e1)); kts.append(ms.value)
    L.lib.ase_gemm_tc_profile(0)
    t, kt = sorted(ts)[len(ts) // 2], sorted(kts)[len(kts) // 2]
    fl = 2.0 * M * N * K
    print(f"M={M:6d} N={N:5d} K={K:6d} at={int(a_trans)} bt={int(b_trans)} acc={int(accumulate)} sk={split_k:2d}  total {t*1e3:8.1f} us  "
          f"main kernel {kt*1e3:8.1f} us  {fl/kt/1e9 if kt else 0:7.1f} TFLOP/s (kernel)  {fl/t/1e9:7.1f} TFLOP/s (with prep)")


print("# forward-like (NT), K sweep at M=32768 N=1024")
for K in (64, 128, 256, 320, 512, 1024, 2048):
    timeit(32768, 1024, K, False, False, bias=True)
print("# M sweep at N=1024 K=1024")
for M in (4096, 8192, 16384, 32768):
    timeit(M, 1024, 1024, False, False, bias=True)
print("# dX-like (B transposed)")
timeit(32768, 1024, 1024, False, True)
timeit(12288, 1400, 1024, False, True)
print("# dW-like (both transposed, split-K accumulate)")
for sk in (1, 2, 5, 10):
    timeit(1024, 1024, 32768, True, True, accumulate=True, split_k=sk)
timeit(1024, 1400, 12288, True, True, accumulate=True, split_k=4)
timeit(512, 1024, 32768, True, True, accumulate=True, split_k=10)
