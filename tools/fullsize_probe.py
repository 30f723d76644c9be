"""Debug aid: run tests/test_gpu_fullsize._follow for one backend and print, per gradient tensor, how ours and the reference's fp32 result
sit relative to fp64 (scale-factor fit on the large elements, median / max relative error).  python tools/fullsize_probe.py <backend> [nsteps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import torch
import test_gpu_fullsize as T

def three_way(mine, g32, g64, tag):
    print(tag)
    for k in g32:
        a, r, t = mine[k].double().flatten(), g32[k].double().flatten(), g64[k].flatten()
        sc = float(t.abs().max()) + 1e-300
        big = t.abs() > 0.2 * sc
        fit = lambda x: float((x[big] * t[big]).sum() / (t[big] * t[big]).sum()) - 1.0
        print(f"  {k:34s} scale-1: ours {fit(a):+.2e} ref {fit(r):+.2e} | median err ours {float(((a - t).abs() / sc).median()):.2e} ref {float(((r - t).abs() / sc).median()):.2e}"
              f" | mean signed err ours {float(((a - t) / sc).mean()):+.2e} ref {float(((r - t) / sc).mean()):+.2e}")
    return [], 0.0, 0.0
T._three_way = three_way
T._follow('ase', int(sys.argv[1]), int(sys.argv[2]) if len(sys.argv) > 2 else 2, (0, 1))
