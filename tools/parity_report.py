"""Prints the parity table (GPU engine vs the reference's golden outputs) for the three GEMM backends.
Usage (on the GPU box): python tools/parity_report.py > profiles/parity_rNN.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import ase_oracle as O
import synth
import golden_util as G
from test_gpu_learner import _make_learner, _cuda


def report(name, backend):
    meta, steps, shapes, P = G.calc_grad_case(name)
    kind = meta['kind']
    ln = _make_learner(kind, meta, P, backend)
    st = O.LearnerState(P, 253, 1400, kind)
    cfg = meta['cfg']
    print(f"== {name}  backend={('SIMT-fp32', 'tcgen05-3xTF32', 'tcgen05-3xFP16-scaled')[backend]}  B={meta['B']} Ba={meta['Ba']}")
    for s, rec in enumerate(steps):
        d, nz = synth.minibatch(st, cfg, meta['B'], meta['Ba'], seed=meta['seed'] * 100 + s, kind=kind)
        out = ln.calc_gradients(_cuda(d), None if nz is None else nz.cuda())
        tr = ln.train_result(out)
        print(f" step {s}: scalar            reference        ours             rel.err")
        for k, v in rec['scalars'].items():
            if k in tr:
                print(f"   {k:22s} {v:+.8e} {tr[k]:+.8e} {abs(tr[k]-v)/max(abs(v),1e-12):.2e}")
        worst = (0, '')
        for k, gv in ln.named_grads().items():
            g = gv.detach().cpu().flatten()
            idx = G.sample_idx(g.numel())
            ref = rec['grad_sample'][k]
            e = float((g[idx] - ref).abs().max()) / max(float(ref.abs().max()), 1e-30)
            ne = abs(float(g.double().norm()) - rec['grad_norm'][k]) / max(rec['grad_norm'][k], 1e-30)
            worst = max(worst, (e, k))
            print(f"   grad {k:40s} max|d|/max|ref| {e:.2e}   norm rel.err {ne:.2e}")
        print(f"   worst gradient tensor: {worst[1]} {worst[0]:.2e}")
        ln.adam_step()
        pw = 0
        for k, pv in ln.named_parameters().items():
            p = pv.detach().cpu().flatten(); idx = G.sample_idx(p.numel())
            pw = max(pw, float((p[idx] - rec['param_sample'][k]).abs().max()))
        print(f"   post-Adam params: max abs diff (sampled) {pw:.2e}")
        O.calc_gradients(st, d, cfg, nz)


if __name__ == '__main__':
    for name in ('calc_grad_ase_cfg1.pt', 'calc_grad_amp_cfg.pt'):
        for backend in (0, 1, 2):
            report(name, backend)
