"""Per-tensor gradient errors of the HLC learner (BASELINE config 5, tests/golden/calc_grad_hrl_small.pt) for the three GEMM backends:
ours vs the reference's fp32 gradients, ours vs the oracle in fp64, the reference vs fp64 (all relative to max|g| of the tensor)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import torch
import ase_oracle as O, synth, golden_util as G
from ase_b200 import Learner
fx = G.load('calc_grad_hrl_small.pt'); meta = fx['meta']
for backend in (0, 1, 2):
    P = synth.params(O.amp_param_shapes(obs=258, act=64, amp=0, units=meta['units']), seed=meta['seed'])
    st = O.LearnerState(P, 258, 0, 'ppo'); st64 = O.LearnerState({k: v.double() for k, v in P.items()}, 258, 0, 'ppo')
    ln = Learner('ppo', 258, 64, meta['B'], units=tuple(meta['units']), hparams={'learning_rate': meta['cfg']['lr']}, gemm_backend=backend, mu_activation='tanh')
    ln.load_named(P)
    for s, rec in enumerate(fx['steps']):
        d, _ = synth.minibatch(st, meta['cfg'], meta['B'], 0, seed=meta['seed'] * 100 + s, kind='ppo', obs_dim=258, act=64)
        out = ln.calc_gradients({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()})
        for k in st.p:
            st64.p[k] = st.p[k].double()
        d64 = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()}
        _, g64 = O.calc_gradients(st64, d64, meta['cfg'], None, apply_adam=False)
        for k, g in rec['grads'].items():
            mine = ln.named_grads()[k].cpu(); sc = max(float(g.abs().max()), 1e-30)
            e32 = float((mine - g).abs().max()) / sc; em = float((mine.double() - g64[k]).abs().max()) / sc; er = float((g.double() - g64[k]).abs().max()) / sc
            q = float(((mine.double() - g64[k]).abs() / sc).flatten().quantile(0.99)) if mine.numel() < 2 ** 24 else -1
            print(f"backend {backend} step {s} {k:28s} ours-ref32 {e32:.2e}  ours-fp64 {em:.2e} (q99 {q:.1e})  ref32-fp64 {er:.2e}")
        ln.adam_step()
        O.calc_gradients(st, d, meta['cfg'], None)
        for k, v in ln.named_parameters().items():
            v.copy_(st.p[k].to(v.device).reshape(v.shape))
        ln.params_changed()
