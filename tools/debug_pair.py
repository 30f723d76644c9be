import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')): sys.path.insert(0, p)
import torch
import ase_oracle as O, synth
from ase_b200 import Learner
B, Ba = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 64
P = synth.params(O.ase_param_shapes(), seed=7)
cfg = dict(O.DEFAULT_CFG); cfg['amp_minibatch_size'] = Ba
for pair in ('1', '0'):
    os.environ['ASE_TC_PAIR'] = pair
    st = O.LearnerState(P, 253, 1400, 'ase')
    ln = Learner('ase', 253, 31, B, amp_dim=1400, latent_dim=64, amp_batch=Ba, gemm_backend=2)
    ln.load_named(P)
    for s in range(3):
        d, nz = synth.minibatch(st, cfg, B, Ba, seed=700 + s)
        out = ln.calc_gradients({k: v.cuda() for k, v in d.items()}, nz.cuda())
        res, grads = O.calc_gradients(st, d, cfg, nz)
        torch.cuda.synchronize()
        bad = []
        for k, g in grads.items():
            e = float((ln.named_grads()[k].cpu() - g).abs().max()) / max(float(g.abs().max()), 1e-12)
            if e > 1e-4: bad.append((k, f'{e:.1e}'))
        print(f'pair={pair} step {s}: scalars actor_loss {float(out["scalars"][0]):.6f} vs {float(res["actor_loss"]):.6f}; bad tensors: {bad}')
        ln.adam_step()
        for k, v in ln.named_parameters().items(): v.copy_(st.p[k].cuda().reshape(v.shape))
