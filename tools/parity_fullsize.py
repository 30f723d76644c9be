"""Full-size (B = 16384, B_amp = 4096) parity diagnostic: for each GEMM backend, per-tensor relative error (to the tensor's max) of the
gradients against (a) the fp32 CPU oracle (= the reference's own arithmetic, pinned by tests/golden/calc_grad_*_full.pt) and (b) the SAME
oracle evaluated in fp64 ("truth"), next to the fp32 oracle's own distance from fp64.  If the reference's fp32 result is as far from
fp64 as ours, the difference between the two is conditioning (clip / ReLU decisions flipping on ~1 of 16384 samples), not a defect.
   python tools/parity_fullsize.py [ase|amp] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import torch
import ase_oracle as O
import synth
from ase_b200 import Learner

kind = sys.argv[1] if len(sys.argv) > 1 else 'ase'
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.set_num_threads(min(32, len(os.sched_getaffinity(0))))
B, Ba = 16384, 4096
shapes = O.ase_param_shapes() if kind == 'ase' else O.amp_param_shapes()
units = (1024, 1024, 512) if kind == 'ase' else (1024, 512)
P = synth.params(shapes, seed=23)
cfg = dict(O.DEFAULT_CFG); cfg['amp_minibatch_size'] = Ba
if kind == 'amp':
    cfg['enc_coef'] = 0.0; cfg['amp_diversity_bonus'] = 0.0
hp = {k: cfg[k] for k in ('e_clip', 'critic_coef', 'entropy_coef', 'bounds_loss_coef', 'disc_coef', 'disc_logit_reg', 'disc_grad_penalty',
                          'disc_weight_decay', 'enc_coef', 'amp_diversity_bonus', 'amp_diversity_tar')}
hp['learning_rate'] = cfg['lr']


def to64(x):
    return x.double() if torch.is_tensor(x) and x.is_floating_point() else x


st32 = O.LearnerState(P, 253, 1400, kind)
st64 = O.LearnerState({k: v.double() for k, v in P.items()}, 253, 1400, kind)
lns = {}
for be in (0, 1, 2):
    lns[be] = Learner(kind, 253, 31, B, amp_dim=1400, latent_dim=64, amp_batch=Ba, units=units, disc_units=units, hparams=hp, gemm_backend=be)
    lns[be].load_named(P)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.double().abs().max()), 1e-30))


for s in range(nsteps):
    d, nz = synth.minibatch(st32, cfg, B, Ba, seed=2300 + s, kind=kind)
    d64 = {k: to64(v) for k, v in d.items()}
    res32, g32 = O.calc_gradients(st32, d, cfg, nz, apply_adam=False)
    res64, g64 = O.calc_gradients(st64, d64, cfg, None if nz is None else nz.double(), apply_adam=False)
    outs = {}
    for be, ln in lns.items():
        out = ln.calc_gradients({k: v.cuda() for k, v in d.items() if v is not None}, None if nz is None else nz.cuda(), update_rms=True)
        torch.cuda.synchronize()
        outs[be] = ({k: v.cpu().clone() for k, v in ln.named_grads().items()}, dict(zip(__import__('ase_b200').lib.TR_NAMES, out['scalars'].tolist())))
    print(f"\n== {kind} step {s}: relative error of each gradient tensor (to the tensor's max |g|)")
    print(f"{'tensor':44s} {'ref32-vs-fp64':>13s} | " + ' | '.join(f'be{be}-vs-ref32  be{be}-vs-fp64' for be in lns))
    worst = {be: [0.0, 0.0] for be in lns}; w32 = 0.0
    for k in g32:
        e32 = rel(g32[k], g64[k]); w32 = max(w32, e32)
        row = f"{k:44s} {e32:13.2e} | "
        for be in lns:
            a, b = rel(outs[be][0][k], g32[k]), rel(outs[be][0][k], g64[k])
            worst[be][0] = max(worst[be][0], a); worst[be][1] = max(worst[be][1], b)
            row += f"{a:12.2e} {b:12.2e} | "
        print(row)
    print(f"{'WORST':44s} {w32:13.2e} | " + ' | '.join(f'{worst[be][0]:12.2e} {worst[be][1]:12.2e}' for be in lns))
    print("quantiles of |a - fp64| / max|fp64| per tensor:  ref32 [q50 q90 q99 max]  |  be2 [q50 q90 q99 max]  | be2-vs-ref32 [q50 q90 q99 max] | L2: ref32 be2")
    for k in g32:
        def qs(a, b):
            dd = ((a.double() - b.double()).abs() / max(float(b.double().abs().max()), 1e-30)).flatten()
            if dd.numel() > 2000000:
                dd = dd[torch.randperm(dd.numel())[:2000000]]
            q = torch.quantile(dd, torch.tensor([0.5, 0.9, 0.99], dtype=dd.dtype))
            return f"{float(q[0]):.1e} {float(q[1]):.1e} {float(q[2]):.1e} {float(dd.max()):.1e}"
        l2 = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
        print(f"  {k:42s} {qs(g32[k], g64[k])} | {qs(outs[2][0][k], g64[k])} | {qs(outs[2][0][k], g32[k])} | {l2(g32[k], g64[k]):.1e} {l2(outs[2][0][k], g64[k]):.1e}")
    print("scalars (ref32, fp64, be0, be1, be2):")
    for k in ('actor_loss', 'critic_loss', 'b_loss', 'actor_clip_frac', 'kl', 'disc_loss', 'disc_grad_penalty', 'enc_loss', 'amp_diversity_loss'):
        if k in res32:
            print(f"  {k:20s} {float(res32[k]): .8e} {float(res64[k]): .8e} " + ' '.join(f"{outs[be][1].get(k, float('nan')): .8e}" for be in lns))
    # keep all implementations on the SAME parameters for the next step: apply the fp32 oracle's Adam everywhere
    O.adam_step(st32, g32, cfg)
    for k in st32.p:
        st64.p[k] = st32.p[k].double()
    st64.step = st32.step
    for k in st32.m:
        st64.m[k] = st32.m[k].double(); st64.v[k] = st32.v[k].double()
    for ln in lns.values():
        ln.load_named(st32.p)
