"""GPU parity at the BENCHMARKED size (BASELINE configs 3 and 2: B = 16384, B_amp = 4096, full networks) for the default
FP16-plane tcgen05 backend (2) and the TF32-plane backend (1): dW reductions over K = 16384 .. 32768 with split-K RED accumulation,
interior `epilogue_fast` tiles everywhere, plane scales PREDICTED over consecutive calls.

Checkers: (a) tests/golden/calc_grad_{ase,amp}_full.pt -- two consecutive calls of the reference's own ASEAgent / AMPAgent
.calc_gradients at this size (oracle/gen_golden.py gen_calc_gradients_full); (b) oracle/ase_oracle.py run on the box's CPU in
lock-step (every gradient tensor in full, all scalars, post-Adam parameters, RMS state) for 8 consecutive steps.
Tolerance (north_star): 1e-4 relative, fp32; tensors relative to the tensor's max |value|."""
import os

import pytest
import torch

import ase_oracle as O
import synth
import golden_util as G
from test_gpu_learner import _make_learner, _cuda, _check_step, _check_grads, SCALAR_KEYS

pytestmark = pytest.mark.gpu


def _check_params_conditioned(ln, rec, lr, nsteps_done, when):
    """Sampled post-Adam parameters against the reference's, where the update is well conditioned (|g| well above the parity floor of
    its tensor: Adam's first steps are lr * sign(g)); every sampled element is bounded by the steps taken so far."""
    for k, pv in ln.named_parameters().items():
        p = pv.detach().cpu().flatten()
        idx = G.sample_idx(p.numel())
        gs = rec['grad_sample'][k]
        ok = gs.abs() > 0.05 * max(float(gs.abs().max()), rec['grad_norm'][k] / max(p.numel(), 1) ** 0.5)
        assert torch.allclose(p[idx][ok], rec['param_sample'][k][ok], rtol=1e-5, atol=1e-6), (when, k)
        assert float((p[idx] - rec['param_sample'][k]).abs().max()) <= 2.5 * lr * nsteps_done, (when, k)


def _threads():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 8
    torch.set_num_threads(max(1, min(32, n)))


@pytest.mark.parametrize('backend', [2, 1])
@pytest.mark.parametrize('name', ['calc_grad_ase_full.pt', 'calc_grad_amp_full.pt'])
def test_full_size_calc_gradients_vs_reference_golden(name, backend):
    _threads()
    meta, steps, shapes, P = G.calc_grad_case(name)
    kind = meta['kind']
    assert meta['B'] == 16384 and meta['Ba'] == 4096
    ln = _make_learner(kind, meta, P, backend)
    st = O.LearnerState(P, 253, 1400, kind)
    cfg = meta['cfg']
    for s, rec in enumerate(steps):
        d, new_z = synth.minibatch(st, cfg, meta['B'], meta['Ba'], seed=meta['seed'] * 100 + s, kind=kind)
        out = ln.calc_gradients(_cuda(d), None if new_z is None else new_z.cuda())
        torch.cuda.synchronize()
        _check_step(ln, out, rec)
        _check_grads(ln, rec, f'{name} backend {backend} step {s}')
        ln.adam_step()
        _check_params_conditioned(ln, rec, cfg['lr'], s + 1, f'{name} backend {backend} step {s}')
        r = rec['rms']
        assert torch.allclose(ln.running_mean_std.running_mean.cpu(), r['obs_mean'], rtol=1e-6, atol=1e-7)
        assert torch.allclose(ln.running_mean_std.running_var.cpu(), r['obs_var'], rtol=1e-5, atol=1e-9)
        assert torch.allclose(ln.amp_input_mean_std.running_mean.cpu(), r['amp_mean'], rtol=1e-6, atol=1e-7)
        assert torch.allclose(ln.amp_input_mean_std.running_var.cpu(), r['amp_var'], rtol=1e-5, atol=1e-9)
        if s + 1 < len(steps):
            O.calc_gradients(st, d, cfg, new_z)        # advance the input generator's state in lock-step


def _follow(kind, backend, nsteps, report=None):
    """nsteps consecutive full-size updates, GPU and oracle in lock-step from the same parameters; every step compares all scalars,
    every gradient tensor IN FULL and the post-Adam parameters.  Returns the worst relative errors seen."""
    _threads()
    from ase_b200 import Learner
    B, Ba = 16384, 4096
    if kind == 'ase':
        shapes = O.ase_param_shapes()
        units = disc_units = (1024, 1024, 512)
    else:
        shapes = O.amp_param_shapes()
        units = disc_units = (1024, 512)
    P = synth.params(shapes, seed=23)
    cfg = dict(O.DEFAULT_CFG); cfg['amp_minibatch_size'] = Ba
    if kind == 'amp':
        cfg['enc_coef'] = 0.0; cfg['amp_diversity_bonus'] = 0.0
    st = O.LearnerState(P, 253, 1400, kind)
    hp = {k: cfg[k] for k in ('e_clip', 'critic_coef', 'entropy_coef', 'bounds_loss_coef', 'disc_coef', 'disc_logit_reg', 'disc_grad_penalty',
                              'disc_weight_decay', 'enc_coef', 'amp_diversity_bonus', 'amp_diversity_tar')}
    hp['learning_rate'] = cfg['lr']
    ln = Learner(kind, 253, 31, B, amp_dim=1400, latent_dim=64, amp_batch=Ba, units=units, disc_units=disc_units, hparams=hp, gemm_backend=backend)
    ln.load_named(P)
    worst = {'scalar': 0.0, 'grad': 0.0, 'grad_key': '', 'param': 0.0}
    for s in range(nsteps):
        d, new_z = synth.minibatch(st, cfg, B, Ba, seed=2300 + s, kind=kind)
        out = ln.calc_gradients(_cuda(d), None if new_z is None else new_z.cuda())
        res, grads = O.calc_gradients(st, d, cfg, new_z)          # applies Adam to st.p as well
        tr = ln.train_result(out)                                 # raises on an FP16 plane-scale miss
        for k in SCALAR_KEYS:
            if k in res and k in tr:
                e = abs(tr[k] - float(res[k])) / max(1.0, abs(float(res[k])))
                worst['scalar'] = max(worst['scalar'], e)
                assert e <= 1e-4, (kind, backend, s, k, tr[k], float(res[k]))
        assert torch.allclose(out['disc_agent_logit'].cpu(), res['disc_agent_logit'].flatten(), rtol=1e-4, atol=1e-4)
        assert torch.allclose(out['disc_demo_logit'].cpu(), res['disc_demo_logit'].flatten(), rtol=1e-4, atol=1e-4)
        for k, g in grads.items():
            mine = ln.named_grads()[k].cpu()
            e = float((mine - g).abs().max()) / max(float(g.abs().max()), 1e-12)
            if e > worst['grad']:
                worst['grad'], worst['grad_key'] = e, f'{k} (step {s})'
            assert e <= 1e-4, (kind, backend, s, k, e)
        ln.adam_step()
        for k, g in grads.items():
            mine = ln.named_parameters()[k].cpu()
            # Adam's first steps are ~ lr * sign(g): where |g| is far below the tensor's max the update is ill conditioned in ANY fp32
            # implementation (the HRL golden test has the same rule), so elements are compared where the gradient is well above the
            # 1e-4 parity floor, and every element is bounded by the size of the steps taken so far
            ok = g.abs() > 0.05 * g.abs().max()
            e = float((mine - st.p[k])[ok].abs().max()) if bool(ok.any()) else 0.0
            worst['param'] = max(worst['param'], e / cfg['lr'])
            assert e <= 0.02 * cfg['lr'] + 2e-7 * float(st.p[k].abs().max()), (kind, backend, s, k, e)
            assert float((mine - st.p[k]).abs().max()) <= 2.5 * cfg['lr'] * (s + 1), (kind, backend, s, k)
        assert torch.allclose(ln.running_mean_std.running_mean.cpu(), st.obs_rms.mean, rtol=1e-6, atol=1e-7)
        assert torch.allclose(ln.amp_input_mean_std.running_var.cpu(), st.amp_rms.var, rtol=1e-5, atol=1e-9)
    if report is not None:
        report.append((kind, backend, nsteps, worst))
    return worst


def test_full_size_ase_fp16_planes_8_steps_vs_oracle():
    """Config 3 on the default backend: step 0 calibrates the plane scales exactly, steps 1..7 run on scales predicted from the
    previous call, RMS updates on, parameters moving under Adam."""
    w = _follow('ase', 2, 8)
    print('full-size ASE backend 2, 8 steps: worst', w)


def test_full_size_ase_tf32_planes_vs_oracle():
    w = _follow('ase', 1, 2)
    print('full-size ASE backend 1, 2 steps: worst', w)


@pytest.mark.parametrize('backend', [2, 1])
def test_full_size_amp_vs_oracle(backend):
    """Config 2 (AMP only: no encoder / latents / diversity, MLPs [1024, 512])."""
    w = _follow('amp', backend, 3)
    print(f'full-size AMP backend {backend}, 3 steps: worst', w)
