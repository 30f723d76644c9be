"""CPU suite, part 1: the travelling oracle (oracle/ase_oracle.py) against the committed outputs of the
reference's own code (tests/golden/*.pt, produced by oracle/gen_golden.py in the build container)."""
import torch
import pytest

import ase_oracle as O
import synth
import golden_util as G

TOL = 2e-6   # oracle and reference are both fp32 torch CPU; expect (near) bit equality


def test_obs_build_matches_reference_golden():
    fx = G.load('obs_build.pt')
    s = fx['inputs']
    for lro in (True, False):
        for rho in (True, False):
            o = O.compute_humanoid_observations_max(s['body_pos'], s['body_rot'], s['body_vel'], s['body_ang_vel'], lro, rho)
            assert torch.allclose(o, fx[f'obs_l{int(lro)}_h{int(rho)}'], rtol=1e-5, atol=1e-5)
            kp = s['body_pos'][:, O.KEY_BODY_IDS_SWORD_SHIELD]
            a = O.build_amp_observations(s['body_pos'][:, 0], s['body_rot'][:, 0], s['body_vel'][:, 0], s['body_ang_vel'][:, 0],
                                         s['dof_pos'], s['dof_vel'], kp, lro, rho, O.DOF_OFFSETS_SWORD_SHIELD)
            assert torch.allclose(a, fx[f'amp_l{int(lro)}_h{int(rho)}'], rtol=1e-5, atol=1e-5)


def test_rollout_math_matches_reference_golden():
    fx = G.load('rollout_math.pt')
    adv = O.discount_values(fx['fdones'], fx['values'], fx['rewards'], fx['next_values'], 0.99, 0.95)
    assert torch.allclose(adv, fx['advs'], rtol=1e-6, atol=1e-6)
    ret = O.swap_and_flatten01(fx['advs'] + fx['values']); vals = O.swap_and_flatten01(fx['values'])
    assert torch.allclose(O.calc_advs(ret, vals, fx['mask']), fx['advs_norm'], rtol=1e-5, atol=1e-6)
    assert torch.allclose(O.disc_rewards(fx['logits'], 2.0), fx['disc_r'], rtol=1e-6, atol=1e-7)
    assert torch.allclose(O.enc_rewards(fx['enc_pred'], fx['z'], 1.0), fx['enc_r'], rtol=1e-6, atol=1e-7)
    assert torch.allclose(0.5 * fx['disc_r'] + 0.5 * fx['enc_r'], fx['combined'], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('name', ['calc_grad_ase_small.pt', 'calc_grad_ase_cfg1.pt', 'calc_grad_amp_cfg.pt',
                                  'calc_grad_ase_full.pt', 'calc_grad_amp_full.pt'])     # *_full: B = 16384, B_amp = 4096 (the benchmarked size)
def test_calc_gradients_matches_reference_golden(name):
    meta, steps, shapes, P = G.calc_grad_case(name)
    st = O.LearnerState(P, 253, 1400, meta['kind'])
    cfg = meta['cfg']
    for s, rec in enumerate(steps):
        d, new_z = synth.minibatch(st, cfg, meta['B'], meta['Ba'], seed=meta['seed'] * 100 + s, kind=meta['kind'])
        res, grads = O.calc_gradients(st, d, cfg, new_z)
        for k, v in rec['scalars'].items():
            assert abs(float(res[k]) - v) <= 1e-5 * max(1.0, abs(v)), (k, float(res[k]), v)
        assert torch.allclose(res['disc_agent_logit'].flatten(), rec['disc_agent_logit'], rtol=1e-5, atol=1e-5)
        for k, g in grads.items():
            idx = G.sample_idx(g.numel())
            assert G.rel_err(g.flatten()[idx], rec['grad_sample'][k]) < 1e-4 or float(rec['grad_sample'][k].abs().max()) < 1e-12, k
            assert abs(float(g.double().norm()) - rec['grad_norm'][k]) <= 1e-4 * max(rec['grad_norm'][k], 1e-12), k
            assert torch.allclose(st.p[k].flatten()[idx], rec['param_sample'][k], rtol=1e-6, atol=1e-7), k
            if 'grads' in rec:
                assert G.rel_err(g, rec['grads'][k]) < 1e-4, k
        assert torch.allclose(st.obs_rms.mean, rec['rms']['obs_mean']) and torch.allclose(st.obs_rms.var, rec['rms']['obs_var'])
        assert torch.allclose(st.amp_rms.mean, rec['rms']['amp_mean']) and torch.allclose(st.amp_rms.var, rec['rms']['amp_var'])
        assert float(st.amp_rms.count) == float(rec['rms']['amp_count'])


def test_rms_count_arithmetic_matches_checkpoint_identity():
    """AMP RMS is updated 3x per minibatch with amp_minibatch rows, count starts at 1."""
    r = O.RMS(4)
    for _ in range(3):
        r.train_forward(torch.randn(16, 4))
    assert float(r.count) == 1 + 3 * 16


def test_heading_task_matches_reference_golden():
    fx = G.load('heading.pt')
    obs = O.compute_heading_observations(fx['root'], fx['tar_dir'], fx['tar_speed'], fx['tar_face_dir'])
    assert torch.allclose(obs, fx['obs'], rtol=1e-5, atol=1e-6)
    rew = O.compute_heading_reward(fx['root'][:, 0:3], fx['prev'], fx['root'][:, 3:7], fx['tar_dir'], fx['tar_speed'], fx['tar_face_dir'], 1.0 / 30.0)
    assert torch.allclose(rew, fx['reward'], rtol=1e-5, atol=1e-6)


def test_hrl_high_level_learner_matches_reference_golden():
    """CommonAgent.calc_gradients over HRLBuilder's tanh-mu network (BASELINE config 5 learner), small units."""
    fx = G.load('calc_grad_hrl_small.pt')
    meta = fx['meta']
    P = synth.params(O.amp_param_shapes(obs=258, act=64, amp=0, units=meta['units']), seed=meta['seed'])
    st = O.LearnerState(P, 258, 0, 'ppo')
    for s, rec in enumerate(fx['steps']):
        d, _ = synth.minibatch(st, meta['cfg'], meta['B'], 0, seed=meta['seed'] * 100 + s, kind='ppo', obs_dim=258, act=64)
        res, grads = O.calc_gradients(st, d, meta['cfg'], None)
        for k, v in rec['scalars'].items():
            assert abs(float(res[k]) - v) <= 1e-5 * max(1.0, abs(v)), k
        for k, g in grads.items():
            assert G.rel_err(g, rec['grads'][k]) < 1e-4, k
            assert torch.allclose(st.p[k], rec['params_after'][k], rtol=1e-6, atol=1e-7), k


def test_motion_lib_matches_reference_golden():
    """MotionLib.get_motion_state + build_amp_obs_demo (reference, on synthetic clip tables) vs the oracle restatement."""
    fx = G.load('motion_lib.pt')
    mt = O.synthetic_motion_tables(seed=fx['seed'])
    state = O.get_motion_state(mt, fx['ids'], fx['t0'])
    for mine, ref, name in zip(state, fx['state'], ('root_pos', 'root_rot', 'dof_pos', 'root_vel', 'root_ang_vel', 'dof_vel', 'key_pos')):
        assert torch.allclose(mine, ref, rtol=1e-5, atol=1e-5), name
    demo = O.build_amp_obs_demo(mt, fx['ids'], fx['t0'], fx['sim_dt'], fx['steps'])
    assert torch.allclose(demo, fx['demo'], rtol=1e-5, atol=1e-5)


def _shipped_stats_state(fx):
    P = synth.params(O.ase_param_shapes(), seed=fx['param_seed'])
    st = O.LearnerState(P, 253, 1400, 'ase')
    for r, k in ((st.obs_rms, 'running_mean_std'), (st.val_rms, 'reward_mean_std'), (st.amp_rms, 'amp_input_mean_std')):
        r.mean, r.var, r.count = fx['rms'][k]['running_mean'].clone(), fx['rms'][k]['running_var'].clone(), fx['rms'][k]['count'].clone()
    return st


def test_inference_with_shipped_checkpoint_statistics_matches_reference_golden():
    """The reference's inference path under the shipped checkpoint's RunningMeanStd statistics (variances down to 1.3e-11, counts of 1e11):
    normalised inputs incl. the +-5 clamp rows, actor mu, critic value (normalised and un-normalised), disc logit, encoder output, rewards."""
    fx = G.load('inference_shipped_stats.pt')
    st = _shipped_stats_state(fx)
    with torch.no_grad():
        on, an = st.obs_rms.norm(fx['obs']), st.amp_rms.norm(fx['amp'])
        assert torch.allclose(on, fx['obs_norm'], rtol=1e-6, atol=1e-6) and torch.allclose(an, fx['amp_norm'], rtol=1e-6, atol=1e-6)
        assert float(on.abs().max()) == 5.0 and float(an.abs().max()) == 5.0
        assert torch.allclose(O.eval_actor(st.p, on, fx['z']), fx['mu'], rtol=1e-5, atol=1e-5)
        assert torch.allclose(O.eval_critic(st.p, on, fx['z']), fx['value_normed'], rtol=1e-5, atol=1e-5)
        assert torch.allclose(O.eval_critic_unnorm(st, fx['obs'], fx['z']), fx['value'], rtol=1e-5, atol=1e-5)
        assert torch.allclose(O.eval_disc(st.p, an), fx['disc_logit'], rtol=1e-5, atol=1e-5)
        assert torch.allclose(O.eval_enc(st.p, an), fx['enc'], rtol=1e-5, atol=1e-6)
        dr, er = O.calc_amp_rewards(st, fx['amp'], fx['z'], O.DEFAULT_CFG)
        assert torch.allclose(dr, fx['disc_r'], rtol=1e-5, atol=1e-5) and torch.allclose(er, fx['enc_r'], rtol=1e-5, atol=1e-6)
