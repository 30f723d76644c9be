"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt by executing the UNMODIFIED reference
(/root/reference, through oracle/shims) on seeded synthetic inputs.  Runs only in the build container;
the fixtures travel to the GPU box.  Usage: python oracle/gen_golden.py
Every fixture stores the inputs (or the seeds that regenerate them through oracle/synth.py) and the
reference's outputs."""
import os
import sys
import copy

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh       # noqa: E402
import ase_oracle as O         # noqa: E402
import synth                   # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def _sample_idx(numel, k=256):
    g = torch.Generator().manual_seed(numel)
    return torch.randint(0, numel, (min(k, numel),), generator=g)


def gen_obs():
    humanoid, humanoid_amp, _ = rh.import_env_fns()
    s = synth.rigid_body_state(64, seed=11, edge_cases=True)
    out = {'inputs': s}
    for lro in (True, False):
        for rho in (True, False):
            out[f'obs_l{int(lro)}_h{int(rho)}'] = humanoid.compute_humanoid_observations_max(
                s['body_pos'], s['body_rot'], s['body_vel'], s['body_ang_vel'], lro, rho)
            kp = s['body_pos'][:, O.KEY_BODY_IDS_SWORD_SHIELD]
            out[f'amp_l{int(lro)}_h{int(rho)}'] = humanoid_amp.build_amp_observations(
                s['body_pos'][:, 0], s['body_rot'][:, 0], s['body_vel'][:, 0], s['body_ang_vel'][:, 0],
                s['dof_pos'], s['dof_vel'], kp, lro, rho, 78, O.DOF_OFFSETS_SWORD_SHIELD)
    torch.save(out, os.path.join(OUT, 'obs_build.pt'))
    print('obs_build.pt', {k: tuple(v.shape) for k, v in out.items() if torch.is_tensor(v)})


def _load_params(agent, P, ase):
    sd = {'a2c_network.' + k: v.clone() for k, v in P.items()}
    if ase:
        for k in list(sd):
            if '_disc_mlp' in k:
                sd[k.replace('_disc_mlp', '_enc_mlp')] = sd[k]
    agent.model.load_state_dict(sd, strict=True)


def _run_steps(kind, shapes_kw, net_over, B, Ba, nsteps, seed, full):
    """kind: 'ase' | 'amp'.  Runs nsteps consecutive reference calc_gradients calls."""
    import ref_harness
    agent, params = None, None
    # network dims are YAML-driven: patch the YAML dict before the builder reads it
    orig = ref_harness.load_train_cfg

    def patched(name):
        c = orig(name)
        for sect, units in net_over.items():
            c['params']['network'][sect]['units'] = list(units)
        return c
    ref_harness.load_train_cfg = patched
    try:
        agent, params = rh.make_ref_agent(kind, num_envs=B // 32, overrides={'minibatch_size': B, 'amp_minibatch_size': Ba})
    finally:
        ref_harness.load_train_cfg = orig
    shapes = (O.ase_param_shapes if kind == 'ase' else O.amp_param_shapes)(**shapes_kw)
    P = synth.params(shapes, seed=seed)
    _load_params(agent, P, kind == 'ase')
    cfg = dict(O.DEFAULT_CFG); cfg['amp_minibatch_size'] = Ba
    if kind == 'amp':
        cfg['enc_coef'] = 0.0; cfg['amp_diversity_bonus'] = 0.0
    st = O.LearnerState(P, 253, 1400, kind)
    steps = []
    for s in range(nsteps):
        d, new_z = synth.minibatch(st, cfg, B, Ba, seed=seed * 100 + s, kind=kind)
        if kind == 'ase':
            agent._sample_latents = (lambda nz: (lambda n: nz))(new_z)
        agent.calc_gradients(d)
        tr = agent.train_result
        rec = {'scalars': {k: float(v) for k, v in tr.items() if torch.is_tensor(v) and v.numel() == 1},
               'disc_agent_logit': tr['disc_agent_logit'].flatten().clone(),
               'disc_demo_logit': tr['disc_demo_logit'].flatten().clone(),
               'actor_clipped': tr['actor_clipped'].float().clone()}
        g_full, p_full, g_norm, g_samp, p_samp = {}, {}, {}, {}, {}
        for n, prm in agent.model.named_parameters():
            k = n[len('a2c_network.'):]
            if k == 'sigma' or '_enc_mlp' in k:
                continue
            gr = prm.grad.detach()
            idx = _sample_idx(gr.numel())
            g_norm[k] = float(gr.double().norm())
            g_samp[k] = gr.flatten()[idx].clone()
            p_samp[k] = prm.detach().flatten()[idx].clone()
            if full and s == nsteps - 1:
                g_full[k] = gr.clone()
        rec.update(grad_norm=g_norm, grad_sample=g_samp, param_sample=p_samp)
        if full and s == nsteps - 1:
            rec.update(grads=g_full)
        rec['rms'] = {'obs_mean': agent.running_mean_std.running_mean.clone(), 'obs_var': agent.running_mean_std.running_var.clone(),
                      'obs_count': agent.running_mean_std.count.clone(),
                      'amp_mean': agent._amp_input_mean_std.running_mean.clone(), 'amp_var': agent._amp_input_mean_std.running_var.clone(),
                      'amp_count': agent._amp_input_mean_std.count.clone()}
        steps.append(rec)
        # advance the oracle state in lock-step so synth.minibatch (which uses st for old_logp) tracks the reference
        O.calc_gradients(st, d, cfg, new_z)
    meta = dict(kind=kind, shapes_kw=shapes_kw, B=B, Ba=Ba, seed=seed, nsteps=nsteps, cfg=cfg,
                param_checksum={k: float(v.double().sum()) for k, v in P.items()})
    return {'meta': meta, 'steps': steps}


def gen_calc_gradients():
    small = dict(units=(64, 48, 32), disc_units=(48, 40, 24))
    r = _run_steps('ase', dict(units=small['units'], disc_units=small['disc_units']),
                   {'mlp': small['units'], 'disc': small['disc_units']}, B=64, Ba=16, nsteps=2, seed=5, full=True)
    torch.save(r, os.path.join(OUT, 'calc_grad_ase_small.pt'))
    print('calc_grad_ase_small', r['steps'][-1]['scalars'])
    r = _run_steps('ase', {}, {}, B=256, Ba=64, nsteps=2, seed=7, full=False)
    torch.save(r, os.path.join(OUT, 'calc_grad_ase_cfg1.pt'))
    print('calc_grad_ase_cfg1', r['steps'][-1]['scalars'])
    r = _run_steps('amp', {}, {}, B=256, Ba=64, nsteps=2, seed=9, full=False)
    torch.save(r, os.path.join(OUT, 'calc_grad_amp_cfg.pt'))
    print('calc_grad_amp_cfg', r['steps'][-1]['scalars'])


def gen_calc_gradients_full():
    """BASELINE configs 3 and 2 at the benchmarked size (B = 16384, B_amp = 4096, full networks): two consecutive calls of the
    reference's own ASEAgent / AMPAgent.calc_gradients (SURVEY.md 8c "... and at full size").  Sampled gradients / parameters,
    full gradient norms, all scalars and the RMS state are stored (a few hundred kB)."""
    r = _run_steps('ase', {}, {}, B=16384, Ba=4096, nsteps=2, seed=17, full=False)
    torch.save(r, os.path.join(OUT, 'calc_grad_ase_full.pt'))
    print('calc_grad_ase_full', r['steps'][-1]['scalars'])
    r = _run_steps('amp', {}, {}, B=16384, Ba=4096, nsteps=2, seed=19, full=False)
    torch.save(r, os.path.join(OUT, 'calc_grad_amp_full.pt'))
    print('calc_grad_amp_full', r['steps'][-1]['scalars'])


def gen_rollout():
    """The reference's own ASEAgent.play_steps + prepare_dataset (learning/ase_agent.py:36-156, common_agent.py:309-351) over a scripted
    vec-env (ref_harness.FakeVecEnv) with every random draw RECORDED in call order: action noise (Normal.sample), the eps-greedy bernoulli
    mask, the latents of _reset_latents and the randint_like step counts.  Small networks; 16 envs x 8 steps; dones / terminations /
    latent horizons chosen so that resets, latent refreshes and the first-step `done_indices = []` rule all occur."""
    import ref_harness
    N, H, Z, A = 16, 8, 64, 31
    small = dict(units=(64, 48, 32), disc_units=(48, 40, 24))
    orig = ref_harness.load_train_cfg

    def patched(name):
        c = orig(name)
        c['params']['network']['mlp']['units'] = list(small['units']); c['params']['network']['disc']['units'] = list(small['disc_units'])
        return c
    ref_harness.load_train_cfg = patched
    try:
        agent, params = rh.make_ref_agent('ase', num_envs=N, overrides={'minibatch_size': 64, 'amp_minibatch_size': 32, 'horizon_length': H,
                                                                         'latent_steps_min': 1, 'latent_steps_max': 6})
    finally:
        ref_harness.load_train_cfg = orig
    P = synth.params(O.ase_param_shapes(units=small['units'], disc_units=small['disc_units']), seed=31)
    _load_params(agent, P, True)
    g = torch.Generator().manual_seed(77)
    # non-trivial normaliser statistics
    for rms, dim in ((agent.running_mean_std, 253), (agent.value_mean_std, 1), (agent._amp_input_mean_std, 1400)):
        rms.running_mean.copy_(torch.randn(dim, generator=g).double() * 0.3); rms.running_var.copy_((0.5 + torch.rand(dim, generator=g)).double())
        rms.count.fill_(1000.0)
    rms_state = {k: {'running_mean': r.running_mean.clone(), 'running_var': r.running_var.clone(), 'count': r.count.clone()}
                 for k, r in (('running_mean_std', agent.running_mean_std), ('reward_mean_std', agent.value_mean_std), ('amp_input_mean_std', agent._amp_input_mean_std))}
    # scripted env
    obs_t = torch.randn(H + 1, N, 253, generator=g) * 1.5 + 0.3
    reset_obs_t = torch.randn(H + 1, N, 253, generator=g) * 1.5 - 0.2
    amp_t = torch.randn(H, N, 1400, generator=g)
    rew_t = torch.rand(H, N, generator=g)
    dones_t = (torch.rand(H, N, generator=g) < 0.25).to(torch.uint8)
    dones_t[H - 1, 0] = 1                                   # a done on the LAST step: must not be reset at the start of the next rollout
    term_t = (dones_t.bool() & (torch.rand(H, N, generator=g) < 0.5)).to(torch.uint8)
    env = agent.vec_env
    task = env.env.task
    task.progress_buf = torch.randint(0, 4, (N,), generator=g)
    progress0 = task.progress_buf.clone()
    cur = {'obs': obs_t[0].clone()}

    def script(t, actions):
        cur['obs'] = obs_t[t + 1].clone()
        return cur['obs'], rew_t[t].clone(), dones_t[t].clone(), {'amp_obs': amp_t[t].clone(), 'terminate': term_t[t].clone()}

    def reset_fn(env_ids):
        if env_ids is None:
            env_ids = torch.arange(N)
        if len(env_ids) > 0:
            task.progress_buf[env_ids] = 0
            cur['obs'][env_ids] = reset_obs_t[env.t][env_ids]
        return cur['obs']
    env.script, env.reset_fn = script, reset_fn
    # recorded RNG
    rec = {'normal': [], 'bernoulli': [], 'latents': [], 'randint': []}
    real_sample, real_bern, real_randint_like = torch.distributions.Normal.sample, torch.bernoulli, torch.randint_like

    def rec_sample(self, sample_shape=torch.Size()):
        eps = torch.randn(self.loc.shape, generator=g)
        rec['normal'].append(eps.clone())
        return self.loc + self.scale * eps

    def rec_bern(p, *a, **k):
        m = (torch.rand(p.shape, generator=g) < p).float()
        rec['bernoulli'].append(m.clone())
        return m

    def rec_randint_like(x, low=0, high=None, **k):
        r = torch.randint(low, high, x.shape, generator=g, dtype=x.dtype)
        rec['randint'].append(r.clone())
        return r
    real_lat = agent.model.a2c_network.sample_latents

    def rec_latents(n):
        z = torch.nn.functional.normalize(torch.randn(n, Z, generator=g), dim=-1)
        rec['latents'].append(z.clone())
        return z
    torch.distributions.Normal.sample, torch.bernoulli, torch.randint_like = rec_sample, rec_bern, rec_randint_like
    agent.model.a2c_network.sample_latents = rec_latents
    try:
        agent.init_tensors()
        agent.obs = agent.env_reset()                    # train(): resets every env and every latent
        latents0, steps0 = agent._ase_latents.clone(), agent._latent_reset_steps.clone()
        n0 = {k: len(v) for k, v in rec.items()}
        agent.set_eval()
        with torch.no_grad():
            batch_dict = agent.play_steps()
        agent.set_train()
        played = batch_dict.pop('played_frames')
        # (train_epoch adds the demo / replay samples before prepare_dataset, amp_agent.py:194-202: not part of this fixture)
        batch_dict['amp_obs_demo'] = torch.zeros_like(batch_dict['amp_obs']); batch_dict['amp_obs_replay'] = torch.zeros_like(batch_dict['amp_obs'])
        agent.prepare_dataset(batch_dict)
    finally:
        torch.distributions.Normal.sample, torch.bernoulli, torch.randint_like = real_sample, real_bern, real_randint_like
        agent.model.a2c_network.sample_latents = real_lat
    eb = {k: v.clone() for k, v in agent.experience_buffer.tensor_dict.items()}
    assert torch.equal(eb['amp_obs'], amp_t) and torch.equal(eb['next_obses'], obs_t[1:])
    del eb['amp_obs'], eb['next_obses']             # (== the scripted inputs: the test checks them against those)
    ds = {k: v.clone() for k, v in agent.dataset.values_dict.items() if torch.is_tensor(v) and not k.startswith('amp_obs') and k != 'obs'}
    batch_dict = {k: v for k, v in batch_dict.items() if k in ('returns', 'disc_rewards', 'enc_rewards')}
    out = dict(N=N, H=H, units=small['units'], disc_units=small['disc_units'], param_seed=31, rms_state=rms_state,
               obs_t=obs_t, reset_obs_t=reset_obs_t, amp_t=amp_t, rew_t=rew_t, dones_t=dones_t, term_t=term_t, progress0=progress0,
               latents0=latents0, steps0=steps0, rec={k: v[n0[k]:] for k, v in rec.items()}, rec_init={k: v[:n0[k]] for k, v in rec.items()},
               eb=eb, batch={k: v.clone() for k, v in batch_dict.items() if torch.is_tensor(v)}, dataset=ds,
               latents_end=agent._ase_latents.clone(), steps_end=agent._latent_reset_steps.clone(), progress_end=task.progress_buf.clone(),
               value_rms_after={'running_mean': agent.value_mean_std.running_mean.clone(), 'running_var': agent.value_mean_std.running_var.clone(),
                                'count': agent.value_mean_std.count.clone()},
               cfg={k: agent.config[k] for k in ('gamma', 'tau', 'disc_reward_scale', 'enc_reward_scale', 'task_reward_w', 'disc_reward_w', 'enc_reward_w')})
    torch.save(out, os.path.join(OUT, 'rollout_ase.pt'))
    print('rollout_ase.pt', {k: tuple(v.shape) for k, v in eb.items()}, {k: len(v) for k, v in out['rec'].items()}, 'dones', int(dones_t.sum()))


def gen_humanoid_reset():
    """compute_humanoid_reset (env/tasks/humanoid.py:645-670) on seeded inputs incl. the progress_buf <= 1 guard and both early-termination settings."""
    humanoid, _, _ = rh.import_env_fns()
    g = torch.Generator().manual_seed(41)
    N, J = 96, 17
    contact = torch.randn(N, J, 3, generator=g) * 0.08
    contact[torch.rand(N, J, generator=g) < 0.15] *= 40.0
    pos = torch.randn(N, J, 3, generator=g); pos[..., 2] = torch.rand(N, J, generator=g) * 0.6
    heights = torch.rand(J, generator=g) * 0.3
    progress = torch.randint(0, 310, (N,), generator=g); progress[:6] = torch.tensor([0, 1, 2, 298, 299, 300])
    ids = torch.tensor([13, 16, 7, 4])          # contact_body_ids (feet / hands)
    out = dict(contact=contact, pos=pos, heights=heights, progress=progress, contact_body_ids=ids, max_episode_length=300.0)
    for et in (True, False):
        r, t = humanoid.compute_humanoid_reset(torch.zeros(N, dtype=torch.long), progress, contact, ids, pos, 300.0, et, heights)
        out[f'reset_{int(et)}'], out[f'term_{int(et)}'] = r, t
    torch.save(out, os.path.join(OUT, 'humanoid_reset.pt'))
    print('humanoid_reset.pt', int(out['reset_1'].sum()), int(out['term_1'].sum()))


def gen_motion_lib():
    """The reference's own MotionLib.get_motion_state + build_amp_observations on synthetic clip tables (the object is
    assembled field by field so no .npy clip has to travel; the real loader only fills these same tensors)."""
    humanoid, humanoid_amp, _ = rh.import_env_fns()
    from utils.motion_lib import MotionLib
    mt = O.synthetic_motion_tables(seed=7)
    ml = MotionLib.__new__(MotionLib)
    ml._dof_body_ids = O.DOF_BODY_IDS_SWORD_SHIELD; ml._dof_offsets = O.DOF_OFFSETS_SWORD_SHIELD; ml._num_dof = 31
    ml._key_body_ids = torch.tensor(O.KEY_BODY_IDS_SWORD_SHIELD); ml._device = 'cpu'
    ml.gts, ml.grs, ml.lrs, ml.grvs, ml.gravs, ml.dvs = mt.gts, mt.grs, mt.lrs, mt.grvs, mt.gravs, mt.dvs
    ml._motion_lengths, ml._motion_num_frames, ml._motion_dt, ml.length_starts = mt.lengths, mt.num_frames, mt.dts, mt.length_starts

    class _M: num_joints = 17
    ml._motions = [_M()]
    g = torch.Generator().manual_seed(8)
    n = 96
    ids = torch.randint(0, 4, (n,), generator=g)
    sim_dt, steps = 1.0 / 30.0, 10
    trunc = sim_dt * (steps - 1)
    t0 = torch.rand(n, generator=g) * torch.clamp(mt.lengths[ids] - trunc, min=0.0) + trunc        # humanoid_amp.py:64-83
    t0[0] = trunc; t0[1] = mt.lengths[ids[1]] + 0.5          # clip start / beyond the end (phase clipped to 1)
    state = ml.get_motion_state(ids, t0)
    tid = ids.unsqueeze(-1).expand(n, steps).reshape(-1)
    tt = (t0.unsqueeze(-1) - sim_dt * torch.arange(0, steps)).reshape(-1)
    rp, rr, dp, rv, rw, dv, kp = ml.get_motion_state(tid, tt)
    demo = humanoid_amp.build_amp_observations(rp, rr, rv, rw, dp, dv, kp, True, True, 78, O.DOF_OFFSETS_SWORD_SHIELD).reshape(n, -1)
    torch.save(dict(seed=7, ids=ids, t0=t0, sim_dt=sim_dt, steps=steps, state=[s.clone() for s in state], demo=demo),
               os.path.join(OUT, 'motion_lib.pt'))
    print('motion_lib.pt', tuple(demo.shape))


def gen_heading():
    rh.import_env_fns()
    from env.tasks import humanoid_heading
    g = torch.Generator().manual_seed(31)
    n = 64
    root = torch.randn(n, 13, generator=g); root[:, 3:7] = torch.nn.functional.normalize(torch.randn(n, 4, generator=g), dim=-1)
    prev = root[:, 0:3] - 0.03 * torch.randn(n, 3, generator=g)
    td = torch.nn.functional.normalize(torch.randn(n, 2, generator=g), dim=-1); fd = torch.nn.functional.normalize(torch.randn(n, 2, generator=g), dim=-1)
    sp = 1.0 + 4.0 * torch.rand(n, generator=g)
    obs = humanoid_heading.compute_heading_observations(root, td, sp, fd)
    rew = humanoid_heading.compute_heading_reward(root[:, 0:3], prev, root[:, 3:7], td, sp, fd, 1.0 / 30.0)
    torch.save(dict(root=root, prev=prev, tar_dir=td, tar_speed=sp, tar_face_dir=fd, obs=obs, reward=rew), os.path.join(OUT, 'heading.pt'))
    print('heading.pt ok')


def gen_hrl_calc_gradients():
    """HLC learner (config 5): CommonAgent.calc_gradients over HRLBuilder's tanh-mu network, obs 258, act 64, units [1024,512]."""
    B = 128
    import ref_harness
    orig = ref_harness.load_train_cfg

    def patched(name):
        c = orig(name); c['params']['network']['mlp']['units'] = [96, 64]; return c
    ref_harness.load_train_cfg = patched
    try:
        agent, params = rh.make_ref_agent('hrl', num_envs=B // 32, overrides={'minibatch_size': B}, obs_dim=258, act_dim=64)
    finally:
        ref_harness.load_train_cfg = orig
    shapes = O.amp_param_shapes(obs=258, act=64, amp=0, units=(96, 64))
    P = synth.params(shapes, seed=13)
    agent.model.load_state_dict({'a2c_network.' + k: v.clone() for k, v in P.items()}, strict=True)
    cfg = dict(O.DEFAULT_CFG); cfg['mu_tanh'] = True
    st = O.LearnerState(P, 258, 0, 'ppo')
    steps = []
    for s in range(2):
        d, _ = synth.minibatch(st, cfg, B, 0, seed=1300 + s, kind='ppo', obs_dim=258, act=64)
        agent.calc_gradients(d)
        tr = agent.train_result
        rec = {'scalars': {k: float(v) for k, v in tr.items() if torch.is_tensor(v) and v.numel() == 1}, 'grads': {}, 'params_after': {}}
        for n, prm in agent.model.named_parameters():
            k = n[len('a2c_network.'):]
            if k == 'sigma':
                continue
            rec['grads'][k] = prm.grad.detach().clone(); rec['params_after'][k] = prm.detach().clone()
        steps.append(rec)
        O.calc_gradients(st, d, cfg, None)
    torch.save({'meta': dict(B=B, seed=13, units=(96, 64), cfg=cfg), 'steps': steps}, os.path.join(OUT, 'calc_grad_hrl_small.pt'))
    print('calc_grad_hrl_small', steps[-1]['scalars'])


def gen_checkpoint_layout():
    """Layout (keys, shapes, dtypes, optimizer state indices, hyper-parameter keys -- NO weights) of the checkpoints the reference ships
    (ase/data/models/*.pth, written by common_agent.py:157-170 save / a2c_common get_full_state_weights): what restore() and the HRL
    llc_checkpoint path have to accept.  tests build a random checkpoint with exactly this layout."""
    import json
    ref = os.environ.get('ASE_REFERENCE', '/root/reference')
    out = {}
    for name in ('ase_llc_reallusion_sword_shield', 'ase_hlc_heading_reallusion_sword_shield'):
        w = torch.load(os.path.join(ref, 'ase/data/models', name + '.pth'), map_location='cpu', weights_only=False)
        lay = {}
        for k, v in w.items():
            if k == 'optimizer':
                lay[k] = {'state': {str(i): {kk: (list(vv.shape), str(vv.dtype)) if torch.is_tensor(vv) else type(vv).__name__ for kk, vv in st.items()}
                                    for i, st in v['state'].items()},
                          'param_groups': [{kk: (type(vv).__name__ if kk != 'params' else list(vv)) for kk, vv in pg.items()} for pg in v['param_groups']]}
            elif isinstance(v, dict):
                lay[k] = {kk: (list(vv.shape), str(vv.dtype)) for kk, vv in v.items()}
            else:
                lay[k] = type(v).__name__
        out[name] = lay
    with open(os.path.join(OUT, 'checkpoint_layout.json'), 'w') as f:
        json.dump(out, f, indent=0, sort_keys=False)
    print('checkpoint_layout.json ok', {k: len(v['model']) for k, v in out.items()})


def gen_inference_shipped_stats():
    """SURVEY 8(c) "forward of the shipped LLC checkpoint on fixed inputs", without shipping 28 MB of weights: the reference's own inference
    path (ASEAgent._preproc_obs / _preproc_amp_obs in eval mode, a2c_network.eval_actor / eval_critic / eval_disc / eval_enc, value un-normalisation,
    _calc_disc_rewards / _calc_enc_rewards) with the SHIPPED checkpoint's RunningMeanStd statistics (f64; variances from 1.3e-11 to ~1e2 after
    8.4e9 frames -- the realistic conditioning of the normaliser) and seeded synthetic full-size network weights.  Inputs are drawn around the
    shipped means with 1.5 x the shipped standard deviations, plus rows far outside (the +-5 clamp) and exactly at the mean."""
    ref = os.environ.get('ASE_REFERENCE', '/root/reference')
    ck = torch.load(os.path.join(ref, 'ase/data/models/ase_llc_reallusion_sword_shield.pth'), map_location='cpu', weights_only=False)
    agent, _ = rh.make_ref_agent('ase', num_envs=8, overrides={'minibatch_size': 256, 'amp_minibatch_size': 64})
    P = synth.params(O.ase_param_shapes(), seed=31)
    _load_params(agent, P, True)
    agent.running_mean_std.load_state_dict(ck['running_mean_std'])
    agent.value_mean_std.load_state_dict(ck['reward_mean_std'])
    agent._amp_input_mean_std.load_state_dict(ck['amp_input_mean_std'])
    agent.set_eval()
    g = torch.Generator().manual_seed(77)
    n = 96
    def around(st):
        m, sd = st['running_mean'].float(), st['running_var'].float().sqrt()
        x = m + 1.5 * sd * torch.randn(n, m.numel(), generator=g)
        x[0] = m; x[1] = m + 40.0 * sd; x[2] = m - 40.0 * sd; x[3] = m + 1e-3 * torch.randn(m.numel(), generator=g)
        return x
    obs, amp = around(ck['running_mean_std']), around(ck['amp_input_mean_std'])
    z = torch.nn.functional.normalize(torch.randn(n, 64, generator=g), dim=-1)
    net = agent.model.a2c_network
    with torch.no_grad():
        po = agent._preproc_obs(obs)
        mu, _ = net.eval_actor(obs=po, ase_latents=z)
        v_n = net.eval_critic(po, z)
        v = agent.value_mean_std(v_n, True)
        pa = agent._preproc_amp_obs(amp)
        logit = net.eval_disc(pa)
        enc = net.eval_enc(pa)
        dr = agent._calc_disc_rewards(amp)
        er = agent._calc_enc_rewards(amp, z)
    out = dict(param_seed=31, obs=obs, amp=amp, z=z, obs_norm=po, amp_norm=pa, mu=mu, value_normed=v_n, value=v, disc_logit=logit, enc=enc,
               disc_r=dr, enc_r=er,
               rms={k: {kk: vv.clone() for kk, vv in ck[k].items()} for k in ('running_mean_std', 'reward_mean_std', 'amp_input_mean_std')})
    torch.save(out, os.path.join(OUT, 'inference_shipped_stats.pt'))
    print('inference_shipped_stats.pt ok', float(ck['amp_input_mean_std']['running_var'].min()), float(po.abs().max()))


def gen_rollout_math():
    agent, _ = rh.make_ref_agent('ase', num_envs=8, overrides={'minibatch_size': 256, 'amp_minibatch_size': 64})
    g = torch.Generator().manual_seed(21)
    H, N = 32, 8
    fd = (torch.rand(H, N, generator=g) < 0.1).float()
    v = torch.randn(H, N, 1, generator=g); nv = torch.randn(H, N, 1, generator=g); r = torch.rand(H, N, 1, generator=g)
    adv = agent.discount_values(fd, v, r, nv)
    mask = (torch.rand(H * N, generator=g) < 0.8).float()
    ret = O.swap_and_flatten01(adv + v); vals = O.swap_and_flatten01(v)
    advs_n = agent._calc_advs({'returns': ret, 'values': vals, 'rand_action_mask': mask})
    logits = torch.randn(64, 1, generator=g) * 4
    agent._eval_disc = lambda x: logits
    dr = agent._calc_disc_rewards(None)
    z = torch.nn.functional.normalize(torch.randn(64, 64, generator=g), dim=-1)
    ep = torch.nn.functional.normalize(torch.randn(64, 64, generator=g), dim=-1)
    agent._eval_enc = lambda x: ep
    er = agent._calc_enc_rewards(None, z)
    comb = agent._combine_rewards(torch.zeros_like(dr), {'disc_rewards': dr, 'enc_rewards': er})
    out = dict(fdones=fd, values=v, next_values=nv, rewards=r, advs=adv, mask=mask, advs_norm=advs_n,
               logits=logits, disc_r=dr, z=z, enc_pred=ep, enc_r=er, combined=comb)
    torch.save(out, os.path.join(OUT, 'rollout_math.pt'))
    print('rollout_math.pt ok')


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if len(sys.argv) > 1:           # python oracle/gen_golden.py gen_calc_gradients_full gen_rollout ...
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    gen_obs()
    gen_motion_lib()
    gen_heading()
    gen_rollout_math()
    gen_hrl_calc_gradients()
    gen_calc_gradients()
    # round 2 (each can also be run by name, see above): the reference's own play_steps / prepare_dataset over a scripted env, the early-termination
    # test, the inference path under the shipped checkpoint's statistics, the shipped checkpoints' layouts, and calc_gradients at the
    # benchmarked size (minutes of CPU time)
    gen_rollout()
    gen_humanoid_reset()
    gen_inference_shipped_stats()
    gen_checkpoint_layout()
    gen_calc_gradients_full()
