"""GPU tests of the agent-level host mirror (play_steps / prepare_dataset / train_epoch) on the synthetic env.
Parity is checked by recomputing, with the oracle on CPU, everything play_steps derived from what it stored in the
experience buffer (values, rewards, GAE returns, advantages), so no RNG injection is needed."""
import pytest
import torch

import ase_oracle as O

pytestmark = pytest.mark.gpu


def _small_agent(kind='ase', n=64, h=8, mb=128, amb=32):
    from ase_b200 import configs
    from ase_b200.agent import ASEAgent, AMPAgent
    from ase_b200.synthetic_env import SyntheticHumanoidEnv
    env = SyntheticHumanoidEnv(n, device='cuda', seed=3, done_prob=0.05, demo_pool=512, demo_source='motion_lib' if kind == 'ase' else 'pool')
    cfg = configs.make(kind, device='cuda:0', vec_env=env, num_actors=n, horizon_length=h, minibatch_size=mb, amp_minibatch_size=amb,
                       mini_epochs=2, amp_obs_demo_buffer_size=2048, amp_replay_buffer_size=2048, amp_batch_size=64, print_stats=False)
    cfg['net_params']['mlp']['units'] = [128, 64] if kind == 'amp' else [128, 96, 64]
    cfg['net_params']['disc']['units'] = [128, 64] if kind == 'amp' else [128, 96, 64]
    agent = (ASEAgent if kind == 'ase' else AMPAgent)('t', cfg)
    agent.init_tensors()
    agent.obs = agent.env_reset()
    agent._init_train()
    return agent, env


def _oracle_state(agent, kind):
    P = {k: v.detach().cpu().clone() for k, v in agent.model.named_parameters().items()}
    P['sigma'] = agent.model.sigma.cpu().clone()
    st = O.LearnerState(P, 253, 1400, kind)
    for r, src in ((st.obs_rms, agent.model.running_mean_std), (st.val_rms, agent.model.value_mean_std), (st.amp_rms, agent.model.amp_input_mean_std)):
        r.mean, r.var, r.count = src.running_mean.cpu().clone(), src.running_var.cpu().clone(), src.count.cpu().clone()
    return st


@pytest.mark.parametrize('kind', ['ase', 'amp'])
def test_play_steps_matches_oracle_recomputation(kind):
    torch.manual_seed(0)
    agent, env = _small_agent(kind)
    # give the normalisers non-trivial statistics first
    agent.update_epoch(); agent.train_epoch()
    st = _oracle_state(agent, kind)
    with torch.no_grad():
        bd = agent.play_steps()
    eb = {k: v.cpu() for k, v in agent.experience_buffer.items()}
    H, N = eb['obses'].shape[:2]
    z = eb['ase_latents'] if kind == 'ase' else None
    with torch.no_grad():
        for t in (0, H - 1):
            zt = None if z is None else z[t]
            on = st.obs_rms.norm(eb['obses'][t])
            assert torch.allclose(eb['mus'][t], O.eval_actor(st.p, on, zt), rtol=1e-4, atol=1e-4)
            assert torch.allclose(eb['values'][t], st.val_rms.unnorm(O.eval_critic(st.p, on, zt)), rtol=1e-4, atol=1e-4)
            nlp = O.neglogp(eb['actions'][t], eb['mus'][t], st.p['sigma'])
            stochastic = eb['rand_action_mask'][t] == 1.0
            assert torch.allclose(eb['neglogpacs'][t][stochastic], nlp[stochastic], rtol=1e-4, atol=1e-3)
            assert torch.equal(eb['actions'][t][~stochastic], eb['mus'][t][~stochastic])     # eps-greedy rows act deterministically
        amp = eb['amp_obs'].reshape(H * N, -1)
        dr, er = O.calc_amp_rewards(st, amp, None if z is None else z.reshape(H * N, -1), dict(O.DEFAULT_CFG))
        w = (0.5, 0.5) if kind == 'ase' else (1.0, 0.0)
        rew = w[0] * dr + (w[1] * er if er is not None else 0)
        assert torch.allclose(bd['disc_rewards'].cpu(), O.swap_and_flatten01(dr.reshape(H, N, 1)), rtol=1e-4, atol=1e-4)
        adv = O.discount_values(eb['dones'].float(), eb['values'], rew.reshape(H, N, 1), eb['next_values'], 0.99, 0.95)
        assert torch.allclose(bd['returns'].cpu(), O.swap_and_flatten01(adv + eb['values']), rtol=1e-4, atol=1e-4)
    # env-major flatten (swap_and_flatten01): row = env * H + t
    assert torch.equal(bd['obses'][5 * H + 3].cpu(), eb['obses'][3, 5])
    agent.set_train()
    agent.prepare_dataset(bd)
    adv_n = O.calc_advs(bd['returns'].cpu(), bd['values'].cpu(), bd['rand_action_mask'].cpu())
    assert torch.allclose(agent.dataset.values_dict['advantages'].cpu(), adv_n, rtol=1e-3, atol=1e-4)


def test_train_epochs_run_and_stay_finite_and_checkpoint_roundtrip():
    torch.manual_seed(1)
    agent, env = _small_agent('ase')
    for _ in range(3):
        agent.update_epoch()
        info = agent.train_epoch()
    for k, v in info.items():
        assert torch.isfinite(v).all(), k
    assert agent.model.step == 3 * 2 * (64 * 8 // 128)
    assert float(agent.model.amp_input_mean_std.count) == 1 + agent.model.step * 3 * 32      # 3 AMP batches per minibatch
    assert float(agent.model.running_mean_std.count) == 1 + agent.model.step * 128
    assert float(agent.model.value_mean_std.count) == 1 + 3 * 2 * 512                       # two updates per epoch on the full batch
    w = agent.get_full_state_weights()
    assert 'a2c_network._enc_mlp.0.weight' in w['model'] and w['model']['a2c_network.sigma'].shape == (31,)
    assert w['running_mean_std']['running_mean'].dtype == torch.float64 and 'amp_input_mean_std' in w and 'reward_mean_std' in w
    agent2, _ = _small_agent('ase')
    agent2.set_full_state_weights(w)
    for k, v in agent.model.named_parameters().items():
        assert torch.equal(v, agent2.model.named_parameters()[k])
    assert torch.equal(agent2.model.exp_avg, agent.model.exp_avg) and agent2.model.step == agent.model.step


def test_train_loop_logs_through_the_async_ring():
    """agent.train() (common_agent.py:82-155) never reads a device scalar inside the loop: the epochs' train_result series and event
    timings come back through AsyncEpochLog (pinned ring) -- all epochs accounted for, in order, finite."""
    torch.manual_seed(2)
    agent, env = _small_agent('ase')
    agent.max_epochs = 5
    tags = []
    class W:
        def add_scalar(self, tag, value, step):
            tags.append(tag)
    agent.writer = W()
    calls = []
    class Observer:                                   # what run.py passes as config['features']['observer'] (RLGPUAlgoObserver)
        def after_print_stats(self, frame, epoch_num, total_time):
            calls.append((frame, epoch_num, total_time))
    agent.algo_observer = Observer()
    agent.train()
    assert [r['epoch'] for r in agent.epoch_log] == [1, 2, 3, 4, 5, 6]      # stops when epoch_num > max_epochs (common_agent.py:149)
    assert [c[1] for c in calls] == [1, 2, 3, 4, 5, 6] and all(b[2] >= a[2] for a, b in zip(calls, calls[1:]))
    # episodes finish at 5 % per env-step here: the per-epoch episode meter (game_rewards / game_lengths) must have reported some
    done = [r for r in agent.epoch_log if 'mean_rewards' in r]
    assert done and all(r['mean_lengths'] >= 1.0 and r['mean_rewards'] == r['mean_rewards'] for r in done)
    assert 'rewards0/frame' in tags and 'episode_lengths/iter' in tags and 'info/epochs' in tags
    for r in agent.epoch_log:
        assert r['play_time'] > 0 and r['update_time'] > 0 and r['frames'] == 64 * 8
        assert all(v == v and abs(v) < 1e9 for v in r['scalars'].values())
    assert 'performance/total_fps' in tags and 'losses/disc_loss' in tags and 'info/kl' in tags


def test_hrl_agent_epoch_config5_shapes():
    """BASELINE config 5 (HumanoidHeading task-train over a frozen ASE LLC): play_steps + update run end to end; the
    combined reward, LLC stepping and tanh-mu HLC learner are checked against the oracle on the stored buffers."""
    from ase_b200 import configs
    from ase_b200.agent import HRLAgent
    from ase_b200.synthetic_env import SyntheticHumanoidEnv
    torch.manual_seed(2)
    n, h = 64, 8
    env = SyntheticHumanoidEnv(n, device='cuda', seed=4, done_prob=0.05, demo_pool=256, heading_task=True)
    cfg = configs.make('hrl', device='cuda:0', vec_env=env, num_actors=n, horizon_length=h, minibatch_size=128, mini_epochs=2, print_stats=False)
    cfg['net_params']['mlp']['units'] = [128, 64]
    cfg['llc_net_params'] = {'mlp': {'units': [128, 96, 64]}, 'disc': {'units': [128, 96, 64]}}
    ag = HRLAgent('t', cfg)
    assert ag.actions_num == 64 and ag.obs_shape[0] == 258
    ag.init_tensors(); ag.obs = ag.env_reset()
    ag.update_epoch(); info = ag.train_epoch()
    for k, v in info.items():
        assert torch.isfinite(v).all(), k
    with torch.no_grad():
        bd = ag.play_steps()
    eb = {k: v.cpu() for k, v in ag.experience_buffer.items()}
    assert float(eb['mus'].abs().max()) <= 1.0                       # tanh'd means
    rew = 0.9 * eb['rewards'] + 0.1 * eb['disc_rewards']
    adv = O.discount_values(eb['dones'].float(), eb['values'], rew, eb['next_values'], 0.99, 0.95)
    assert torch.allclose(bd['returns'].cpu(), O.swap_and_flatten01(adv + eb['values']), rtol=1e-4, atol=1e-4)
    assert float(eb['disc_rewards'].min()) >= 0.0 and float(eb['rewards'].max()) <= 1.0 + 1e-6     # heading reward in [0,1]


# ------------------------------------------------------------------------------------------------------------------------------------
# drop-in surface: the rl_games factories of run.py and the checkpoints the reference ships
# ------------------------------------------------------------------------------------------------------------------------------------
def _checkpoint_from_layout(name, seed):
    """A random checkpoint with EXACTLY the layout of a shipped reference checkpoint (tests/golden/checkpoint_layout.json, written by
    oracle/gen_golden.py gen_checkpoint_layout from ase/data/models/*.pth): same keys, shapes, dtypes, optimizer state indices."""
    import json, os
    lay = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'checkpoint_layout.json')))[name]
    g = torch.Generator().manual_seed(seed)
    dt = {'torch.float32': torch.float32, 'torch.float64': torch.float64}
    w = {}
    for k, v in lay.items():
        if k == 'optimizer':
            st = {}
            for i, e in v['state'].items():
                st[int(i)] = {'step': 6194400, 'exp_avg': 1e-3 * torch.randn(e['exp_avg'][0], generator=g),
                              'exp_avg_sq': 1e-6 * torch.rand(e['exp_avg_sq'][0], generator=g) + 1e-9}
            w[k] = {'state': st, 'param_groups': [{'lr': 2e-5, 'betas': (0.9, 0.999), 'eps': 1e-8, 'weight_decay': 0.0, 'amsgrad': False,
                                                   'params': v['param_groups'][0]['params']}]}
        elif isinstance(v, dict):
            w[k] = {}
            for kk, (shape, dtype) in v.items():
                t = torch.randn(shape, generator=g, dtype=dt[dtype]) * (0.05 if dtype == 'torch.float32' else 1.0)
                if kk in ('running_var', 'count'):
                    t = t.abs() + 0.5
                if kk.endswith('sigma'):
                    t = torch.full(shape, -2.9)
                w[k][kk] = t
        else:
            w[k] = {'epoch': 129050, 'frame': 8355840000, 'last_mean_rewards': -100500, 'env_state': None}[k]
    # the ASE builder aliases the encoder trunk to the discriminator trunk (ase_network_builder.py:289-303): one storage, two names
    for kk in list(w['model']):
        if '._enc_mlp.' in kk:
            w['model'][kk] = w['model'][kk.replace('._enc_mlp.', '._disc_mlp.')]
    return w, lay


def _same_layout(w, lay):
    assert set(w.keys()) == set(lay.keys()), (sorted(w.keys()), sorted(lay.keys()))
    for k, v in lay.items():
        if k == 'optimizer':
            assert sorted(w[k]['state'].keys()) == sorted(int(i) for i in v['state'])
            for i, e in v['state'].items():
                st = w[k]['state'][int(i)]
                assert set(st.keys()) == set(e.keys()) and isinstance(st['step'], int)
                assert list(st['exp_avg'].shape) == e['exp_avg'][0] and str(st['exp_avg_sq'].dtype) == e['exp_avg_sq'][1]
            pg, pl = w[k]['param_groups'][0], v['param_groups'][0]
            assert set(pg.keys()) == set(pl.keys()) and list(pg['params']) == pl['params']
        elif isinstance(v, dict):
            assert list(w[k].keys()) == list(v.keys()), k                      # same names in the same order
            for kk, (shape, dtype) in v.items():
                assert list(w[k][kk].shape) == shape and str(w[k][kk].dtype) == dtype, (k, kk)


def _full_size_env_cfg(kind, n=64, h=8, **over):
    from ase_b200 import configs
    from ase_b200.synthetic_env import SyntheticHumanoidEnv
    env = SyntheticHumanoidEnv(n, device='cuda', seed=3, done_prob=0.05, demo_pool=512, heading_task=(kind == 'hrl'))
    kw = dict(device='cuda:0', vec_env=env, num_actors=n, horizon_length=h, minibatch_size=128, mini_epochs=1, print_stats=False)
    if kind != 'hrl':
        kw.update(amp_minibatch_size=32, amp_obs_demo_buffer_size=2048, amp_replay_buffer_size=2048, amp_batch_size=64)
    kw.update(over)
    return env, configs.make(kind, **kw)


def test_restore_accepts_the_shipped_checkpoint_layout_and_save_writes_it_back(tmp_path):
    """restore(fn) (common_agent.py:157-170, what run.py does for --checkpoint) on a file with the exact layout of the shipped
    ase_llc_reallusion_sword_shield.pth: 39 model tensors incl. the `_enc_mlp` aliases and the frozen sigma, f64 RMS buffers, Adam state
    1..32 with an int `step`; then save() must write the same layout back with the same numbers."""
    from ase_b200.agent import ASEAgent
    w, lay = _checkpoint_from_layout('ase_llc_reallusion_sword_shield', seed=7)
    fn = str(tmp_path / 'shipped_layout.pth')
    torch.save(w, fn)
    env, cfg = _full_size_env_cfg('ase')
    ag = ASEAgent('t', cfg)
    ag.restore(fn)
    assert ag.epoch_num == 129050 and ag.frame == 8355840000 and ag.model.step == 6194400
    sd = ag.model.state_dict()
    for k, v in w['model'].items():
        assert torch.equal(sd[k].cpu(), v), k
    assert torch.equal(ag.model.amp_input_mean_std.running_var.cpu(), w['amp_input_mean_std']['running_var'])
    assert float(ag.model.value_mean_std.count) == float(w['reward_mean_std']['count'])
    out = ag.save(str(tmp_path / 'resaved'))
    w2 = torch.load(out, map_location='cpu', weights_only=False)
    _same_layout(w2, lay)
    for k in w['model']:
        assert torch.equal(w2['model'][k], w['model'][k]), k
    for i, st in w['optimizer']['state'].items():
        assert torch.equal(w2['optimizer']['state'][i]['exp_avg'], st['exp_avg']) and w2['optimizer']['state'][i]['step'] == st['step']
        assert torch.equal(w2['optimizer']['state'][i]['exp_avg_sq'], st['exp_avg_sq'])
    for r in ('running_mean_std', 'reward_mean_std', 'amp_input_mean_std'):
        for kk in w[r]:
            assert torch.equal(w2[r][kk], w[r][kk]), (r, kk)
    # and training continues from it
    ag.init_tensors(); ag.obs = ag.env_reset(); ag._init_train()
    ag.update_epoch(); info = ag.train_epoch()
    assert all(torch.isfinite(v).all() for v in info.values()) and ag.model.step == 6194400 + 64 * 8 // 128


def test_hrl_restores_llc_and_hlc_from_shipped_layouts(tmp_path):
    """hrl_agent.py:28-41,202-213: the LLC comes from `llc_checkpoint` (an ASE checkpoint: model + running_mean_std + amp_input_mean_std are
    used, frozen, eval mode); the HLC itself restores from an HLC checkpoint (13 model tensors, no AMP statistics)."""
    from ase_b200.agent import HRLAgent
    llc, _ = _checkpoint_from_layout('ase_llc_reallusion_sword_shield', seed=8)
    hlc, hlay = _checkpoint_from_layout('ase_hlc_heading_reallusion_sword_shield', seed=9)
    fl, fh = str(tmp_path / 'llc.pth'), str(tmp_path / 'hlc.pth')
    torch.save(llc, fl); torch.save(hlc, fh)
    env, cfg = _full_size_env_cfg('hrl', llc_checkpoint=fl)
    cfg['llc_net_params'] = {'mlp': {'units': [1024, 1024, 512]}, 'disc': {'units': [1024, 1024, 512]}}
    ag = HRLAgent('t', cfg)
    sd = ag._llc.state_dict()
    for k, v in llc['model'].items():
        assert torch.equal(sd[k].cpu(), v), k
    assert torch.equal(ag._llc.running_mean_std.running_mean.cpu(), llc['running_mean_std']['running_mean'])
    assert torch.equal(ag._llc.amp_input_mean_std.running_var.cpu(), llc['amp_input_mean_std']['running_var'])
    ag.restore(fh)
    sd = ag.model.state_dict()
    for k, v in hlc['model'].items():
        assert torch.equal(sd[k].cpu(), v), k
    w2 = torch.load(ag.save(str(tmp_path / 'hlc_resaved')), map_location='cpu', weights_only=False)
    _same_layout(w2, hlay)
    ag.init_tensors(); ag.obs = ag.env_reset()
    ag.update_epoch(); info = ag.train_epoch()
    assert all(torch.isfinite(v).all() for v in info.values())


def test_agent_is_created_and_trained_through_the_rl_games_factories(tmp_path):
    """The call sequence of ase/run.py:153-170 + rl_games Runner.run_train with the reference's registrations swapped for this package's
    agent: the algo factory builds the agent from **kwargs (base_name, config); config['network'] is the model
    object the model builder returns (its network_builder holds the YAML `network` section); there is no 'vec_env' key -- the agent creates
    the env through rl_games.common.vecenv from config['env_name'] like A2CBase does; then restore(checkpoint) and train()."""
    import os, sys
    shims = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', 'shims')
    if shims not in sys.path:
        sys.path.insert(0, shims)
    from rl_games.common import object_factory, vecenv
    from rl_games.algos_torch import network_builder
    from ase_b200 import configs
    from ase_b200.agent import ASEAgent
    from ase_b200.synthetic_env import SyntheticHumanoidEnv

    # --- run.py:147-150: vec-env registration under the config's env_name
    made = []
    def create_env(config_name, num_actors, **kw):
        made.append((config_name, num_actors, kw))
        return SyntheticHumanoidEnv(num_actors, device='cuda', seed=kw.get('seed', 0), done_prob=0.05, demo_pool=512, demo_source='motion_lib')
    vecenv.register('rlgpu', create_env)

    # --- run.py:153-170 build_alg_runner: factories by name
    algo_factory, model_factory, network_factory = object_factory.ObjectFactory(), object_factory.ObjectFactory(), object_factory.ObjectFactory()
    algo_factory.register_builder('ase', lambda **kwargs: ASEAgent(**kwargs))

    class ASEBuilderStub(network_builder.NetworkBuilder):       # stands in for learning/ase_network_builder.ASEBuilder: load() keeps the YAML section
        def load(self, params):
            self.params = params

    class ModelStub:                                             # stands in for ase_models.ModelASEContinuous(network)
        def __init__(self, network):
            self.network_builder = network
    network_factory.register_builder('ase', lambda **kwargs: ASEBuilderStub())
    model_factory.register_builder('ase', lambda network, **kwargs: ModelStub(network))

    # --- Runner.load: params = YAML `params`; ModelBuilder.load(params) -> config['network']
    base = configs.make('ase')
    yaml_params = {'algo': {'name': 'ase'}, 'model': {'name': 'ase'}, 'network': dict(base['net_params'], name='ase'),
                   'config': {k: v for k, v in base.items() if k != 'net_params'}}
    yaml_params['network']['mlp'] = dict(yaml_params['network']['mlp'], units=[128, 96, 64])
    yaml_params['network']['disc'] = dict(yaml_params['network']['disc'], units=[128, 96, 64])
    network = network_factory.create(yaml_params['network']['name'])
    network.load(yaml_params['network'])
    config = dict(yaml_params['config'])
    config['network'] = model_factory.create(yaml_params['model']['name'], network=network)
    config.update(env_name='rlgpu', env_config={'seed': 5}, num_actors=64, horizon_length=8, minibatch_size=128, amp_minibatch_size=32, mini_epochs=2,
                  amp_obs_demo_buffer_size=2048, amp_replay_buffer_size=2048, amp_batch_size=64, print_stats=False, device='cuda:0',
                  max_epochs=3, save_frequency=2, train_dir=str(tmp_path), full_experiment_name='dropin', name='Humanoid')
    assert 'vec_env' not in config and 'net_params' not in config

    # --- Runner.run_train
    agent = algo_factory.create(yaml_params['algo']['name'], base_name='run', config=config)
    assert made == [('rlgpu', 64, {'seed': 5})] and isinstance(agent, ASEAgent)
    assert agent.model.named_parameters()['actor_mlp._dense_layers.0.weight'].shape == (128, 253 + 64)      # the builder object's YAML section was used
    agent.train()
    nn_dir = os.path.join(str(tmp_path), 'dropin', 'nn')
    saved = sorted(os.listdir(nn_dir))
    assert 'Humanoid.pth' in saved, saved
    # a second agent created the same way resumes from the file (Runner: agent.restore(args['checkpoint']))
    agent2 = algo_factory.create('ase', base_name='run', config=dict(config, max_epochs=5))
    agent2.restore(os.path.join(nn_dir, 'Humanoid.pth'))
    assert agent2.epoch_num == agent.epoch_num and agent2.model.step == agent.model.step
    assert torch.equal(agent2.model.params, agent.model.params)
