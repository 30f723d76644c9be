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
    agent.train()
    assert [r['epoch'] for r in agent.epoch_log] == [1, 2, 3, 4, 5, 6]      # stops when epoch_num > max_epochs (common_agent.py:149)
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
