"""Build-container only: pins oracle/ase_oracle.py to the reference's own code, executed live through
oracle/shims (isaacgym / rl_games stand-ins).  Skipped where /root/reference is absent (GPU box)."""
import pytest
import torch

pytestmark = pytest.mark.reference


def test_obs_functions_live():
    import ref_harness as rh, ase_oracle as O, synth
    humanoid, humanoid_amp, _ = rh.import_env_fns()
    s = synth.rigid_body_state(257, seed=3, edge_cases=True)
    ref = humanoid.compute_humanoid_observations_max(s['body_pos'], s['body_rot'], s['body_vel'], s['body_ang_vel'], True, True)
    mine = O.compute_humanoid_observations_max(s['body_pos'], s['body_rot'], s['body_vel'], s['body_ang_vel'], True, True)
    assert torch.allclose(ref, mine, rtol=1e-5, atol=1e-5)
    kp = s['body_pos'][:, O.KEY_BODY_IDS_SWORD_SHIELD]
    ref = humanoid_amp.build_amp_observations(s['body_pos'][:, 0], s['body_rot'][:, 0], s['body_vel'][:, 0], s['body_ang_vel'][:, 0],
                                              s['dof_pos'], s['dof_vel'], kp, True, True, 78, O.DOF_OFFSETS_SWORD_SHIELD)
    mine = O.build_amp_observations(s['body_pos'][:, 0], s['body_rot'][:, 0], s['body_vel'][:, 0], s['body_ang_vel'][:, 0],
                                    s['dof_pos'], s['dof_vel'], kp, True, True, O.DOF_OFFSETS_SWORD_SHIELD)
    assert torch.allclose(ref, mine, rtol=1e-5, atol=1e-5)


def test_quat_shim_against_poselib():
    """isaacgym.torch_utils restatement vs the in-tree independent poselib implementation."""
    import ref_harness as rh
    rh.import_env_fns()
    from isaacgym import torch_utils as tu
    from poselib.poselib.core import rotation3d as r3
    g = torch.Generator().manual_seed(0)
    a = torch.nn.functional.normalize(torch.randn(100, 4, generator=g), dim=-1)
    b = torch.nn.functional.normalize(torch.randn(100, 4, generator=g), dim=-1)
    v = torch.randn(100, 3, generator=g)
    assert torch.allclose(tu.quat_mul(a, b), r3.quat_mul(a, b), atol=1e-6)
    assert torch.allclose(tu.quat_rotate(a, v), r3.quat_rotate(a, v), atol=1e-5)


def test_checkpoint_loads_strictly_into_reference_and_oracle_forward_matches():
    import ref_harness as rh, ase_oracle as O
    ag, _ = rh.make_ref_agent('ase', num_envs=8, overrides={'minibatch_size': 256, 'amp_minibatch_size': 64})
    ck = torch.load('/root/reference/ase/data/models/ase_llc_reallusion_sword_shield.pth', weights_only=True, map_location='cpu')
    ag.model.load_state_dict(ck['model'], strict=True)
    assert ck['running_mean_std']['running_mean'].dtype == torch.float64
    assert float(ck['reward_mean_std']['count']) == 1 + 129050 * 2 * 131072
    assert float(ck['running_mean_std']['count']) == 1 + 129050 * 6 * 131072
    P = {k[len('a2c_network.'):]: v for k, v in ck['model'].items() if '_enc_mlp' not in k}
    g = torch.Generator().manual_seed(0)
    obs = torch.randn(32, 253, generator=g); z = torch.nn.functional.normalize(torch.randn(32, 64, generator=g), dim=-1)
    amp = torch.randn(32, 1400, generator=g)
    net = ag.model.a2c_network
    with torch.no_grad():
        mu_ref, _ = net.eval_actor(obs, z)
        assert torch.allclose(O.eval_actor(P, obs, z), mu_ref, atol=1e-5)
        assert torch.allclose(O.eval_critic(P, obs, z), net.eval_critic(obs, z), atol=1e-5)
        assert torch.allclose(O.eval_disc(P, amp), net.eval_disc(amp), atol=1e-4)
        assert torch.allclose(O.eval_enc(P, amp), net.eval_enc(amp), atol=1e-5)


def test_calc_gradients_live_ase():
    import ref_harness as rh, ase_oracle as O, synth
    B, Ba = 128, 32
    ag, _ = rh.make_ref_agent('ase', num_envs=4, overrides={'minibatch_size': B, 'amp_minibatch_size': Ba})
    P = synth.params(O.ase_param_shapes(), seed=2)
    sd = {'a2c_network.' + k: v.clone() for k, v in P.items()}
    for k in list(sd):
        if '_disc_mlp' in k:
            sd[k.replace('_disc_mlp', '_enc_mlp')] = sd[k]
    ag.model.load_state_dict(sd, strict=True)
    st = O.LearnerState(P, 253, 1400, 'ase')
    cfg = dict(O.DEFAULT_CFG); cfg['amp_minibatch_size'] = Ba
    for s in range(2):
        d, nz = synth.minibatch(st, cfg, B, Ba, seed=s)
        ag._sample_latents = lambda n, nz=nz: nz
        ag.calc_gradients(d)
        res, grads = O.calc_gradients(st, d, cfg, nz)
        for k in ('actor_loss', 'critic_loss', 'b_loss', 'disc_loss', 'enc_loss', 'amp_diversity_loss', 'kl', 'disc_grad_penalty'):
            assert abs(float(ag.train_result[k]) - float(res[k])) < 1e-5 * max(1, abs(float(res[k]))), k
        for n, prm in ag.model.named_parameters():
            k = n[len('a2c_network.'):]
            if k == 'sigma' or '_enc_mlp' in k:
                continue
            assert torch.allclose(prm.grad, grads[k], rtol=1e-4, atol=1e-7), k
            assert torch.allclose(prm.detach(), st.p[k], rtol=1e-6, atol=1e-7), k


def test_dataset_and_replay_buffer_host_logic_matches_reference_classes():
    """SURVEY 8f row 1 host logic, live against the reference's own classes on CPU: learning/amp_datasets.py AMPDataset (global
    permutation, contiguous slices, reshuffle when exhausted) and learning/replay_buffer.py ReplayBuffer (ring store with wrap-around,
    permutation sampling, modulo while the buffer is not full) produce the SAME row sequences as ase_b200's mirrors under one seed."""
    import ref_harness as rh
    rh._setup_path()
    from learning.amp_datasets import AMPDataset as RefDataset
    from learning.replay_buffer import ReplayBuffer as RefReplay
    from ase_b200.agent import AMPDataset
    from ase_b200.replay_buffer import ReplayBuffer
    B, mb = 96, 32
    x = torch.arange(B, dtype=torch.float32)
    torch.manual_seed(5)
    ref = RefDataset(B, mb, False, False, 'cpu', 4)
    ref.update_values_dict({'x': x})
    ref_seq = [ref._get_item(i)['x'].clone() for _ in range(3) for i in range(len(ref))]
    torch.manual_seed(5)
    mine = AMPDataset(B, mb, 'cpu')
    mine.update_values_dict({'x': x})
    my_seq = [x[mine.sample_indices(i)] for _ in range(3) for i in range(len(mine))]
    assert len(ref) == len(mine) == 3 and all(torch.equal(a, b) for a, b in zip(ref_seq, my_seq))

    torch.manual_seed(9)
    r = RefReplay(50, 'cpu')
    ref_out = []
    g = torch.Generator().manual_seed(1)
    chunks = [torch.randn(n, 3, generator=g) for n in (20, 20, 20, 7, 50, 13)]       # fills, wraps around, exact-size store
    for c in chunks:
        r.store({'amp_obs': c})
        ref_out.append(r.sample(16)['amp_obs'].clone())
    torch.manual_seed(9)
    m = ReplayBuffer(50, 'cpu')
    for c, ro in zip(chunks, ref_out):
        m.store({'amp_obs': c})
        idx = m.sample_indices(16)
        assert torch.equal(m.rows('amp_obs', idx), ro)
    assert m.get_total_count() == r.get_total_count() and m.get_buffer_size() == r.get_buffer_size()
