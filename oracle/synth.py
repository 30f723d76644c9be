"""TEST INFRASTRUCTURE ONLY -- seeded synthetic inputs shared by oracle/gen_golden.py, tests/ and
bench.py's CPU legs (SURVEY.md section 8(d) "Synthetic inputs").  torch CPU generators only, so the same
seed regenerates the same tensors on the GPU box (same image, same torch)."""
import math
import torch
import torch.nn.functional as F

import ase_oracle as O


def rigid_body_state(n, seed=0, bodies=17, dofs=31, edge_cases=False):
    g = torch.Generator().manual_seed(seed)
    pos = torch.randn(n, bodies, 3, generator=g)
    pos[:, 0, 2] = 0.5 + 0.7 * torch.rand(n, generator=g)
    rot = F.normalize(torch.randn(n, bodies, 4, generator=g), dim=-1)
    vel = torch.randn(n, bodies, 3, generator=g)
    ang = torch.randn(n, bodies, 3, generator=g)
    dof_pos = torch.rand(n, dofs, generator=g) * 2 - 1
    dof_vel = torch.randn(n, dofs, generator=g) * 2
    if edge_cases and n >= 8:
        rot[0, :] = torch.tensor([0., 0., 0., 1.])          # identity
        rot[1, 0] = torch.tensor([0., 0., 1., 0.])          # heading = pi
        rot[2, 0] = F.normalize(torch.tensor([1., 0., 0., 1.]), dim=0)   # pitch-free roll
        dof_pos[3, :] = 0.0                                 # zero exp-map -> default-axis branch
        dof_pos[4, 0:3] = torch.tensor([1e-6, 0., 0.])      # |angle| <= 1e-5 branch
        dof_pos[5, 0:3] = torch.tensor([3.5, 0.5, -0.2])    # angle > pi -> normalize_angle wraps
        dof_pos[6, 9] = 4.0                                 # 1-dof joint beyond pi
    return dict(body_pos=pos, body_rot=rot, body_vel=vel, body_ang_vel=ang, dof_pos=dof_pos, dof_vel=dof_vel)


def params(shapes, seed=0):
    p = O.synthetic_params(shapes, seed=seed)
    g = torch.Generator().manual_seed(seed + 1000)
    # push some action means beyond +-1 so bound loss and the diversity clip are exercised
    p['mu.bias'] = (torch.rand(p['mu.bias'].shape, generator=g) * 2 - 1) * 1.3
    return p


def minibatch(st, cfg, B, Ba, seed, kind='ase', obs_dim=253, amp_dim=1400, act=31, zdim=64):
    """One minibatch dict with the reference's key names (ase_agent.py:162-186).  old_logp_actions is
    set near the current policy's neglogp (ratio in ~[0.7,1.4]) so that clipped and unclipped samples,
    both signs of advantage, masked rows and out-of-bound means all occur."""
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(B, obs_dim, generator=g) * 1.5 + 0.3
    z = F.normalize(torch.randn(B, zdim, generator=g), dim=-1) if kind == 'ase' else None
    new_z = F.normalize(torch.randn(B, zdim, generator=g), dim=-1) if kind == 'ase' else None
    with torch.no_grad():
        rms = st.obs_rms.clone(); rms.update(obs)
        mu_cur = O.eval_actor(st.p, rms.norm(obs), z, mu_tanh=cfg.get('mu_tanh', False))
    sig = math.exp(-2.9)
    actions = mu_cur + sig * torch.randn(B, act, generator=g)
    logstd = torch.full((act,), -2.9)
    nlp = O.neglogp(actions, mu_cur, logstd)
    d = dict(old_values=torch.randn(B, 1, generator=g),
             old_logp_actions=nlp + 0.15 * torch.randn(B, generator=g),
             advantages=torch.randn(B, generator=g),
             mu=mu_cur + 0.02 * torch.randn(B, act, generator=g),
             sigma=torch.full((B, act), sig),
             returns=torch.randn(B, 1, generator=g),
             actions=actions, obs=obs)
    if kind != 'ppo':
        d['amp_obs'] = torch.randn(B, amp_dim, generator=g)
        d['amp_obs_replay'] = torch.randn(B, amp_dim, generator=g) * 1.2 - 0.1
        d['amp_obs_demo'] = torch.randn(B, amp_dim, generator=g) * 0.8 + 0.3
        d['rand_action_mask'] = (torch.rand(B, generator=g) < 0.85).float()
    if kind == 'ase':
        d['ase_latents'] = z
    return d, new_z
