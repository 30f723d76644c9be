"""TEST / BENCH INFRASTRUCTURE ONLY -- times the reference's hot path on the host CPU through the oracle port
(oracle/ase_oracle.py; the reference itself is Python and cannot travel to the GPU box, so kind = "port").
A bounded sample of the BASELINE config-3 epoch is timed and extrapolated:

    epoch = 32 x rollout_step(4096 envs: obs build + AMP obs build + actor + 2 x critic)
          + reward pass (131072 rows of disc + enc trunk)  + GAE / advantage normalisation
          + 48 x calc_gradients(B = 16384, Ba = 4096) incl. Adam

sample = 1 rollout step, a 4096-row reward slice, 1 minibatch update (each preceded by one un-timed warm-up of the
same call when `warm` is set)."""
import time

import torch

import ase_oracle as O
import synth


def pick_threads():
    """The GPU box advertises 128 logical CPUs but torch CPU kernels collapse when oversubscribed: probe a GEMM of the
    learner's shape with a few thread counts (<= the affinity mask) and keep the fastest.  This is the core count the
    CPU arm reports."""
    import os
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    cands = sorted({c for c in (avail, 64, 32, 16, 8) if 1 <= c <= avail}, reverse=True)
    a = torch.randn(4096, 1024); b = torch.randn(1024, 1024)
    best, best_t = cands[-1], float('inf')
    for c in cands:
        torch.set_num_threads(c)
        a @ b
        t0 = time.perf_counter()
        for _ in range(3):
            a @ b
        t = time.perf_counter() - t0
        if t < best_t:
            best, best_t = c, t
    torch.set_num_threads(best)
    return best


def fixed_threads():
    """The CPU arm's thread policy: min(32, cores in the affinity mask) -- fixed, stated in the JSON line (round 1 picked the count with a
    noisy probe: 16 / 32 / 64 across arms of one round, the number moved by 3x)."""
    import os
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    n = max(1, min(32, avail))
    torch.set_num_threads(n)
    return n


def make_state(num_envs=4096, minibatch=16384, amp_minibatch=4096, seed=0):
    cfg = dict(O.DEFAULT_CFG); cfg['amp_minibatch_size'] = amp_minibatch
    P = synth.params(O.ase_param_shapes(), seed=seed)
    st = O.LearnerState(P, 253, 1400, 'ase')
    s = synth.rigid_body_state(num_envs, seed=seed)
    g = torch.Generator().manual_seed(seed)
    z = torch.nn.functional.normalize(torch.randn(num_envs, 64, generator=g), dim=-1)
    d, new_z = synth.minibatch(st, cfg, minibatch, amp_minibatch, seed=seed + 1)
    return dict(cfg=cfg, st=st, s=s, z=z, noise=torch.randn(num_envs, 31, generator=g), mask=torch.ones(num_envs), hist=torch.zeros(num_envs, 10, 140),
                amp_rows=torch.randn(2 * num_envs, 1400, generator=g), z2=torch.cat([z, z]), d=d, new_z=new_z, num_envs=num_envs)


def fraction_step(S, fraction=16, horizon=32, mini_epochs=6, batches_per_epoch=8):
    """An exact 1 / fraction of one training epoch in the epoch's own proportions (fraction = 16: 2 of the 32 rollout steps, 8192 of the
    131072 reward rows, 3 of the 48 minibatch updates incl. Adam).  Returns its wall time in seconds."""
    st, s, z, cfg = S['st'], S['s'], S['z'], S['cfg']
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(horizon // fraction):
            obs = O.compute_humanoid_observations_max(s['body_pos'], s['body_rot'], s['body_vel'], s['body_ang_vel'], True, True)
            kp = s['body_pos'][:, O.KEY_BODY_IDS_SWORD_SHIELD]
            fr = O.build_amp_observations(s['body_pos'][:, 0], s['body_rot'][:, 0], s['body_vel'][:, 0], s['body_ang_vel'][:, 0],
                                          s['dof_pos'], s['dof_vel'], kp, True, True, O.DOF_OFFSETS_SWORD_SHIELD)
            O.amp_hist_step(S['hist'], fr)
            O.get_action_values(st, obs, z, S['noise'], S['mask'])
            O.eval_critic_unnorm(st, obs, z)
        rows = S['num_envs'] * horizon // fraction
        O.calc_amp_rewards(st, S['amp_rows'][:rows], S['z2'][:rows], cfg)
    for _ in range(mini_epochs * batches_per_epoch // fraction):
        O.calc_gradients(st, S['d'], cfg, S['new_z'])
    return time.perf_counter() - t0


def fraction_description(num_envs, horizon, minibatch, amp_minibatch, mini_epochs, fraction):
    nmb = mini_epochs * (num_envs * horizon // minibatch)
    return (f"1/{fraction} of an epoch, timed whole: {horizon // fraction} rollout steps ({num_envs} envs: obs + AMP obs build, actor + 2 critic passes), "
            f"{num_envs * horizon // fraction}-row reward pass, {nmb // fraction} minibatch updates (B={minibatch}, Ba={amp_minibatch}, incl. Adam) "
            f"= {num_envs * horizon // fraction} env-steps; no extrapolation")


def _t(fn, reps=1):
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def sample_epoch_seconds(num_envs=4096, horizon=32, minibatch=16384, amp_minibatch=4096, mini_epochs=6, threads=None,
                         warm=True, seed=0):
    if threads:
        torch.set_num_threads(threads)
    cfg = dict(O.DEFAULT_CFG); cfg['amp_minibatch_size'] = amp_minibatch
    P = synth.params(O.ase_param_shapes(), seed=seed)
    st = O.LearnerState(P, 253, 1400, 'ase')
    s = synth.rigid_body_state(num_envs, seed=seed)
    g = torch.Generator().manual_seed(seed)
    z = torch.nn.functional.normalize(torch.randn(num_envs, 64, generator=g), dim=-1)
    noise = torch.randn(num_envs, 31, generator=g)
    mask = torch.ones(num_envs)
    amp_hist = torch.zeros(num_envs, 10, 140)

    def rollout_step():
        with torch.no_grad():
            obs = O.compute_humanoid_observations_max(s['body_pos'], s['body_rot'], s['body_vel'], s['body_ang_vel'], True, True)
            kp = s['body_pos'][:, O.KEY_BODY_IDS_SWORD_SHIELD]
            fr = O.build_amp_observations(s['body_pos'][:, 0], s['body_rot'][:, 0], s['body_vel'][:, 0], s['body_ang_vel'][:, 0],
                                          s['dof_pos'], s['dof_vel'], kp, True, True, O.DOF_OFFSETS_SWORD_SHIELD)
            O.amp_hist_step(amp_hist, fr)
            O.get_action_values(st, obs, z, noise, mask)
            O.eval_critic_unnorm(st, obs, z)

    amp_rows = torch.randn(num_envs, 1400, generator=g)

    def reward_slice():
        with torch.no_grad():
            O.calc_amp_rewards(st, amp_rows, z, cfg)

    d, new_z = synth.minibatch(st, cfg, minibatch, amp_minibatch, seed=seed + 1)

    def minibatch_update():
        O.calc_gradients(st, d, cfg, new_z)

    if warm:
        rollout_step(); reward_slice()
    t_roll = _t(rollout_step)
    t_rew = _t(reward_slice)
    if warm:
        minibatch_update()
    t_mb = _t(minibatch_update)
    batch = num_envs * horizon
    n_mb = mini_epochs * (batch // minibatch)
    epoch = horizon * t_roll + (batch / num_envs) * t_rew + n_mb * t_mb
    return {'epoch_s': epoch, 'rollout_step_s': t_roll, 'reward_4096rows_s': t_rew, 'minibatch_s': t_mb,
            'env_steps_per_s': batch / epoch, 'threads': torch.get_num_threads(),
            'sample': f'1 rollout step ({num_envs} envs) + {num_envs}-row reward slice + 1 minibatch update (B={minibatch}, Ba={amp_minibatch}), '
                      f'extrapolated to {horizon} steps + {batch // num_envs} reward slices + {n_mb} minibatches'}


if __name__ == '__main__':
    print(sample_epoch_seconds())
