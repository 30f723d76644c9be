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
