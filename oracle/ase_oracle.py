"""TEST INFRASTRUCTURE ONLY -- CPU (torch fp32/f64) restatement of the reference's PPO + adversarial
update hot path (nv-tlabs/ASE @ 28952a3).  It is the *checker* for the CUDA engine in ase_b200/:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
it.  The product path never does (ase_b200 raises if its CUDA library is missing).

Parity status: PINNED.  Every function below is checked (tests/test_oracle_vs_reference.py, in the
build container) against the reference's own unmodified Python executed through oracle/shims, and
against the committed outputs of that run (tests/golden/*.pt, made by oracle/gen_golden.py).
The un-vendored dependencies (rl-games==1.1.4, isaacgym.torch_utils) are restated from recollection
of their public sources -- see oracle/shims/*/__init__.py -- and pinned by the shipped checkpoints'
identities (strict state_dict load, f64 RMS buffers, count arithmetic).

All file:line citations are relative to /root/reference/ase/.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------
# quaternion helpers (xyzw)                      isaacgym.torch_utils [public IsaacGymEnvs restated]
# --------------------------------------------------------------------------------------------------


def quat_mul(a, b):
    """Hamilton product, 8-multiply factorisation (isaacgym.torch_utils.quat_mul)."""
    x1, y1, z1, w1 = a.unbind(-1)
    x2, y2, z2, w2 = b.unbind(-1)
    ww = (z1 + x1) * (x2 + y2)
    yy = (w1 - y1) * (w2 + z2)
    zz = (w1 + y1) * (w2 - z2)
    xx = ww + yy + zz
    qq = 0.5 * (xx + (z1 - x1) * (x2 - y2))
    w = qq - ww + (z1 - y1) * (y2 - z2)
    x = qq - xx + (x1 + w1) * (x2 + w2)
    y = qq - yy + (w1 - x1) * (y2 + z2)
    z = qq - zz + (z1 + y1) * (w2 - x2)
    return torch.stack([x, y, z, w], dim=-1)


def quat_rotate(q, v):
    """v(2w^2-1) + 2w(q_v x v) + 2 q_v (q_v . v)   (isaacgym.torch_utils.quat_rotate)."""
    qw = q[..., 3:4]
    qv = q[..., :3]
    a = v * (2.0 * qw * qw - 1.0)
    b = torch.cross(qv, v, dim=-1) * qw * 2.0
    c = qv * (qv * v).sum(-1, keepdim=True) * 2.0
    return a + b + c


def _normalize(x, eps=1e-9):
    return x / x.norm(p=2, dim=-1).clamp(min=eps).unsqueeze(-1)


def quat_from_angle_axis(angle, axis):
    theta = (angle / 2).unsqueeze(-1)
    xyz = _normalize(axis) * theta.sin()
    return _normalize(torch.cat([xyz, theta.cos()], dim=-1))


def normalize_angle(x):
    return torch.atan2(torch.sin(x), torch.cos(x))


def quat_to_tan_norm(q):
    """utils/torch_utils.py:46-59: [rot(q, x^) || rot(q, z^)]."""
    ex = torch.zeros_like(q[..., :3]); ex[..., 0] = 1
    ez = torch.zeros_like(q[..., :3]); ez[..., 2] = 1
    return torch.cat([quat_rotate(q, ex), quat_rotate(q, ez)], dim=-1)


def calc_heading_quat_inv(q):
    """utils/torch_utils.py:117-154: rotation by -heading about z."""
    ex = torch.zeros_like(q[..., :3]); ex[..., 0] = 1
    d = quat_rotate(q, ex)
    heading = torch.atan2(d[..., 1], d[..., 0])
    ez = torch.zeros_like(q[..., :3]); ez[..., 2] = 1
    return quat_from_angle_axis(-heading, ez)


def exp_map_to_quat(e):
    """utils/torch_utils.py:68-91."""
    angle = torch.norm(e, dim=-1)
    axis = e / angle.unsqueeze(-1)
    angle = normalize_angle(angle)
    default_axis = torch.zeros_like(e); default_axis[..., 2] = 1
    mask = torch.abs(angle) > 1e-5
    angle = torch.where(mask, angle, torch.zeros_like(angle))
    axis = torch.where(mask.unsqueeze(-1), axis, default_axis)
    return quat_from_angle_axis(angle, axis)


# --------------------------------------------------------------------------------------------------
# observation build                                  env/tasks/humanoid.py, env/tasks/humanoid_amp.py
# --------------------------------------------------------------------------------------------------

DOF_OFFSETS_SWORD_SHIELD = [0, 3, 6, 9, 10, 13, 16, 17, 20, 21, 24, 27, 28, 31]   # humanoid.py:192
KEY_BODY_IDS_SWORD_SHIELD = [5, 10, 13, 16, 6, 9]  # r_hand l_hand r_foot l_foot sword shield (getup.yaml:20)


def compute_humanoid_observations_max(body_pos, body_rot, body_vel, body_ang_vel,
                                      local_root_obs, root_height_obs):
    """env/tasks/humanoid.py:591-635.  [N,J,3],[N,J,4],[N,J,3],[N,J,3] -> [N, 1+(J-1)*3+J*6+J*3+J*3]."""
    N, J, _ = body_pos.shape
    root_pos = body_pos[:, 0]
    root_rot = body_rot[:, 0]
    h = root_pos[:, 2:3]
    if not root_height_obs:
        h = torch.zeros_like(h)
    hq = calc_heading_quat_inv(root_rot).unsqueeze(1).expand(N, J, 4)
    lp = quat_rotate(hq, body_pos - root_pos.unsqueeze(1)).reshape(N, J * 3)[:, 3:]
    lr = quat_to_tan_norm(quat_mul(hq, body_rot)).reshape(N, J * 6).clone()
    if local_root_obs:
        lr[:, 0:6] = quat_to_tan_norm(root_rot)          # humanoid.py:622-624 (global root rot, sic)
    lv = quat_rotate(hq, body_vel).reshape(N, J * 3)
    lw = quat_rotate(hq, body_ang_vel).reshape(N, J * 3)
    return torch.cat([h, lp, lr, lv, lw], dim=-1)


def dof_to_obs(pose, dof_offsets):
    """env/tasks/humanoid.py:522-552."""
    out = []
    for j in range(len(dof_offsets) - 1):
        o, sz = dof_offsets[j], dof_offsets[j + 1] - dof_offsets[j]
        if sz == 3:
            q = exp_map_to_quat(pose[:, o:o + 3])
        else:
            assert sz == 1
            axis = torch.tensor([0.0, 1.0, 0.0], dtype=pose.dtype).expand(pose.shape[0], 3)
            q = quat_from_angle_axis(pose[:, o], axis)
        out.append(quat_to_tan_norm(q))
    return torch.cat(out, dim=-1)


def build_amp_observations(root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, key_body_pos,
                           local_root_obs, root_height_obs, dof_offsets):
    """env/tasks/humanoid_amp.py:282-316 -> [N, 13 + 6*joints + dofs + 3*keys]."""
    N, K, _ = key_body_pos.shape
    h = root_pos[:, 2:3]
    if not root_height_obs:
        h = torch.zeros_like(h)
    hq = calc_heading_quat_inv(root_rot)
    rr = quat_mul(hq, root_rot) if local_root_obs else root_rot
    rr = quat_to_tan_norm(rr)
    lv = quat_rotate(hq, root_vel)
    lw = quat_rotate(hq, root_ang_vel)
    lk = quat_rotate(hq.unsqueeze(1).expand(N, K, 4), key_body_pos - root_pos.unsqueeze(1)).reshape(N, K * 3)
    return torch.cat([h, rr, lv, lw, dof_to_obs(dof_pos, dof_offsets), dof_vel, lk], dim=-1)


def calc_heading_quat(q):
    """utils/torch_utils.py:130-141."""
    ex = torch.zeros_like(q[..., :3]); ex[..., 0] = 1
    d = quat_rotate(q, ex)
    heading = torch.atan2(d[..., 1], d[..., 0])
    ez = torch.zeros_like(q[..., :3]); ez[..., 2] = 1
    return quat_from_angle_axis(heading, ez)


def compute_heading_observations(root_states, tar_dir, tar_speed, tar_face_dir):
    """env/tasks/humanoid_heading.py:232-248 -> [N,5]."""
    hq = calc_heading_quat_inv(root_states[:, 3:7])
    z = torch.zeros_like(tar_dir[..., 0:1])
    ltd = quat_rotate(hq, torch.cat([tar_dir, z], dim=-1))[..., 0:2]
    lfd = quat_rotate(hq, torch.cat([tar_face_dir, z], dim=-1))[..., 0:2]
    return torch.cat([ltd, tar_speed.unsqueeze(-1), lfd], dim=-1)


def compute_heading_reward(root_pos, prev_root_pos, root_rot, tar_dir, tar_speed, tar_face_dir, dt):
    """env/tasks/humanoid_heading.py:250-285."""
    root_vel = (root_pos - prev_root_pos) / dt
    tar_dir_speed = (tar_dir * root_vel[..., :2]).sum(-1)
    tangent_speed = (root_vel[..., :2] - tar_dir_speed.unsqueeze(-1) * tar_dir).sum(-1)
    verr = tar_speed - tar_dir_speed
    dir_reward = torch.exp(-0.25 * (verr * verr + 0.1 * tangent_speed * tangent_speed))
    dir_reward = torch.where(tar_dir_speed <= 0, torch.zeros_like(dir_reward), dir_reward)
    ex = torch.zeros_like(root_pos); ex[..., 0] = 1.0
    facing = quat_rotate(calc_heading_quat(root_rot), ex)
    facing_reward = torch.clamp_min((tar_face_dir * facing[..., 0:2]).sum(-1), 0.0)
    return 0.7 * dir_reward + 0.3 * facing_reward


# --------------------------------------------------------------------------------------------------
# motion library: (clip, time) -> interpolated reference state -> demo AMP observations      utils/motion_lib.py
# --------------------------------------------------------------------------------------------------
DOF_BODY_IDS_SWORD_SHIELD = [1, 2, 3, 4, 5, 7, 8, 11, 12, 13, 14, 15, 16]     # env/tasks/humanoid.py:191


def slerp(q0, q1, t):
    """utils/torch_utils.py:93-115."""
    cos_half = (q0 * q1).sum(-1)
    neg = cos_half < 0
    q1 = torch.where(neg.unsqueeze(-1), -q1, q1)
    cos_half = torch.abs(cos_half).unsqueeze(-1)
    half = torch.acos(cos_half)
    sin_half = torch.sqrt(1.0 - cos_half * cos_half)
    ra = torch.sin((1 - t) * half) / sin_half
    rb = torch.sin(t * half) / sin_half
    new_q = ra * q0 + rb * q1
    new_q = torch.where(torch.abs(sin_half) < 0.001, 0.5 * q0 + 0.5 * q1, new_q)
    new_q = torch.where(torch.abs(cos_half) >= 1, q0, new_q)
    return new_q


def quat_to_angle_axis(q):
    """utils/torch_utils.py:6-27."""
    sin_theta = torch.sqrt(1 - q[..., 3] * q[..., 3])
    angle = normalize_angle(2 * torch.acos(q[..., 3]))
    axis = q[..., 0:3] / sin_theta.unsqueeze(-1)
    mask = torch.abs(sin_theta) > 1e-5
    default_axis = torch.zeros_like(axis); default_axis[..., 2] = 1
    angle = torch.where(mask, angle, torch.zeros_like(angle))
    axis = torch.where(mask.unsqueeze(-1), axis, default_axis)
    return angle, axis


def quat_to_exp_map(q):
    """utils/torch_utils.py:37-44."""
    angle, axis = quat_to_angle_axis(q)
    return angle.unsqueeze(-1) * axis


class MotionTables:
    """The flat per-frame tensors MotionLib builds (utils/motion_lib.py:65-89,174-238): gts [F,J,3], grs/lrs [F,J,4],
    grvs/gravs [F,3], dvs [F,dofs]; per motion: length (s), num_frames, dt, first-frame offset."""

    def __init__(self, gts, grs, lrs, grvs, gravs, dvs, lengths, num_frames, dts):
        self.gts, self.grs, self.lrs, self.grvs, self.gravs, self.dvs = gts, grs, lrs, grvs, gravs, dvs
        self.lengths, self.num_frames, self.dts = lengths, num_frames, dts
        shifted = num_frames.roll(1); shifted[0] = 0
        self.length_starts = shifted.cumsum(0)


def synthetic_motion_tables(seed=0, frames=(40, 75, 23, 2), fps=(30.0, 30.0, 60.0, 30.0), bodies=17, dofs=31):
    """Smooth random clips in MotionLib's table format (incl. a 2-frame clip: the shortest a clip can be)."""
    g = torch.Generator().manual_seed(seed)
    gts, grs, lrs, grvs, gravs, dvs = [], [], [], [], [], []
    for nf in frames:
        def smooth_quat():
            q = torch.randn(1, bodies, 4, generator=g) + 0.15 * torch.randn(nf, bodies, 4, generator=g).cumsum(0)
            return F.normalize(q, dim=-1)
        gts.append(torch.randn(1, bodies, 3, generator=g) + 0.05 * torch.randn(nf, bodies, 3, generator=g).cumsum(0))
        gr = smooth_quat()
        # root: mostly-upright character (yaw + small tilt) as in real clips -- a root x-axis pointing straight up makes the
        # heading angle (atan2 of its xy projection) ill-conditioned for ANY fp32 implementation
        yaw = torch.rand(1, generator=g) * 6.28 + 0.1 * torch.randn(nf, generator=g).cumsum(0)
        tilt = 0.15 * torch.randn(nf, 2, generator=g)
        gr[:, 0] = F.normalize(torch.stack([tilt[:, 0], tilt[:, 1], torch.sin(yaw / 2), torch.cos(yaw / 2)], dim=-1), dim=-1)
        grs.append(gr); lrs.append(smooth_quat())
        grvs.append(torch.randn(nf, 3, generator=g)); gravs.append(torch.randn(nf, 3, generator=g)); dvs.append(torch.randn(nf, dofs, generator=g))
    nfr = torch.tensor(frames, dtype=torch.long)
    dts = torch.tensor([1.0 / f for f in fps], dtype=torch.float32)
    lengths = torch.tensor([1.0 / f * (n - 1) for f, n in zip(fps, frames)], dtype=torch.float32)
    lrs = torch.cat(lrs); lrs[5, 3] = torch.tensor([0., 0., 0., 1.])      # identity joint rotation: default-axis branch of quat_to_angle_axis
    return MotionTables(torch.cat(gts), torch.cat(grs), lrs, torch.cat(grvs), torch.cat(gravs), torch.cat(dvs), lengths, nfr, dts)


def get_motion_state(mt, motion_ids, motion_times, dof_body_ids=DOF_BODY_IDS_SWORD_SHIELD, dof_offsets=DOF_OFFSETS_SWORD_SHIELD,
                     key_body_ids=KEY_BODY_IDS_SWORD_SHIELD):
    """utils/motion_lib.py:123-172 (+ _calc_frame_blend :263-272, _local_rotation_to_dof :296-324)."""
    mlen, nfr, dt = mt.lengths[motion_ids], mt.num_frames[motion_ids], mt.dts[motion_ids]
    phase = torch.clip(motion_times / mlen, 0.0, 1.0)
    f0 = (phase * (nfr - 1)).long()
    f1 = torch.min(f0 + 1, nfr - 1)
    blend = ((motion_times - f0 * dt) / dt).unsqueeze(-1)
    f0l, f1l = f0 + mt.length_starts[motion_ids], f1 + mt.length_starts[motion_ids]
    root_pos = (1.0 - blend) * mt.gts[f0l, 0] + blend * mt.gts[f1l, 0]
    root_rot = slerp(mt.grs[f0l, 0], mt.grs[f1l, 0], blend)
    kid = torch.tensor(key_body_ids)
    be = blend.unsqueeze(-1)
    key_pos = (1.0 - be) * mt.gts[f0l.unsqueeze(-1), kid.unsqueeze(0)] + be * mt.gts[f1l.unsqueeze(-1), kid.unsqueeze(0)]
    local_rot = slerp(mt.lrs[f0l], mt.lrs[f1l], be)
    n = motion_ids.shape[0]
    dof_pos = torch.zeros(n, dof_offsets[-1])
    for j, body in enumerate(dof_body_ids):
        o, sz = dof_offsets[j], dof_offsets[j + 1] - dof_offsets[j]
        q = local_rot[:, body]
        if sz == 3:
            dof_pos[:, o:o + 3] = quat_to_exp_map(q)
        else:
            theta, axis = quat_to_angle_axis(q)
            dof_pos[:, o] = normalize_angle(theta * axis[..., 1])
    return root_pos, root_rot, dof_pos, mt.grvs[f0l], mt.gravs[f0l], mt.dvs[f0l], key_pos


def build_amp_obs_demo(mt, motion_ids, motion_times0, sim_dt, num_steps, local_root_obs=True, root_height_obs=True):
    """env/tasks/humanoid_amp.py:85-101: `num_steps` frames going back in time by sim_dt -> [n, num_steps * 140]."""
    n = motion_ids.shape[0]
    ids = motion_ids.unsqueeze(-1).expand(n, num_steps).reshape(-1)
    times = (motion_times0.unsqueeze(-1) - sim_dt * torch.arange(0, num_steps)).reshape(-1)
    rp, rr, dp, rv, rw, dv, kp = get_motion_state(mt, ids, times)
    obs = build_amp_observations(rp, rr, rv, rw, dp, dv, kp, local_root_obs, root_height_obs, DOF_OFFSETS_SWORD_SHIELD)
    return obs.reshape(n, -1)


def amp_hist_step(amp_buf, new_frame):
    """env/tasks/humanoid_amp.py:248-275: shift history by one slot, newest frame at slot 0.
    amp_buf [N,S,F] (modified in place), new_frame [N,F]."""
    amp_buf[:, 1:] = amp_buf[:, :-1].clone()
    amp_buf[:, 0] = new_frame
    return amp_buf


# --------------------------------------------------------------------------------------------------
# RunningMeanStd                                   rl_games 1.1.4 algos_torch/running_mean_std.py
# --------------------------------------------------------------------------------------------------

class RMS:
    """f64 running (mean, var, count); count starts at 1 (confirmed by shipped checkpoint)."""

    def __init__(self, size, eps=1e-5):
        self.mean = torch.zeros(size, dtype=torch.float64)
        self.var = torch.ones(size, dtype=torch.float64)
        self.count = torch.ones((), dtype=torch.float64)
        self.eps = eps

    def clone(self):
        r = RMS(self.mean.shape[0], self.eps)
        r.mean, r.var, r.count = self.mean.clone(), self.var.clone(), self.count.clone()
        return r

    def update(self, x):
        bm = x.mean(0)
        bv = x.var(0)            # unbiased
        n = x.shape[0]
        delta = bm - self.mean
        tot = self.count + n
        self.mean = self.mean + delta * n / tot
        m2 = self.var * self.count + bv * n + delta ** 2 * self.count * n / tot
        self.var = m2 / tot
        self.count = tot

    def norm(self, x):
        y = (x - self.mean.float()) / torch.sqrt(self.var.float() + self.eps)
        return torch.clamp(y, -5.0, 5.0)

    def unnorm(self, x):
        y = torch.clamp(x, -5.0, 5.0)
        return torch.sqrt(self.var.float() + self.eps) * y + self.mean.float()

    def train_forward(self, x):
        self.update(x)
        return self.norm(x)


# --------------------------------------------------------------------------------------------------
# networks                      learning/amp_network_builder.py, learning/ase_network_builder.py
# --------------------------------------------------------------------------------------------------

def ase_param_shapes(obs=253, z=64, act=31, amp=1400, units=(1024, 1024, 512), disc_units=(1024, 1024, 512),
                     style_units=(512, 256)):
    """Trainable tensors in `model.parameters()` order (= Adam param order).
    Names are the reference's state_dict names with the `a2c_network.` prefix removed."""
    s = OrderedDict()
    s['sigma'] = (act,)
    i = z
    for k, u in enumerate(style_units):
        s[f'actor_mlp._style_mlp.{2 * k}.weight'] = (u, i); s[f'actor_mlp._style_mlp.{2 * k}.bias'] = (u,); i = u
    s['actor_mlp._style_dense.weight'] = (z, i); s['actor_mlp._style_dense.bias'] = (z,)
    i = obs + z
    for k, u in enumerate(units):
        s[f'actor_mlp._dense_layers.{k}.weight'] = (u, i); s[f'actor_mlp._dense_layers.{k}.bias'] = (u,); i = u
    i = obs + z
    for k, u in enumerate(units):
        s[f'critic_mlp._mlp.{2 * k}.weight'] = (u, i); s[f'critic_mlp._mlp.{2 * k}.bias'] = (u,); i = u
    s['value.weight'] = (1, units[-1]); s['value.bias'] = (1,)
    s['mu.weight'] = (act, units[-1]); s['mu.bias'] = (act,)
    i = amp
    for k, u in enumerate(disc_units):
        s[f'_disc_mlp.{2 * k}.weight'] = (u, i); s[f'_disc_mlp.{2 * k}.bias'] = (u,); i = u
    s['_disc_logits.weight'] = (1, i); s['_disc_logits.bias'] = (1,)
    s['_enc.weight'] = (z, i); s['_enc.bias'] = (z,)
    return s


def amp_param_shapes(obs=253, act=31, amp=1400, units=(1024, 512), disc_units=(1024, 512)):
    """AMP / plain-PPO network (amp_network_builder.py:16-49 over rl_games A2CBuilder)."""
    s = OrderedDict()
    s['sigma'] = (act,)
    i = obs
    for k, u in enumerate(units):
        s[f'actor_mlp.{2 * k}.weight'] = (u, i); s[f'actor_mlp.{2 * k}.bias'] = (u,); i = u
    i = obs
    for k, u in enumerate(units):
        s[f'critic_mlp.{2 * k}.weight'] = (u, i); s[f'critic_mlp.{2 * k}.bias'] = (u,); i = u
    s['value.weight'] = (1, units[-1]); s['value.bias'] = (1,)
    s['mu.weight'] = (act, units[-1]); s['mu.bias'] = (act,)
    if amp:
        i = amp
        for k, u in enumerate(disc_units):
            s[f'_disc_mlp.{2 * k}.weight'] = (u, i); s[f'_disc_mlp.{2 * k}.bias'] = (u,); i = u
        s['_disc_logits.weight'] = (1, i); s['_disc_logits.bias'] = (1,)
    return s


def synthetic_params(shapes, seed=0, sigma_val=-2.9):
    """Deterministic stand-in for the reference initialisation (torch default Linear init
    U(+-1/sqrt(fan_in)) for weights, zero biases made slightly non-zero here so bias gradients are
    exercised; `_disc_logits` U(-1,1), `_enc` U(-0.1,0.1), `_style_dense` U(-1,1):
    amp_network_builder.py:112-120, ase_network_builder.py:196-210,326-336)."""
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, shp in shapes.items():
        if name == 'sigma':
            p[name] = torch.full(shp, float(sigma_val))
        elif name.endswith('.bias'):
            p[name] = (torch.rand(shp, generator=g) * 2 - 1) * 0.05
        else:
            bound = 1.0 / math.sqrt(shp[1])
            if name.startswith('_disc_logits') or name.startswith('actor_mlp._style_dense'):
                bound = 1.0
            elif name.startswith('_enc.'):
                bound = 0.1
            p[name] = (torch.rand(shp, generator=g) * 2 - 1) * bound
    return p


def _mlp(x, p, names, act_last=True):
    n = len(names)
    for k, nm in enumerate(names):
        x = F.linear(x, p[nm + '.weight'], p[nm + '.bias'])
        if act_last or k < n - 1:
            x = torch.relu(x)
    return x


def _layer_names(p, prefix, sep):
    ks = sorted({int(k[len(prefix):].split('.')[0]) for k in p if k.startswith(prefix) and k.endswith('.weight')})
    return [prefix + str(k) for k in ks]


def is_ase(p):
    return 'actor_mlp._style_dense.weight' in p


def eval_style(p, z):
    """ase_network_builder.py:317-321."""
    h = _mlp(z, p, _layer_names(p, 'actor_mlp._style_mlp.', '.'))
    return torch.tanh(F.linear(h, p['actor_mlp._style_dense.weight'], p['actor_mlp._style_dense.bias']))


def eval_actor(p, obs_n, z=None, mu_tanh=False):
    """ASE: ase_network_builder.py:123-144,305-324.  AMP: amp_network_builder.py:51-72.  HRL high-level policy:
    hrl_network_builder.py:26-29 (mu_tanh).  -> mu."""
    if is_ase(p):
        h = torch.cat([obs_n, eval_style(p, z)], dim=-1)
        h = _mlp(h, p, _layer_names(p, 'actor_mlp._dense_layers.', '.'))
    else:
        h = _mlp(obs_n, p, _layer_names(p, 'actor_mlp.', '.'))
    mu = F.linear(h, p['mu.weight'], p['mu.bias'])
    return torch.tanh(mu) if mu_tanh else mu


def eval_critic(p, obs_n, z=None):
    """ase_network_builder.py:115-121,255-259; amp_network_builder.py:74-79."""
    if is_ase(p):
        h = _mlp(torch.cat([obs_n, z], dim=-1), p, _layer_names(p, 'critic_mlp._mlp.', '.'))
    else:
        h = _mlp(obs_n, p, _layer_names(p, 'critic_mlp.', '.'))
    return F.linear(h, p['value.weight'], p['value.bias'])


def disc_trunk(p, x):
    return _mlp(x, p, _layer_names(p, '_disc_mlp.', '.'))


def eval_disc(p, x):
    """amp_network_builder.py:81-84."""
    return F.linear(disc_trunk(p, x), p['_disc_logits.weight'], p['_disc_logits.bias'])


def eval_enc(p, x):
    """ase_network_builder.py:214-219 (enc.separate False => trunk is the disc trunk, :202-203)."""
    return F.normalize(F.linear(disc_trunk(p, x), p['_enc.weight'], p['_enc.bias']), dim=-1)


def neglogp(a, mu, logstd):
    """rl_games ModelA2CContinuousLogStd.neglogp; sigma = exp(logstd)."""
    sigma = torch.exp(logstd)
    return 0.5 * (((a - mu) / sigma) ** 2).sum(-1) + 0.5 * math.log(2.0 * math.pi) * a.shape[-1] + logstd.sum(-1)


def policy_kl(mu0, s0, mu1, s1):
    """rl_games torch_ext.policy_kl (reduce=True)."""
    c1 = torch.log(s1 / s0 + 1e-5)
    c2 = (s0 ** 2 + (mu1 - mu0) ** 2) / (2.0 * (s1 ** 2 + 1e-5))
    return (c1 + c2 - 0.5).sum(-1).mean()


# --------------------------------------------------------------------------------------------------
# rollout-side math
# --------------------------------------------------------------------------------------------------

def discount_values(fdones, values, rewards, next_values, gamma, tau):
    """learning/common_agent.py:437-449.  All [H,N,1] except fdones [H,N]."""
    adv = torch.zeros_like(rewards)
    last = 0
    for t in reversed(range(rewards.shape[0])):
        nd = (1.0 - fdones[t]).unsqueeze(1)
        delta = rewards[t] + gamma * next_values[t] - values[t]
        last = delta + gamma * tau * nd * last
        adv[t] = last
    return adv


def disc_rewards(logits, scale):
    """learning/amp_agent.py:570-577."""
    prob = 1 / (1 + torch.exp(-logits))
    return -torch.log(torch.maximum(1 - prob, torch.tensor(0.0001))) * scale


def enc_rewards(enc_pred, z, scale):
    """learning/ase_agent.py:404-411,469-472."""
    err = -(enc_pred * z).sum(-1, keepdim=True)
    return torch.clamp_min(-err, 0.0) * scale


def normalization_with_masks(v, m):
    """rl_games torch_ext.normalization_with_masks."""
    sm = m.sum()
    vm = v * m
    mean = vm.sum() / sm
    min_sqr = ((vm ** 2) / sm).sum() - ((vm / sm).sum()) ** 2
    var = min_sqr * sm / (sm - 1)
    return (v - mean) / (torch.sqrt(var) + 1e-8)


def calc_advs(returns, values, mask=None):
    """amp_agent.py:551-561 (masked) / common_agent.py:536-546 (mask None: plain mean/std)."""
    adv = (returns - values).sum(1)
    if mask is None:
        return (adv - adv.mean()) / (adv.std() + 1e-8)
    return normalization_with_masks(adv, mask)


def swap_and_flatten01(x):
    s = x.shape
    return x.transpose(0, 1).reshape(s[0] * s[1], *s[2:])


# --------------------------------------------------------------------------------------------------
# calc_gradients                     learning/{common,amp,ase}_agent.py
# --------------------------------------------------------------------------------------------------

DEFAULT_CFG = dict(   # data/cfg/train/rlg/ase_humanoid.yaml:59-114
    e_clip=0.2, critic_coef=5.0, entropy_coef=0.0, bounds_loss_coef=10.0, disc_coef=5.0,
    disc_logit_reg=0.01, disc_grad_penalty=5.0, disc_weight_decay=1e-4, enc_coef=5.0,
    amp_diversity_bonus=0.01, amp_diversity_tar=1.0, lr=2e-5, beta1=0.9, beta2=0.999, adam_eps=1e-8,
    gamma=0.99, tau=0.95, disc_reward_scale=2.0, enc_reward_scale=1.0,
    task_reward_w=0.0, disc_reward_w=0.5, enc_reward_w=0.5, amp_minibatch_size=4096,
)


class LearnerState:
    """params (name->tensor), Adam moments, RMS stats.  kind in {'ase','amp','ppo'}."""

    def __init__(self, params, obs_dim, amp_dim=0, kind='ase'):
        self.kind = kind
        self.p = OrderedDict((k, v.clone()) for k, v in params.items())
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in params.items() if k != 'sigma')
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in params.items() if k != 'sigma')
        self.step = 0
        self.obs_rms = RMS(obs_dim)
        self.val_rms = RMS(1)
        self.amp_rms = RMS(amp_dim) if amp_dim else None


def calc_gradients(st, d, cfg, new_z=None, apply_adam=True, world=1):
    """One minibatch update.
    ASE:  learning/ase_agent.py:159-308 (+ _enc_loss :413-443, _diversity_loss :445-467)
    AMP:  learning/amp_agent.py:266-390 (+ _disc_loss :442-479)
    PPO:  learning/common_agent.py:353-435 (losses :456-464,505-534)
    d: minibatch dict with the reference's key names; new_z: the latents `_diversity_loss` would draw.
    Returns (train_result dict, grads dict)."""
    kind = st.kind
    p = OrderedDict((k, v.detach().clone().requires_grad_(k != 'sigma')) for k, v in st.p.items())
    obs_n = st.obs_rms.train_forward(d['obs'])
    Ba = cfg['amp_minibatch_size']
    if kind != 'ppo':
        amp_a = st.amp_rms.train_forward(d['amp_obs'][0:Ba])
        amp_r = st.amp_rms.train_forward(d['amp_obs_replay'][0:Ba])
        amp_d = st.amp_rms.train_forward(d['amp_obs_demo'][0:Ba]).requires_grad_(True)
        mask = d['rand_action_mask']
        msum = mask.sum()
    z = d.get('ase_latents') if kind == 'ase' else None
    mu = eval_actor(p, obs_n, z, mu_tanh=cfg.get('mu_tanh', False))
    logstd = mu * 0.0 + p['sigma']
    sigma = torch.exp(logstd)
    values = eval_critic(p, obs_n, z)
    nlp = neglogp(d['actions'], mu, logstd)
    entropy = (0.5 + 0.5 * math.log(2 * math.pi) + logstd).sum(-1)

    ratio = torch.exp(d['old_logp_actions'] - nlp)
    adv = d['advantages']
    a_loss = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1.0 - cfg['e_clip'], 1.0 + cfg['e_clip']))
    clipped = (torch.abs(ratio - 1.0) > cfg['e_clip']).float()
    c_loss = ((d['returns'] - values) ** 2).mean()                       # clip_value False
    b_loss = (torch.clamp_max(mu + 1.0, 0.0) ** 2 + torch.clamp_min(mu - 1.0, 0.0) ** 2).sum(-1)
    res = {}
    if kind == 'ppo':
        a_loss, b_loss, entropy, clip_frac = a_loss.mean(), b_loss.mean(), entropy.mean(), clipped.mean()
    else:
        a_loss = (mask * a_loss).sum() / msum
        entropy = (mask * entropy).sum() / msum
        b_loss = (mask * b_loss).sum() / msum
        clip_frac = (mask * clipped).sum() / msum
    loss = a_loss + cfg['critic_coef'] * c_loss - cfg['entropy_coef'] * entropy + cfg['bounds_loss_coef'] * b_loss

    if kind != 'ppo':
        la, lr_, ld = eval_disc(p, amp_a), eval_disc(p, amp_r), eval_disc(p, amp_d)
        lcat = torch.cat([la, lr_], dim=0)
        bce = F.binary_cross_entropy_with_logits
        disc_loss = 0.5 * (bce(lcat, torch.zeros_like(lcat)) + bce(ld, torch.ones_like(ld)))
        logit_w = p['_disc_logits.weight'].flatten()
        logit_loss = (logit_w ** 2).sum()
        disc_loss = disc_loss + cfg['disc_logit_reg'] * logit_loss
        g = torch.autograd.grad(ld, amp_d, grad_outputs=torch.ones_like(ld), create_graph=True,
                                retain_graph=True, only_inputs=True)[0]
        gp = (g ** 2).sum(-1).mean()
        disc_loss = disc_loss + cfg['disc_grad_penalty'] * gp
        if cfg['disc_weight_decay'] != 0:
            ws = [p[n + '.weight'].flatten() for n in _layer_names(p, '_disc_mlp.', '.')] + [logit_w]
            disc_loss = disc_loss + cfg['disc_weight_decay'] * (torch.cat(ws) ** 2).sum()
        loss = loss + cfg['disc_coef'] * disc_loss
        res.update(disc_loss=disc_loss, disc_grad_penalty=gp.detach(), disc_logit_loss=logit_loss.detach(),
                   disc_agent_acc=(lcat < 0).float().mean(), disc_demo_acc=(ld > 0).float().mean(),
                   disc_agent_logit=lcat.detach(), disc_demo_logit=ld.detach())
    if kind == 'ase':
        enc_pred = eval_enc(p, amp_a)
        enc_loss = (-(enc_pred * z[0:Ba]).sum(-1, keepdim=True)).mean()
        loss = loss + cfg['enc_coef'] * enc_loss
        res['enc_loss'] = enc_loss
        if cfg['amp_diversity_bonus'] != 0:
            mu2 = eval_actor(p, obs_n, new_z)
            a_diff = ((torch.clamp(mu, -1.0, 1.0) - torch.clamp(mu2, -1.0, 1.0)) ** 2).mean(-1)
            z_diff = 0.5 - 0.5 * (new_z * z).sum(-1)
            div = (cfg['amp_diversity_tar'] - a_diff / (z_diff + 1e-5)) ** 2
            div = (mask * div).sum() / msum
            loss = loss + cfg['amp_diversity_bonus'] * div
            res['amp_diversity_loss'] = div

    names = [k for k in p if k != 'sigma']
    gl = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    grads = OrderedDict((k, (g_ if g_ is not None else torch.zeros_like(p[k])) / world) for k, g_ in zip(names, gl))
    if apply_adam:
        adam_step(st, grads, cfg)
    kl = policy_kl(mu.detach(), sigma.detach(), d['mu'], d['sigma'])
    res.update(entropy=entropy, kl=kl, b_loss=b_loss, actor_loss=a_loss, actor_clip_frac=clip_frac,
               critic_loss=c_loss, loss=loss.detach(), mus=mu.detach(), values=values.detach())
    return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in res.items()}, grads


def adam_step(st, grads, cfg):
    """torch.optim.Adam(lr, betas, eps, weight_decay=0, amsgrad=False) (common_agent.py:45)."""
    st.step += 1
    b1, b2, lr, eps = cfg['beta1'], cfg['beta2'], cfg['lr'], cfg['adam_eps']
    bc1 = 1 - b1 ** st.step
    bc2 = 1 - b2 ** st.step
    for k, g in grads.items():
        st.m[k].mul_(b1).add_(g, alpha=1 - b1)
        st.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (st.v[k].sqrt() / math.sqrt(bc2)).add_(eps)
        st.p[k] = st.p[k] - (lr / bc1) * st.m[k] / denom


# --------------------------------------------------------------------------------------------------
# rollout-side inference                         learning/ase_agent.py:117-148,385-411
# --------------------------------------------------------------------------------------------------

def get_action_values(st, obs, z, noise, rand_mask):
    """Eval-mode policy step.  noise ~ N(0,1) [N,act] and rand_mask (bernoulli draw) are inputs.
    -> dict(actions, mus, sigmas, neglogpacs, values)."""
    p = st.p
    obs_n = st.obs_rms.norm(obs)
    mu = eval_actor(p, obs_n, z)
    logstd = mu * 0.0 + p['sigma']
    sigma = torch.exp(logstd)
    a = mu + sigma * noise
    nlp = neglogp(a, mu, logstd)
    v = st.val_rms.unnorm(eval_critic(p, obs_n, z))
    a = torch.where((rand_mask == 0.0).unsqueeze(-1), mu, a)
    return dict(actions=a, mus=mu, sigmas=sigma, neglogpacs=nlp, values=v)


def eval_critic_unnorm(st, obs, z):
    return st.val_rms.unnorm(eval_critic(st.p, st.obs_rms.norm(obs), z))


def calc_amp_rewards(st, amp_obs, z, cfg):
    """ase_agent.py:395-411 + amp_agent.py:570-577 -> (disc_r, enc_r)."""
    x = st.amp_rms.norm(amp_obs)
    dr = disc_rewards(eval_disc(st.p, x), cfg['disc_reward_scale'])
    er = enc_rewards(eval_enc(st.p, x), z, cfg['enc_reward_scale']) if st.kind == 'ase' else None
    return dr, er
