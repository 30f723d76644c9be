"""GPU parity tests (through the C ABI) for the HBM-bound kernels: observation build, RunningMeanStd,
GAE, AMP/ASE rewards, advantage normalisation.  Checker = oracle/ase_oracle.py (CPU fp32) + the reference's
own outputs in tests/golden.  Tolerance: 1e-4 relative (north_star), with an absolute floor of 1e-5 on
O(1) quantities; these kernels are expected to land within a few fp32 ulps."""
import pytest
import torch

import ase_oracle as O
import synth
import golden_util as G

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 1e-5


def _pack(s, dev='cuda'):
    return torch.cat([s['body_pos'], s['body_rot'], s['body_vel'], s['body_ang_vel']], dim=-1).to(dev).contiguous()


def test_obs_build_golden_and_oracle():
    from ase_b200 import ops
    fx = G.load('obs_build.pt')
    s = fx['inputs']
    bs = _pack(s)
    for lro in (True, False):
        for rho in (True, False):
            o = ops.compute_humanoid_observations_max(bs, lro, rho).cpu()
            assert torch.allclose(o, fx[f'obs_l{int(lro)}_h{int(rho)}'], rtol=RTOL, atol=ATOL)
            buf = torch.zeros(64, 10, 140, device='cuda')
            ops.build_amp_observations(bs, s['dof_pos'].cuda(), s['dof_vel'].cuda(), buf, lro, rho, shift_history=False)
            assert torch.allclose(buf[:, 0].cpu(), fx[f'amp_l{int(lro)}_h{int(rho)}'], rtol=RTOL, atol=ATOL)
            assert float(buf[:, 1:].abs().max()) == 0.0


def test_obs_build_full_size_strided_and_subset():
    from ase_b200 import ops
    n = 4096
    s = synth.rigid_body_state(n, seed=5)
    # Isaac Gym layout [N, bodies_per_env, 13] with extra bodies: the view [:, :17] is strided (humanoid.py:86-89)
    full = torch.zeros(n, 19, 13)
    full[:, :17] = torch.cat([s['body_pos'], s['body_rot'], s['body_vel'], s['body_ang_vel']], dim=-1)
    view = full.cuda()[:, :17]
    ref = O.compute_humanoid_observations_max(s['body_pos'], s['body_rot'], s['body_vel'], s['body_ang_vel'], True, True)
    out = ops.compute_humanoid_observations_max(view, True, True)
    assert torch.allclose(out.cpu(), ref, rtol=RTOL, atol=ATOL)
    # reset path: only env_ids rows are rewritten
    ids = torch.tensor([3, 17, 4095, 1000], dtype=torch.int32, device='cuda')
    out2 = torch.full((n, 253), 7.0, device='cuda')
    ops.compute_humanoid_observations_max(view, True, True, out=out2, env_ids=ids)
    assert torch.allclose(out2[ids.long()].cpu(), ref[ids.long().cpu()], rtol=RTOL, atol=ATOL)
    mask = torch.ones(n, dtype=torch.bool); mask[ids.long().cpu()] = False
    assert float((out2.cpu()[mask] - 7.0).abs().max()) == 0.0
    # empty subset is a no-op
    ops.compute_humanoid_observations_max(view, True, True, out=out2, env_ids=torch.zeros(0, dtype=torch.int32, device='cuda'))


def test_amp_obs_history_shift_matches_reference_semantics():
    from ase_b200 import ops
    n = 300
    buf = torch.zeros(n, 10, 140, device='cuda')
    ref = torch.zeros(n, 10, 140)
    for t in range(12):    # more steps than history slots
        s = synth.rigid_body_state(n, seed=100 + t)
        bs = _pack(s)
        ops.build_amp_observations(bs, s['dof_pos'].cuda(), s['dof_vel'].cuda(), buf, True, True, shift_history=True)
        kp = s['body_pos'][:, O.KEY_BODY_IDS_SWORD_SHIELD]
        fr = O.build_amp_observations(s['body_pos'][:, 0], s['body_rot'][:, 0], s['body_vel'][:, 0], s['body_ang_vel'][:, 0],
                                      s['dof_pos'], s['dof_vel'], kp, True, True, O.DOF_OFFSETS_SWORD_SHIELD)
        O.amp_hist_step(ref, fr)
    assert torch.allclose(buf.cpu(), ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize('rows,cols', [(256, 253), (4096, 1400), (131072, 1), (2, 7)])
def test_running_mean_std_train_eval_unnorm(rows, cols):
    from ase_b200 import ops
    g = torch.Generator().manual_seed(rows + cols)
    r_gpu = ops.RunningMeanStd(cols, 'cuda')
    r_cpu = O.RMS(cols)
    for it in range(3):
        x = torch.randn(rows, cols, generator=g) * (1.0 + it) + 0.5 * it
        x[:, 0] = 3.0 + 1e-6 * torch.randn(rows, generator=g)     # near-constant column (tiny variance, like the checkpoint's 1.3e-11)
        y = r_gpu(x.cuda())
        y_ref = r_cpu.train_forward(x)
        assert torch.allclose(r_gpu.running_mean.cpu(), r_cpu.mean, rtol=1e-6, atol=1e-7)
        assert torch.allclose(r_gpu.running_var.cpu(), r_cpu.var, rtol=1e-5, atol=1e-10)
        assert float(r_gpu.count) == float(r_cpu.count)
        # the normalised output of a near-constant column amplifies the fp32 rounding of the batch mean: compare the others tightly
        assert torch.allclose(y.cpu()[:, 1:], y_ref[:, 1:], rtol=RTOL, atol=ATOL)
    r_gpu.eval()
    x = torch.randn(rows, cols, generator=g) * 20
    assert torch.allclose(r_gpu(x.cuda()).cpu()[:, 1:], r_cpu.norm(x)[:, 1:], rtol=RTOL, atol=ATOL)
    assert torch.allclose(r_gpu(x.cuda(), unnorm=True).cpu(), r_cpu.unnorm(x), rtol=RTOL, atol=ATOL)
    assert float(r_gpu.count) == float(r_cpu.count)      # eval mode must not update


def test_rollout_math_golden():
    from ase_b200 import ops
    fx = G.load('rollout_math.pt')
    adv = ops.discount_values(fx['fdones'].to(torch.uint8).cuda(), fx['values'].cuda(), fx['rewards'].cuda(), fx['next_values'].cuda(), 0.99, 0.95)
    assert torch.allclose(adv.cpu(), fx['advs'], rtol=RTOL, atol=ATOL)
    ret = O.swap_and_flatten01(fx['advs'] + fx['values']); vals = O.swap_and_flatten01(fx['values'])
    a = ops.calc_advs(ret.cuda(), vals.cuda(), fx['mask'].cuda())
    assert torch.allclose(a.cpu(), fx['advs_norm'], rtol=RTOL, atol=ATOL)
    dr, er, comb = ops.amp_rewards(fx['logits'].cuda(), fx['enc_pred'].cuda(), fx['z'].cuda(), 2.0, 1.0)
    assert torch.allclose(dr.cpu(), fx['disc_r'], rtol=RTOL, atol=ATOL)
    assert torch.allclose(er.cpu(), fx['enc_r'], rtol=RTOL, atol=ATOL)
    assert torch.allclose(comb.cpu(), fx['combined'], rtol=RTOL, atol=ATOL)


def test_gae_full_size_properties():
    """BASELINE size [32, 4096]: oracle comparison + the dones=1 reset property (advantage == delta where done)."""
    from ase_b200 import ops
    g = torch.Generator().manual_seed(3)
    H, N = 32, 4096
    d = (torch.rand(H, N, generator=g) < 1 / 30).to(torch.uint8)
    v = torch.randn(H, N, 1, generator=g); nv = torch.randn(H, N, 1, generator=g); r = torch.rand(H, N, 1, generator=g)
    adv, ret = ops.discount_values(d.cuda(), v.cuda(), r.cuda(), nv.cuda(), 0.99, 0.95, want_returns=True)
    ref = O.discount_values(d.float(), v, r, nv, 0.99, 0.95)
    assert torch.allclose(adv.cpu(), ref, rtol=RTOL, atol=ATOL)
    assert torch.allclose(ret.cpu(), ref + v, rtol=RTOL, atol=ATOL)
    delta = r + 0.99 * nv - v
    done = d.bool().unsqueeze(-1)
    assert torch.allclose(adv.cpu()[done], delta[done], rtol=1e-6, atol=1e-6)
    # unmasked advantage normalisation = (x-mean)/(std+1e-8)
    a = ops.calc_advs(O.swap_and_flatten01(ret.cpu()).cuda(), O.swap_and_flatten01(v).cuda(), None).cpu()
    assert abs(float(a.mean())) < 1e-4 and abs(float(a.std()) - 1.0) < 1e-3
    assert torch.allclose(a, O.calc_advs(O.swap_and_flatten01(ref + v), O.swap_and_flatten01(v), None), rtol=1e-3, atol=1e-4)


def test_heading_task_kernels_golden():
    from ase_b200 import ops
    fx = G.load('heading.pt')
    root = fx['root'].cuda()
    obs = ops.compute_heading_observations(root, fx['tar_dir'].cuda(), fx['tar_speed'].cuda(), fx['tar_face_dir'].cuda())
    assert torch.allclose(obs.cpu(), fx['obs'], rtol=RTOL, atol=ATOL)
    # written behind the humanoid features of a wider observation buffer (humanoid_amp_task.py:51-64)
    buf = torch.zeros(64, 258, device='cuda')
    ops.compute_heading_observations(root, fx['tar_dir'].cuda(), fx['tar_speed'].cuda(), fx['tar_face_dir'].cuda(), out=buf, col0=253)
    assert torch.allclose(buf[:, 253:].cpu(), fx['obs'], rtol=RTOL, atol=ATOL) and float(buf[:, :253].abs().max()) == 0.0
    rew = ops.compute_heading_reward(root[:, 0:3], fx['prev'].cuda(), root[:, 3:7], fx['tar_dir'].cuda(), fx['tar_speed'].cuda(),
                                     fx['tar_face_dir'].cuda(), 1.0 / 30.0)
    assert torch.allclose(rew.cpu(), fx['reward'], rtol=RTOL, atol=ATOL)


def test_motion_lib_state_and_demo_obs_golden():
    """ase_motion_state / ase_amp_obs_demo vs the reference MotionLib + build_amp_observations outputs (tests/golden/motion_lib.pt)
    on the same synthetic clip tables; incl. a 2-frame clip, a query beyond the clip end and an identity joint rotation."""
    from ase_b200.motion_lib import MotionLib
    fx = G.load('motion_lib.pt')
    mt = O.synthetic_motion_tables(seed=fx['seed'])
    ml = MotionLib(mt.gts, mt.grs, mt.lrs, mt.grvs, mt.gravs, mt.dvs, mt.lengths, mt.num_frames, mt.dts)
    state = ml.get_motion_state(fx['ids'], fx['t0'])
    # tolerance: 1e-4 of the quantity's natural scale -- exponential-map joint angles live on a pi scale (acos / atan2 chains
    # after a slerp: measured 8.6e-5 absolute = 3e-5 of pi), tangent / normal vectors and quaternions on a unit scale
    for mine, ref, name in zip(state, fx['state'], ('root_pos', 'root_rot', 'dof_pos', 'root_vel', 'root_ang_vel', 'dof_vel', 'key_pos')):
        assert torch.allclose(mine.cpu(), ref, rtol=RTOL, atol=3e-4 if name == 'dof_pos' else 2e-5), name
    demo = ml.build_amp_obs_demo(fx['ids'], fx['t0'], fx['sim_dt'], fx['steps'])
    assert torch.allclose(demo.cpu(), fx['demo'], rtol=RTOL, atol=1e-4)
    # sampler plumbing: shapes, times inside [truncate, len]
    d = ml.fetch_amp_obs_demo(512, 1.0 / 30.0, 10)
    assert d.shape == (512, 1400) and torch.isfinite(d).all()


def test_gather_rows_bit_exact():
    """ase_gather_rows (AMPDataset._get_item, amp_datasets.py:14-27): index work is bit-exact; several tensors per launch, vector and
    scalar paths (row widths 253 / 31 / 1 / 1400 / 64), identity index, empty batch, more tensors than one launch holds."""
    from ase_b200 import ops
    g = torch.Generator().manual_seed(7)
    n, rows = 5000, 777
    idx = torch.randint(0, n, (rows,), generator=g).cuda()
    srcs = [torch.randn(n, c, generator=g).cuda() for c in (253, 31, 1400, 64)] + [torch.randn(n, generator=g).cuda(), torch.randn(n, 1, generator=g).cuda()]
    dsts = [torch.full((rows,) + tuple(s.shape[1:]), -7.0, device='cuda') for s in srcs]
    ops.gather_rows([(s, d, idx) for s, d in zip(srcs, dsts)])
    for s, d in zip(srcs, dsts):
        assert torch.equal(d, s[idx])
    ident = torch.empty(rows, 253, device='cuda')
    ops.gather_rows([(srcs[0], ident, None)])
    assert torch.equal(ident, srcs[0][:rows])
    ops.gather_rows([])
    many = [(srcs[1], torch.empty(rows, 31, device='cuda'), idx) for _ in range(19)]       # > ASE_GATHER_MAX: split over two launches
    ops.gather_rows(many)
    assert all(torch.equal(d, srcs[1][idx]) for _, d, _ in many)
