import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')): sys.path.insert(0, p)
import torch, ase_oracle as O, golden_util as G
from ase_b200.motion_lib import MotionLib
fx = G.load('motion_lib.pt'); mt = O.synthetic_motion_tables(seed=fx['seed'])
ml = MotionLib(mt.gts, mt.grs, mt.lrs, mt.grvs, mt.gravs, mt.dvs, mt.lengths, mt.num_frames, mt.dts)
state = ml.get_motion_state(fx['ids'], fx['t0'])
for mine, ref, name in zip(state, fx['state'], ('root_pos', 'root_rot', 'dof_pos', 'root_vel', 'root_ang_vel', 'dof_vel', 'key_pos')):
    d = (mine.cpu() - ref).abs()
    print(name, float(d.max()), (d > 2e-5).nonzero()[:5].tolist())
    if float(d.max()) > 2e-5:
        i = (d.reshape(d.shape[0], -1).max(1)[0] > 2e-5).nonzero().flatten()[:4]
        for r in i.tolist(): print('  row', r, 'id', int(fx['ids'][r]), 't0', float(fx['t0'][r]), 'len', float(mt.lengths[fx['ids'][r]]), mine.cpu()[r].flatten()[:6].tolist(), ref[r].flatten()[:6].tolist())
demo = ml.build_amp_obs_demo(fx['ids'], fx['t0'], fx['sim_dt'], fx['steps'])
d = (demo.cpu() - fx['demo']).abs(); print('demo', float(d.max()), (d > 2e-5).nonzero()[:8].tolist())
r = 2
ids = fx['ids'][r].repeat(10); tt = fx['t0'][r] - fx['sim_dt'] * torch.arange(0, 10)
print('row', r, 'id', int(fx['ids'][r]), 'nf', int(mt.num_frames[fx['ids'][r]]), 'times', tt.tolist())
sg = ml.get_motion_state(ids, tt); so = O.get_motion_state(mt, ids, tt)
for a, b, name in zip(sg, so, ('root_pos', 'root_rot', 'dof_pos', 'root_vel', 'root_ang_vel', 'dof_vel', 'key_pos')):
    d = (a.cpu() - b).abs().reshape(10, -1); print(name, d.max(1)[0].tolist())
dg = demo.cpu()[r].reshape(10, 140); do = fx['demo'][r].reshape(10, 140)
dd = (dg - do).abs()
for s in range(10):
    bad = (dd[s] > 2e-5).nonzero().flatten().tolist()
    if bad: print('step', s, 'cols', bad, dg[s, bad].tolist(), do[s, bad].tolist())
# dof_pos of the offending joint and its round trip
j = 12; o = O.DOF_OFFSETS_SWORD_SHIELD[j]
print('dof_pos gpu', sg[2][:, o:o+3].cpu().tolist()); print('dof_pos ref', so[2][:, o:o+3].tolist())
d = (demo.cpu() - fx['demo']).abs()
idx = int(d.argmax()); r, c = idx // 1400, idx % 1400; s, cc = c // 140, c % 140
print('ARGMAX row', r, 'step', s, 'col', cc, 'gpu', float(demo.cpu()[r, c]), 'ref', float(fx['demo'][r, c]), 'id', int(fx['ids'][r]), 't0', float(fx['t0'][r]))
ids = fx['ids'][r].repeat(10); tt = fx['t0'][r] - fx['sim_dt'] * torch.arange(0, 10)
sg = ml.get_motion_state(ids, tt); so = O.get_motion_state(mt, ids, tt)
j = (cc - 13) // 6 if 13 <= cc < 91 else -1
print('joint', j, 'times', tt.tolist())
if j >= 0:
    o = O.DOF_OFFSETS_SWORD_SHIELD[j]; sz = O.DOF_OFFSETS_SWORD_SHIELD[j+1]-o
    print('dof gpu', sg[2][s, o:o+sz].cpu().tolist(), 'ref', so[2][s, o:o+sz].tolist())
    mlen, nfr, dt = mt.lengths[ids], mt.num_frames[ids], mt.dts[ids]
    phase = torch.clip(tt / mlen, 0, 1); f0 = (phase * (nfr - 1)).long(); print('f0', f0.tolist(), 'phase*(nf-1)', (phase*(nfr-1)).tolist())
    body = O.DOF_BODY_IDS_SWORD_SHIELD[j]; f0l = f0 + mt.length_starts[ids]; f1l = torch.min(f0+1, nfr-1) + mt.length_starts[ids]
    q0 = mt.lrs[f0l[s], body]; q1 = mt.lrs[f1l[s], body]; print('q0', q0.tolist(), 'q1', q1.tolist(), 'dot', float((q0*q1).sum()))
print('demo row gpu', demo.cpu()[r, s*140+13+6*max(j,0): s*140+19+6*max(j,0)].tolist()); print('demo row ref', fx['demo'][r, s*140+13+6*max(j,0): s*140+19+6*max(j,0)].tolist())
