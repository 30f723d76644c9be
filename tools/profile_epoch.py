"""Phase breakdown of one training epoch (CUDA-event timers around the agent's phases)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from ase_b200 import configs, ops
from ase_b200.agent import ASEAgent
from ase_b200.synthetic_env import SyntheticHumanoidEnv

env = SyntheticHumanoidEnv(4096, device='cuda', seed=0)
cfg = configs.make('ase', device='cuda:0', vec_env=env, num_actors=4096, print_stats=False, gemm_backend=2)
ag = ASEAgent('p', cfg); ag.init_tensors(); ag.obs = ag.env_reset(); ag._init_train()
for _ in range(2):
    ag.update_epoch(); ag.train_epoch()
torch.cuda.synchronize()

T = {}
def timed(name, fn, *a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); r = fn(*a, **k); e1.record(); t1 = time.perf_counter()
    T.setdefault(name, []).append((e0, e1, t1 - t0)); return r

# wrap phases
for name in ('get_action_values', '_eval_critic', 'env_step', 'env_reset', '_final_rewards', '_pre_action'):
    orig = getattr(ag, name)
    setattr(ag, name, (lambda o, n: (lambda *a, **k: timed(n, o, *a, **k)))(orig, name))
od = ops.discount_values
ops.discount_values = lambda *a, **k: timed('gae', od, *a, **k)
for name in ('prepare_dataset', '_minibatch', '_pre_update', '_post_update'):
    orig = getattr(ag, name)
    setattr(ag, name, (lambda o, n: (lambda *a, **k: timed(n, o, *a, **k)))(orig, name))
ocg, oad = ag.model.calc_gradients, ag.model.adam_step
ag.model.calc_gradients = lambda *a, **k: timed('learner.calc_gradients', ocg, *a, **k)
ag.model.adam_step = lambda *a, **k: timed('learner.adam_step', oad, *a, **k)
ops_ps = ag.play_steps
ag.play_steps = lambda: timed('play_steps(total)', ops_ps)
ag.update_epoch(); t0 = time.perf_counter(); ag.train_epoch(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
p, u, tot = ag.epoch_times()
print(f"epoch wall {wall*1e3:.1f} ms; events: play {p*1e3:.1f} update {u*1e3:.1f}")
for k, v in T.items():
    g = sum(e0.elapsed_time(e1) for e0, e1, _ in v); c = sum(x[2] for x in v) * 1e3
    print(f"  {k:26s} n={len(v):4d}  gpu {g:8.2f} ms   host-issue {c:8.2f} ms")
