"""One config-3 rollout (4096 envs x 32 steps + reward pass + GAE) launched EAGERLY (no CUDA graph) between cudaProfilerStart / Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rollout_launches.csv python tools/profile_rollout.py
Without ncu it prints the CUDA-event time of the eager and of the graph-replayed rollout."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from ase_b200 import configs
from ase_b200.agent import ASEAgent
from ase_b200.synthetic_env import SyntheticHumanoidEnv

env = SyntheticHumanoidEnv(4096, device='cuda', seed=0)
cfg = configs.make('ase', device='cuda:0', vec_env=env, num_actors=4096, print_stats=False, gemm_backend=2, mini_epochs=1)
ag = ASEAgent('p', cfg); ag.init_tensors(); ag.obs = ag.env_reset(); ag._init_train()
def rollout():
    ag.set_eval()
    with torch.no_grad():
        ag.play_steps()
    ag.set_train()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for _ in range(4):
    rollout()
torch.cuda.synchronize()
ev[0].record(); rollout(); ev[1].record()
ag.set_graphs(False)
rollout(); torch.cuda.synchronize()
torch.cuda.profiler.start()
ev[2].record(); rollout(); ev[3].record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"rollout: graph replay {ev[0].elapsed_time(ev[1]):.2f} ms, eager {ev[2].elapsed_time(ev[3]):.2f} ms")
