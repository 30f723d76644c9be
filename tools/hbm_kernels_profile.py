"""One short config-3 epoch (4096 envs x 32 steps, 8 minibatches) for `ncu` captures of the HBM-bound (non-GEMM) kernels:
   ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
       -k regex:'^(?!.*gemm_)' --csv --log-file gpurun_out/hbm_kernels.csv python tools/hbm_kernels_profile.py
The profiled region (cudaProfilerStart/Stop) is the third epoch; summarise with tools/summarize_hbm.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from ase_b200 import configs
from ase_b200.agent import ASEAgent
from ase_b200.synthetic_env import SyntheticHumanoidEnv

env = SyntheticHumanoidEnv(4096, device='cuda', seed=0)
cfg = configs.make('ase', device='cuda:0', vec_env=env, num_actors=4096, print_stats=False, gemm_backend=2, mini_epochs=1)
ag = ASEAgent('p', cfg); ag.init_tensors(); ag.obs = ag.env_reset(); ag._init_train()
for _ in range(2):
    ag.update_epoch(); ag.train_epoch()
torch.cuda.synchronize()
torch.cuda.profiler.start()
ag.update_epoch(); ag.train_epoch()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
