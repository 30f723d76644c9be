"""One full-size ASE minibatch update (B=16384, Ba=4096) repeated N times -- the unit ncu captures.
  python tools/profile_minibatch.py [n_iters] [backend]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import torch
from ase_b200 import Learner, lib as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
backend = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B, Ba = 16384, 4096
g = torch.Generator(device='cuda').manual_seed(0)
r = lambda *s: torch.randn(*s, device='cuda', generator=g)
ln = Learner('ase', 253, 31, B, amp_dim=1400, latent_dim=64, amp_batch=Ba, gemm_backend=backend)
ln.init_reference(0)
if os.environ.get('ASE_GRADS_IPC', '0') == '1':       # experiment: gradient arena in a cudaMalloc'ed, IPC-exported allocation (what dist_utils.init_peer uses)
    import ctypes as C
    from ase_b200.dist_utils import _DeviceArray
    local, handle = C.c_void_p(), (C.c_uint8 * 64)()
    L.check(L.lib.ase_peer_alloc(ln.grads.numel(), C.byref(local), handle), 'ase_peer_alloc')
    _arr = _DeviceArray(L.lib.ase_peer_grads(local), (ln.grads.numel() + 3) // 4 * 4)
    ln.use_grads_arena(torch.as_tensor(_arr, device='cuda'))
    print('gradient arena: IPC-exported cudaMalloc allocation')
z = torch.nn.functional.normalize(r(B, 64), dim=-1)
d = dict(obs=r(B, 253), actions=r(B, 31) * 0.1, old_logp_actions=r(B) * 0.1 - 46, advantages=r(B), mu=r(B, 31) * 0.1,
         sigma=torch.full((B, 31), 0.055, device='cuda'), returns=r(B, 1), old_values=r(B, 1), rand_action_mask=(torch.rand(B, device='cuda') < 0.9).float(),
         ase_latents=z, amp_obs=r(Ba, 1400), amp_obs_replay=r(Ba, 1400), amp_obs_demo=r(Ba, 1400))
nz = torch.nn.functional.normalize(r(B, 64), dim=-1)
torch.cuda.synchronize()
c0 = L.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(n):
    if i == n - 1:
        e0.record()
    ln.calc_gradients(d, nz)
    ln.adam_step()
e1.record()
torch.cuda.synchronize()
print(f"launches per minibatch: {(L.launch_count() - c0) // n}; last minibatch {e0.elapsed_time(e1):.3f} ms "
      f"({0.9685 / (e0.elapsed_time(e1) / 1e3):.1f} TFLOP/s algorithmic)")
