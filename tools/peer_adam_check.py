"""Run under torchrun with N >= 2 ranks on one node:  the fused allreduce + Adam kernel over NVLink peer memory (csrc/peer.cu) against (a) the
gradient arenas of all ranks added in rank order with plain torch adds followed by adam_kernel -- parameters, moments and summed gradients
must be BIT-identical, on every rank and across ranks -- and (b) NCCL's own allreduce (bit-identical for two ranks, to rounding for more:
NCCL adds in ring / tree order); then the timings of NCCL allreduce + adam_kernel against the fused kernel.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/peer_adam_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
from ase_b200 import Learner, dist_utils as D
D.init_comm()

def make():
    ln = Learner('ase', 253, 31, 256, amp_dim=1400, latent_dim=64, amp_batch=64, gemm_backend=0)      # full-size arena (7.04 M floats, odd tail)
    ln.init_reference(seed=3)
    return ln
a, b = make(), make()
on = D.init_peer(b)
print(f"rank {rank}: peer path {'ON' if on else 'OFF'}; arena {a.params.numel()} floats", flush=True)
assert on, "peer path not available"
assert torch.equal(a.params, b.params)
bad = 0
for s in range(6):
    g = torch.Generator(device='cuda').manual_seed(100 * s + rank)
    gr = torch.randn(a.grads.numel(), device='cuda', generator=g) * (10.0 ** (-(s % 3)))
    # reference: the gradient arenas of all ranks added in rank order ((g0 + g1) + g2) + ... with plain torch adds, then adam_kernel
    parts = [torch.empty_like(gr) for _ in range(world)]
    dist.all_gather(parts, gr)
    tot = parts[0].clone()
    for q in range(1, world):
        tot += parts[q]
    a.grads.copy_(tot)
    a.adam_step(grad_scale=1.0 / world)
    b.grads.copy_(gr)
    assert D.allreduce_grads(b.grads) == 1.0 / world      # no-op: the sum happens inside b.adam_step
    b.adam_step(grad_scale=1.0 / world)
    # NCCL's own sum (any order) must agree to rounding
    nc = gr.clone(); D.allreduce_grads(nc)
    torch.cuda.synchronize()
    same = torch.equal(a.params, b.params) and torch.equal(a.exp_avg, b.exp_avg) and torch.equal(a.exp_avg_sq, b.exp_avg_sq) and torch.equal(a.grads, b.grads)
    ref = b.params.clone(); dist.broadcast(ref, 0)
    cross = torch.equal(ref, b.params)
    close = float((nc - b.grads).abs().max()) <= 1e-5 * float(nc.abs().max())
    if world == 2:
        close = close and torch.equal(nc, b.grads)        # two addends: every order gives the same bits
    if not (same and cross and close):
        bad += 1
        print(f"rank {rank} step {s}: rank-order reference bitwise={same} cross-rank bitwise={cross} vs NCCL sum={close} "
              f"max|dp|={float((a.params - b.params).abs().max()):.3e} max|dg|={float((a.grads - b.grads).abs().max()):.3e}", flush=True)
import ctypes as C
from ase_b200 import lib as L
err = C.c_int(0); L.check(L.lib.ase_peer_status(b._peer, C.byref(err), None), 'peer_status')
assert err.value == 0, f"peer error word {err.value}"
# timings: 50 iterations each, CUDA events, after a barrier
def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) * 1e3
def nccl_path():
    a.grads.copy_(gr); sc = D.allreduce_grads(a.grads); a.adam_step(grad_scale=sc)
def peer_path():
    b.grads.copy_(gr); b.adam_step(grad_scale=1.0 / world)
def copy_only():
    a.grads.copy_(gr)
t_n, t_p, t_c = timed(nccl_path), timed(peer_path), timed(copy_only)
L.lib.ase_peer_debug.argtypes = [C.c_void_p, C.c_void_p]
dbg = (C.c_longlong * 8)(); L.lib.ase_peer_debug(b._peer, dbg)
print(f"rank {rank} phase clocks of block 0 (last call): wait-ready {dbg[1]-dbg[0]}, reduce {dbg[2]-dbg[1]}, fence {dbg[3]-dbg[2]}, wait-done {dbg[4]-dbg[3]}, adam {dbg[5]-dbg[4]}", flush=True)
if rank == 0:
    print(f"RESULT world={world} mismatches={bad}  NCCL allreduce + adam_kernel: {t_n - t_c:.1f} us   peer allreduce+Adam kernel: {t_p - t_c:.1f} us   (a 28 MB refill of the gradients, {t_c:.1f} us, subtracted from both)", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if bad else 0)
