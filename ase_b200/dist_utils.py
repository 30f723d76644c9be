"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on the box, gloo in CPU tests).
Mirrors what the reference gets from Horovod through rl_games:
  * broadcast of parameters (+ Adam state) from rank 0 at start       (hvd.setup_algo, common_agent.py:94-95)
  * ONE sum-allreduce of the flat gradient arena per minibatch, Adam applies 1/world (amp_agent.py:357-363)
  * once per epoch: average of the RunningMeanStd buffers              (hvd.sync_stats, common_agent.py:106-107)
Advantage normalisation stays per rank, as in the reference."""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def broadcast_state(tensors, src=0):
    for t in tensors:
        dist.broadcast(t, src)


def allreduce_grads(flat_grads):
    """Sum over ranks in place; returns the scale (1/world) the optimizer must apply."""
    w = world()
    if w > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return 1.0 / w


def sync_running_stats(rms_modules):
    """Average running_mean / running_var / count over ranks (the reference averages the buffers themselves)."""
    w = world()
    if w == 1:
        return
    for r in rms_modules:
        for t in (r.running_mean, r.running_var, r.count):
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.div_(w)
