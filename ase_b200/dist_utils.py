"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on the box, gloo in CPU tests).
Mirrors what the reference gets from Horovod through rl_games:
  * broadcast of parameters (+ Adam state) from rank 0 at start       (hvd.setup_algo, common_agent.py:94-95)
  * ONE sum-allreduce of the flat gradient arena per minibatch, Adam applies 1/world (amp_agent.py:357-363)
  * once per epoch: average of the RunningMeanStd buffers              (hvd.sync_stats, common_agent.py:106-107)
Advantage normalisation stays per rank, as in the reference."""
import torch
import torch.distributed as dist

_COMM = None          # AseComm* of libase_b200.so (the gradient allreduce goes through the C ABI: include/ase_b200.h ase_grad_allreduce)


def init_comm():
    """Create the library's own NCCL communicator (CUDA ranks only): rank 0 draws the unique id, torch.distributed carries the 128 bytes to the
    other ranks, every rank calls ase_comm_create.  libnccl.so.2 is the one PyTorch already loaded."""
    global _COMM
    if _COMM is not None or world() == 1:
        return _COMM
    import ctypes as C
    import glob
    import os
    from . import lib as L
    os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')       # NCCL's own log lines go to stderr, not into the caller's stdout
    cands = glob.glob(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), 'nvidia', 'nccl', 'lib', 'libnccl.so.2'))
    L.check(L.lib.ase_comm_load(cands[0].encode() if cands else None), 'ase_comm_load')
    buf = (C.c_uint8 * 128)()
    if dist.get_rank() == 0:
        L.check(L.lib.ase_comm_unique_id(buf), 'ase_comm_unique_id')
    box = [bytes(buf)]
    dist.broadcast_object_list(box, src=0)
    buf = (C.c_uint8 * 128).from_buffer_copy(box[0])
    h = C.c_void_p()
    L.check(L.lib.ase_comm_create(buf, dist.get_rank(), dist.get_world_size(), C.byref(h)), 'ase_comm_create')
    _COMM = h
    return _COMM


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def broadcast_state(tensors, src=0):
    for t in tensors:
        dist.broadcast(t, src)


def allreduce_grads(flat_grads):
    """Sum over ranks in place; returns the scale (1/world) the optimizer must apply."""
    w = world()
    if w > 1:
        if flat_grads.is_cuda and _COMM is not None:
            from . import lib as L
            L.check(L.lib.ase_grad_allreduce(_COMM, flat_grads.data_ptr(), flat_grads.numel(), torch.cuda.current_stream().cuda_stream), 'ase_grad_allreduce')
        else:       # gloo (CPU tests of the host logic)
            dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return 1.0 / w


def sync_running_stats(rms_modules):
    """Average running_mean / running_var / count over ranks (the reference averages the buffers themselves)."""
    w = world()
    if w == 1:
        return
    for r in rms_modules:
        for t in (r.running_mean, r.running_var, r.count):
            if t.is_cuda and _COMM is not None:
                from . import lib as L
                L.check(L.lib.ase_comm_allreduce_f64(_COMM, t.data_ptr(), t.numel(), torch.cuda.current_stream().cuda_stream), 'ase_comm_allreduce_f64')
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.div_(w)
