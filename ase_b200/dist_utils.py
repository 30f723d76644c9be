"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on the box, gloo in CPU tests).
Mirrors what the reference gets from Horovod through rl_games:
  * broadcast of parameters (+ Adam state) from rank 0 at start       (hvd.setup_algo, common_agent.py:94-95)
  * ONE sum over ranks of the flat gradient arena per minibatch, Adam applies 1/world (amp_agent.py:357-363): on one NVSwitch node the sum and
    Adam are one kernel over NVLink peer memory (init_peer / csrc/peer.cu); otherwise an NCCL allreduce through the library's communicator
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


class _DeviceArray:
    """A raw device allocation seen through __cuda_array_interface__, so that torch can wrap it without copying."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {'shape': (int(n_floats),), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}


_PEER = None          # (AsePeer*, local allocation, wrapped tensor)


def init_peer(learner):
    """Move the learner's gradient arena into NVLink peer memory and attach the fused allreduce + Adam kernel (csrc/peer.cu): every rank
    cudaMallocs its arena through the library, the 64-byte CUDA IPC handles are gathered over torch.distributed, every rank maps the others.
    All ranks take the peer path or none does (e.g. no IPC in the container, no P2P between the GPUs): then the NCCL allreduce stays.
    Returns True when the peer path is on."""
    global _PEER
    if world() == 1 or not learner.params.is_cuda or not (2 <= world() <= 8):
        return False
    import ctypes as C
    import os
    import sys
    want = os.environ.get('ASE_PEER_ADAM', '')
    if want == '0':
        return False
    if world() > 4 and want != '1':
        # measured and bit-checked on 2 and 4 GPUs (profiles/peer_adam_r02.txt); 8 ranks stay on the NCCL allreduce (measured in round 1) until
        # the kernel has been run there: ASE_PEER_ADAM=1 opts in
        return False
    from . import lib as L
    n = learner.grads.numel()
    local, handle = C.c_void_p(), (C.c_uint8 * 64)()
    ok = L.lib.ase_peer_alloc(n, C.byref(local), handle) == 0
    boxes = [None] * world()
    dist.all_gather_object(boxes, bytes(handle) if ok else None)
    ok = all(b is not None for b in boxes)
    peer = C.c_void_p()
    if ok:
        blob = (C.c_uint8 * (64 * world())).from_buffer_copy(b''.join(boxes))
        ok = L.lib.ase_peer_open(blob, world(), dist.get_rank(), local, n, C.byref(peer)) == 0
    flag = torch.tensor([1 if ok else 0], device=learner.params.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        if dist.get_rank() == 0:
            sys.stderr.write(f"ase_b200: NVLink peer-memory optimizer step not available ({L.lib.ase_last_error().decode()}); using the NCCL allreduce\n")
        if peer:
            L.lib.ase_peer_close(peer, 0)
        return False
    pad = (n + 3) // 4 * 4
    arr = _DeviceArray(L.lib.ase_peer_grads(local), pad)
    flat = torch.as_tensor(arr, device=learner.params.device)
    learner.use_grads_arena(flat)
    learner._peer = peer
    _PEER = (peer, local, arr, flat)
    dist.barrier()          # everybody's arena is zeroed and mapped before the first kernel touches a peer
    return True


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def broadcast_state(tensors, src=0):
    for t in tensors:
        dist.broadcast(t, src)


def allreduce_grads(flat_grads):
    """Sum over ranks in place; returns the scale (1/world) the optimizer must apply."""
    w = world()
    if w > 1 and _PEER is not None and flat_grads.data_ptr() == _PEER[3].data_ptr():
        return 1.0 / w          # the sum happens inside the optimizer kernel (Learner.adam_step -> ase_learner_peer_adam_step)
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
