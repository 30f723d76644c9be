"""CPU suite, part 3: the N>1 host-side logic (gradient sum-allreduce + 1/world Adam scale, parameter broadcast,
RunningMeanStd buffer averaging) under torch.distributed with the gloo backend, world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


class _Rms:
    def __init__(self, rank):
        self.running_mean = torch.full((5,), float(rank + 1), dtype=torch.float64)
        self.running_var = torch.full((5,), float(2 * rank + 1), dtype=torch.float64)
        self.count = torch.tensor(float(10 * (rank + 1)), dtype=torch.float64)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('dist_utils', os.path.join(root, 'ase_b200', 'dist_utils.py'))
    du = importlib.util.module_from_spec(spec); spec.loader.exec_module(du)
    params = torch.arange(8, dtype=torch.float32) * (1.0 if rank == 0 else -3.0)
    du.broadcast_state([params])
    grads = torch.arange(8, dtype=torch.float32) + 100.0 * rank
    scale = du.allreduce_grads(grads)
    r = _Rms(rank)
    du.sync_running_stats([r])
    q.put((rank, params.tolist(), (grads * scale).tolist(), r.running_mean.tolist(), r.running_var.tolist(), float(r.count)))
    dist.destroy_process_group()


def test_world2_gloo_gradient_average_broadcast_and_stats_sync():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs: p.join(timeout=60)
    expect_grad = [(i + (i + 100.0)) / 2 for i in range(8)]
    for rank, params, g, m, v, c in out:
        assert params == [float(i) for i in range(8)]            # rank 0's parameters everywhere
        assert g == expect_grad                                   # sum-allreduce x 1/world == Horovod average
        assert m == [1.5] * 5 and v == [2.0] * 5 and c == 15.0    # buffers averaged
