"""Multi-GPU (>= 2 GPUs on the box; skipped otherwise): the fused allreduce + Adam kernel over NVLink peer memory (csrc/peer.cu,
include/ase_b200.h ase_learner_peer_adam_step) against the NCCL allreduce + adam_kernel pair it replaces (Horovod's averaging inside
optimizer.step, learning/amp_agent.py:348-363): bit-identical parameters, moments and summed gradients on every rank and across ranks, over
six steps with gradients of three magnitudes on the full-size 7.04 M-float arena (odd tail)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one node")
def test_peer_allreduce_adam_is_bit_identical_to_nccl_allreduce_plus_adam():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr', '127.0.0.1',
           '--master-port', '29547', os.path.join(root, 'tools', 'peer_adam_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=dict(os.environ, ASE_PEER_ADAM='1'))      # opt in beyond 4 ranks too
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert 'mismatches=0' in out and 'peer path ON' in out, out[-3000:]
