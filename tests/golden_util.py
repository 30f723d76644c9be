"""Helpers to load tests/golden fixtures (reference outputs) and regenerate their seeded inputs."""
import os
import torch
import ase_oracle as O
import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def sample_idx(numel, k=256):
    g = torch.Generator().manual_seed(numel)
    return torch.randint(0, numel, (min(k, numel),), generator=g)


def calc_grad_case(name):
    """-> (meta, steps, shapes, params P).  Checks the regenerated params against the fixture checksum."""
    fx = load(name)
    meta = fx['meta']
    shapes = (O.ase_param_shapes if meta['kind'] == 'ase' else O.amp_param_shapes)(**meta['shapes_kw'])
    P = synth.params(shapes, seed=meta['seed'])
    for k, v in P.items():
        assert abs(float(v.double().sum()) - meta['param_checksum'][k]) < 1e-9, f"synthetic param drift in {k}"
    return meta, fx['steps'], shapes, P


def rel_err(a, b):
    a = a.double().flatten(); b = b.double().flatten()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
