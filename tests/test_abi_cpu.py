"""CPU suite, part 2: the C-ABI shared library loads and exports every symbol include/ase_b200.h declares.
No compute calls (no GPU here); pure host-side queries only."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def aselib():
    from ase_b200 import build
    build.build()
    from ase_b200 import lib
    return lib


def _declared():
    src = open(os.path.join(ROOT, 'include', 'ase_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ase_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported(aselib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(aselib.lib, n), f"{n} declared in include/ase_b200.h but not exported"
    assert sorted(aselib.EXPORTS) == names


def test_abi_version_and_layout_queries(aselib):
    L = aselib
    assert L.lib.ase_abi_version() == 4
    cfg = L.LearnerConfig()
    cfg.kind = L.KIND_ASE
    cfg.obs_dim, cfg.act_dim, cfg.amp_dim, cfg.latent_dim = 253, 31, 1400, 64
    cfg.n_units = 3; cfg.units[0], cfg.units[1], cfg.units[2] = 1024, 1024, 512
    cfg.n_disc_units = 3; cfg.disc_units[0], cfg.disc_units[1], cfg.disc_units[2] = 1024, 1024, 512
    cfg.n_style_units = 2; cfg.style_units[0], cfg.style_units[1] = 512, 256
    cfg.batch, cfg.amp_batch = 16384, 4096
    assert L.lib.ase_learner_num_params(C.byref(cfg)) == 32            # 32 trainable tensors
    total = 0
    for i in range(32):
        off, r, c = C.c_int64(), C.c_int(), C.c_int()
        assert L.lib.ase_learner_param_desc(C.byref(cfg), i, C.byref(off), C.byref(r), C.byref(c)) == 0
        assert off.value % 32 == 0
        total += r.value * c.value
    assert total == 7039905 - 0                                         # 7,039,905 trainable parameters
    assert L.lib.ase_learner_arena_floats(C.byref(cfg)) >= total
    assert L.lib.ase_learner_workspace_bytes(C.byref(cfg)) > 0
    bad = L.LearnerConfig()
    assert L.lib.ase_learner_num_params(C.byref(bad)) < 0
    assert b'learner' in L.lib.ase_last_error()


def test_param_names_match_reference_state_dict():
    from ase_b200.learner import param_names
    import ase_oracle as O
    assert ['sigma'] + param_names('ase', 3, 3, 2) == list(O.ase_param_shapes().keys())
    assert ['sigma'] + param_names('amp', 2, 2, 0) == list(O.amp_param_shapes().keys())
    assert ['sigma'] + param_names('ppo', 2, 0, 0) == list(O.amp_param_shapes(amp=0).keys())


def test_async_epoch_log_ring_cpu():
    """AsyncEpochLog on CPU tensors: order, means, ring overflow drains the oldest epoch first."""
    import torch
    from ase_b200.async_log import AsyncEpochLog
    log = AsyncEpochLog(['a', 'b'], depth=2)
    got = []
    for e in range(1, 6):
        series = torch.tensor([[float(e), 2.0 * e], [float(e) + 2, 2.0 * e]])
        got += log.push(e, series, frames=10 * e)
        if e == 3:
            got += log.poll()
    got += log.flush()
    assert [r['epoch'] for r in got] == [1, 2, 3, 4, 5]
    assert [r['frames'] for r in got] == [10, 20, 30, 40, 50]
    assert all(abs(r['scalars']['a'] - (r['epoch'] + 1)) < 1e-6 and abs(r['scalars']['b'] - 2 * r['epoch']) < 1e-6 for r in got)
    assert log.poll() == [] and log.flush() == []


def test_ctypes_signatures_have_the_declared_arity(aselib):
    """Every prototype of include/ase_b200.h against the ctypes binding (ase_b200/lib.py): same number of parameters wherever argtypes are
    set -- a binding that drifted from the header would otherwise corrupt the call silently."""
    src = open(os.path.join(ROOT, 'include', 'ase_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    src = re.sub(r'//[^\n]*', '', src)
    protos = re.findall(r'\b(ase_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;', src, flags=re.S)
    assert len(protos) >= 40
    checked = 0
    for name, args in protos:
        args = ' '.join(args.split())
        n = 0 if args in ('', 'void') else args.count(',') + 1
        fn = getattr(aselib.lib, name)
        if fn.argtypes is not None:
            assert len(fn.argtypes) == n, f"{name}: header has {n} parameters, ctypes binding {len(fn.argtypes)}"
            checked += 1
    assert checked >= 30
