"""ase_b200 -- B200-native engine for the PPO + adversarial update hot path of nv-tlabs/ASE.

The CUDA library (ase_b200/csrc/libase_b200.so, C ABI in include/ase_b200.h) is the product; this package
is the thin host-side mirror of the reference's agent / env helper interfaces.  Importing it without the
built library raises: there is no CPU fallback."""
from . import lib, ops                                   # noqa: F401
from .learner import Learner, param_names                # noqa: F401

__all__ = ['lib', 'ops', 'Learner', 'param_names']
