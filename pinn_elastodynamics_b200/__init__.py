"""pinn_elastodynamics_b200 -- B200-native residual/training engine for PINN elastodynamics.

Drop-in for the hot path of Raocp/PINN-elastodynamics: the model classes below mirror the reference's
`PINN` / `DeepHPM` / `DeepElasticWave` surface (`PhysicsInformedNN` is the alias BASELINE.json uses); their
TF1 graph is replaced by hand-written sm_100a CUDA kernels behind the C ABI in include/pinn_elasto.h.
Importing this package does not touch the GPU; constructing a model requires one (no CPU fallback).
"""
from . import _lib
from ._lib import PeError

__all__ = ['PINN', 'PhysicsInformedNN', 'DeepHPM', 'DeepElasticWave', 'Network', 'LossEngine', 'PeError', 'preprocess']


def __getattr__(name):
    if name in ('PINN', 'PhysicsInformedNN', 'DeepHPM', 'DeepElasticWave'):
        from . import models
        return getattr(models, name)
    if name in ('Network', 'LossEngine'):
        from . import engine
        return getattr(engine, name)
    if name == 'preprocess':
        import importlib
        return importlib.import_module('.preprocess', __name__)
    raise AttributeError(name)
