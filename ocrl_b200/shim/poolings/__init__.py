"""Registry shim: a package importable as ``poolings`` whose ``Transformer_Module`` is the B200-native drop-in.

``sb3s/ocr_extractor.py:33-35`` builds the pooling with ``getattr(poolings, config.pooling.name + "_Module")(rep_dim,
num_slots, config.pooling)``; with this directory in front of the reference on ``sys.path`` that call returns
``ocrl_b200.pooling.Transformer_Module`` (same constructor, sub-module names and ``state_dict``; one fused kernel on
the rollout path, torch ops when PPO needs gradients).  Every other name (``Transformer`` wrapper, ``MLP``, ``RN`` ...)
and every sub-package is served from the reference's own ``poolings`` package found further down ``sys.path``.
"""
import importlib
import os
import pkgutil
import sys

from ocrl_b200.pooling import Transformer_Module  # noqa: F401

__all__ = ["Transformer_Module"]
_HERE = os.path.dirname(os.path.abspath(__file__))

for _entry in sys.path:
    _cand = os.path.join(_entry or os.getcwd(), "poolings")
    if os.path.isfile(os.path.join(_cand, "__init__.py")) and os.path.abspath(_cand) != _HERE and os.path.abspath(_cand) not in __path__:
        __path__.append(os.path.abspath(_cand))  # ``poolings.common.transformer`` etc. keep resolving to the reference's files


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    for d in __path__[1:]:
        for info in pkgutil.iter_modules([d]):
            if info.name == "common":
                continue
            try:
                mod = importlib.import_module(f"{__name__}.{info.name}")
            except Exception:
                continue
            if hasattr(mod, name):
                return getattr(mod, name)
    raise AttributeError(f"module 'poolings' (ocrl_b200 shim) has no attribute {name!r}")
