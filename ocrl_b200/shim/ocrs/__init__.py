"""Registry shim: a package importable as ``ocrs`` whose ``SLATE`` / ``SLATE_Module`` are the B200-native drop-ins.

The reference resolves its OCRs by name -- ``getattr(ocrs, config.ocr.name)(config.ocr, config.dataset)``
(train_ocr.py:37, sb3s/ocr_extractor.py:19) and ``getattr(ocrs, name + "_Module")`` (utils/tools.py:326-335).  Put
this directory in front of the reference on ``sys.path`` / ``PYTHONPATH`` and those call sites pick up
``ocrl_b200.SLATE`` without any edit to the reference:

    PYTHONPATH=/path/to/repo/ocrl_b200/shim:/path/to/repo python train_ocr.py ocr=slate ocr.slotattr.num_slots=6

Every other name of the registry (GT, Iodine, VAE, ...) and every sub-package (``ocrs.common`` ...) is served from the
reference's own ``ocrs`` package when one is found further down ``sys.path``: its directory is appended to this
package's ``__path__`` and its classes are looked up lazily, so importing the shim never pulls in dependencies of OCRs
that are not used.
"""
import importlib
import os
import pkgutil
import sys

from ocrl_b200 import SLATE, SLATE_Module  # noqa: F401  (the B200-native slot-attention OCR)

__all__ = ["SLATE", "SLATE_Module"]
_HERE = os.path.dirname(os.path.abspath(__file__))


def _reference_dirs():
    out = []
    for entry in sys.path:
        cand = os.path.join(entry or os.getcwd(), "ocrs")
        if os.path.isfile(os.path.join(cand, "__init__.py")) and os.path.abspath(cand) != _HERE:
            out.append(os.path.abspath(cand))
    return out


for _d in _reference_dirs():
    if _d not in __path__:
        __path__.append(_d)  # ``ocrs.common.slot_attn`` etc. keep resolving to the reference's files


def __getattr__(name):
    """Names this shim does not provide: search the reference's sub-packages (``ocrs/gt``, ``ocrs/vaes`` ...)."""
    if name.startswith("__"):
        raise AttributeError(name)
    for d in __path__[1:]:
        for info in pkgutil.iter_modules([d]):
            if info.name in ("slate", "common"):
                continue
            try:
                mod = importlib.import_module(f"{__name__}.{info.name}")
            except Exception:  # an OCR whose own dependencies are missing (timm, ...) is simply not offered
                continue
            if hasattr(mod, name):
                return getattr(mod, name)
    raise AttributeError(f"module 'ocrs' (ocrl_b200 shim) has no attribute {name!r}")
