"""Import the real reference (read-only at /root/reference) in the BUILD CONTAINER only.

TEST INFRASTRUCTURE.  The reference is Python and cannot travel to the GPU box, so this module
is used only by ``oracle/make_golden.py`` (to freeze fixtures) and by CPU tests that skip when
the tree is absent.  Nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports it.

``import ocrs`` fails here (h5py / omegaconf / timm are not installed), so three stubs are put
in ``sys.modules`` before importing the leaf modules.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace as NS

REFERENCE_ROOT = os.environ.get("OCRL_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ocrs", "common", "slot_attn.py"))


_loaded = False


def load():
    """Returns a namespace with the reference classes on the hot path."""
    global _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if not _loaded:
        for name in ("h5py", "omegaconf"):  # utils/tools.py:8,10 import them at module scope
            sys.modules.setdefault(name, types.ModuleType(name))
        pkg = types.ModuleType("ocrs")  # skip ocrs/__init__.py (timm via ocrs/mae)
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "ocrs")]
        sys.modules["ocrs"] = pkg
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        _loaded = True
    from ocrs.common.slot_attn import SlotAttention, SlotAttentionEncoder
    from ocrs.common.models import SlotAttnCNNEncoder, BroadCastDecoder, dVAE
    from ocrs.common.utils import PositionalEmbedding
    from ocrs.slate.slate_module import SLATE_Module
    from ocrs.slate.slate import SLATE

    return NS(SlotAttention=SlotAttention, SlotAttentionEncoder=SlotAttentionEncoder,
              SlotAttnCNNEncoder=SlotAttnCNNEncoder, PositionalEmbedding=PositionalEmbedding,
              BroadCastDecoder=BroadCastDecoder, dVAE=dVAE, SLATE_Module=SLATE_Module, SLATE=SLATE)


def _ns(d):
    return NS(**{k: _ns(v) if isinstance(v, dict) else v for k, v in d.items()})


def slate_config(num_slots=6, num_iterations=3, slot_size=192, mlp_hidden_size=192,
                 use_bcdec=False, obs_size=64):
    """configs/ocr/slate.yaml:1-35 with the README's num_slots=6 override, as attribute objects."""
    ocr = _ns(dict(
        name="SLATE", tau_start=1.0, tau_final=0.1, tau_steps=30000, hard=False,
        use_cnn_feat=False, use_bcdec=use_bcdec,
        dvae=dict(vocab_size=4096, d_model=192), cnn=dict(hidden_size=64),
        slotattr=dict(num_iterations=num_iterations, num_slots=num_slots, num_slot_heads=1,
                      slot_size=slot_size, mlp_hidden_size=mlp_hidden_size, pos_channels=4),
        tfdec=dict(num_dec_blocks=4, num_dec_heads=4),
        learning=dict(lr_half_life=250000, lr_dvae=3e-4, lr_enc=1e-4, lr_dec=3e-4,
                      lr_warmup_steps=30000, dropout=0.1, clip=0.05)))
    env = _ns(dict(obs_size=obs_size, obs_channels=3))
    return ocr, env
