"""Tiny stand-in for the Hydra config objects the reference passes around.  Anything with
attribute access works (omegaconf DictConfig, SimpleNamespace); hydra/omegaconf are not required.
Values mirror configs/ocr/slate.yaml:1-35 and the README's num_slots=6 override."""
from types import SimpleNamespace

import yaml


def to_namespace(d):
    if isinstance(d, dict):
        return SimpleNamespace(**{k: to_namespace(v) for k, v in d.items()})
    return d


def load_yaml(path, **overrides):
    """Load one of the reference's YAML files (no Hydra composition) and apply dotted overrides,
    e.g. ``load_yaml("configs/ocr/slate.yaml", **{"slotattr.num_slots": 6})``."""
    with open(path) as f:
        cfg = yaml.safe_load(f)
    for key, val in overrides.items():
        node = cfg
        parts = key.split(".")
        for part in parts[:-1]:
            node = node[part]
        node[parts[-1]] = val
    return to_namespace(cfg)


def slate_config(num_slots=6, num_iterations=3, slot_size=192, mlp_hidden_size=192, use_bcdec=False,
                 use_cnn_feat=False, obs_size=64, obs_channels=3, kv_dtype=None):
    ocr = to_namespace(dict(
        name="SLATE", tau_start=1.0, tau_final=0.1, tau_steps=30000, hard=False, use_cnn_feat=use_cnn_feat,
        use_bcdec=use_bcdec, dvae=dict(vocab_size=4096, d_model=192), cnn=dict(hidden_size=64),
        slotattr=dict(num_iterations=num_iterations, num_slots=num_slots, num_slot_heads=1, slot_size=slot_size,
                      mlp_hidden_size=mlp_hidden_size, pos_channels=4,
                      **({} if kv_dtype is None else {"kv_dtype": kv_dtype})),
        tfdec=dict(num_dec_blocks=4, num_dec_heads=4),
        learning=dict(lr_half_life=250000, lr_dvae=3e-4, lr_enc=1e-4, lr_dec=3e-4, lr_warmup_steps=30000,
                      dropout=0.1, clip=0.05)))
    env = to_namespace(dict(obs_size=obs_size, obs_channels=obs_channels))
    return ocr, env


def slot_attention_config(large=False, **kw):
    """The paper's "Slot-Attention" rows = SLATE with the broadcast decoder (for_running.json:55-81)."""
    if large:
        return slate_config(num_iterations=7, slot_size=192, mlp_hidden_size=192, use_bcdec=True, **kw)
    return slate_config(num_iterations=7, slot_size=64, mlp_hidden_size=128, use_bcdec=True, **kw)
