"""CPU-side checks of the drop-in boundary: checkpoint layout, seeded init, OCR wrapper API,
C-ABI exports, error conventions (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import ocrl_b200
from ocrl_b200 import abi, synth
from ocrl_b200.config import slate_config, slot_attention_config
from tests.golden_io import load_json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tag,cfg", [("slate", slate_config()), ("bcdec", slot_attention_config(obs_size=32))])
def test_state_dict_layout_matches_reference(tag, cfg):
    ref = load_json(f"state_dict_{tag}.json")
    model = ocrl_b200.SLATE(*cfg)
    mine = {k: list(v.shape) for k, v in model._module.state_dict().items()}
    assert list(mine.keys()) == list(ref["keys"].keys())
    assert mine == ref["keys"]
    assert model.rep_dim == ref["rep_dim"] and model.num_slots == ref["num_slots"]
    assert sum(p.numel() for p in model._module.parameters()) == ref["n_params"]


def test_seeded_init_equals_reference_init():
    """Same seed -> same weights as the reference constructors (parity runs use random init)."""
    sums = load_json("seed0_param_sums.json")
    torch.manual_seed(0)
    model = ocrl_b200.SLATE(*slate_config())
    for k, v in model._module.state_dict().items():
        assert abs(float(v.double().sum()) - sums[k]) <= 1e-9 * max(1.0, abs(sums[k])), k


def test_wrapper_api_surface():
    model = ocrl_b200.SLATE(*slate_config())
    assert model.name == "SLATE"
    for attr in ("__call__", "get_loss", "update", "train", "eval", "to", "set_zero_grad", "do_step", "get_samples",
                 "save", "load", "wandb_watch"):
        assert callable(getattr(model, attr))
    assert isinstance(model._module, torch.nn.Module)
    assert [len(g["params"]) for g in model._opt.param_groups] == [36, 36, 82]
    ckpt = model.save()
    assert set(ckpt) == {"ocr_module_state_dict", "ocr_opt_state_dict"}
    other = ocrl_b200.SLATE(*slate_config())
    other.load(ckpt)
    for (k, a), (_, b) in zip(model._module.state_dict().items(), other._module.state_dict().items()):
        assert torch.equal(a, b), k


def test_use_cnn_feat_shapes():
    model = ocrl_b200.SLATE(*slate_config(use_cnn_feat=True, obs_size=16))
    assert model.num_slots == 256 and model.rep_dim == 67
    out = model(torch.rand(2, 3, 16, 16))  # this branch never runs slot attention (slate_module.py:183-185)
    assert out.shape == (2, 256, 67)


def test_no_cpu_fallback():
    model = ocrl_b200.SLATE(*slate_config(obs_size=16))
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.rand(2, 3, 16, 16))
    sa = ocrl_b200.SlotAttention(3, 6, 64, 192, 192, 2)
    with pytest.raises(NotImplementedError):
        sa(torch.rand(1, 4, 64), torch.rand(1, 6, 192))
    with pytest.raises(AssertionError):
        ocrl_b200.SLATE(*slate_config(obs_size=16))._module(torch.rand(1, 3, 16, 16), True, True)


def test_abi_exports_every_declared_symbol():
    lib = abi.lib()
    header = open(os.path.join(ROOT, "include", "ocrl_sa.h")).read()
    declared = set(re.findall(r"\b(ocrl_[a-z0-9_]+)\s*\(", header))
    assert declared == set(abi.EXPORTS), declared ^ set(abi.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ocrl_version() == 6 and lib.ocrl_built_arch() == b"sm_100a"
    assert ctypes.sizeof(abi.SaDims) == 56 and ctypes.sizeof(abi.SaWeights) == 13 * 8
    assert ctypes.sizeof(abi.TokenWeights) == 10 * 8 and ctypes.sizeof(abi.LaunchOpts) == 24


def test_abi_rejects_bad_dims_without_a_gpu():
    lib = abi.lib()
    f, b, s = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    d = abi.make_dims(2, 100, 64, 192, 192, 6, 3)
    assert lib.ocrl_sa_query_workspace(ctypes.byref(d), ctypes.byref(f), ctypes.byref(b), ctypes.byref(s)) == 0
    assert s.value == 4 * 2 * 3 * ((8 * 6 * 192 + 6 * 192 + 6 + 3) // 4 * 4)
    for bad in (abi.make_dims(2, 100, 64, 192, 192, 17, 3), abi.make_dims(2, 100, 64, 100, 192, 6, 3),
                abi.make_dims(2, 100, 64, 192, 192, 6, 3, heads=2)):
        assert lib.ocrl_sa_query_workspace(ctypes.byref(bad), ctypes.byref(f), ctypes.byref(b), ctypes.byref(s)) == -1
        assert len(lib.ocrl_last_error()) > 0


def test_synthetic_frames_are_deterministic():
    a = synth.random_objs_frames(3, 64, seed=5)
    b = synth.random_objs_frames(3, 64, seed=5)
    assert a.shape == (3, 64, 64, 3) and a.dtype.name == "uint8" and (a == b).all()
    assert (a.reshape(3, -1).max(1) > 0).all()
    assert synth.push_frames(2, 64, seed=1).shape == (2, 64, 64, 3)
    obs = synth.to_obs(torch.from_numpy(a))
    assert obs.shape == (3, 3, 64, 64) and float(obs.max()) <= 1.0


def test_schedules():
    from ocrl_b200.adjacent import cosine_anneal, linear_warmup

    assert cosine_anneal(0, 1.0, 0.1, 0, 100) == pytest.approx(1.0)
    assert cosine_anneal(50, 1.0, 0.1, 0, 100) == pytest.approx(0.55)
    assert cosine_anneal(100, 1.0, 0.1, 0, 100) == 0.1
    assert linear_warmup(0, 0, 1, 0, 10) == pytest.approx(0.1)
    assert linear_warmup(10, 0, 1, 0, 10) == 1


def test_ocrs_registry_shim():
    """The reference's registry call sites (train_ocr.py:37, utils/tools.py:326-335) resolve to the drop-ins when
    ocrl_b200/shim is in front on sys.path: ``getattr(ocrs, cfg.ocr.name)(cfg.ocr, cfg.env)``."""
    import importlib
    import sys

    shim = os.path.join(ROOT, "ocrl_b200", "shim")
    saved = {k: v for k, v in sys.modules.items() if k == "ocrs" or k.startswith("ocrs.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, shim)
    try:
        ocrs = importlib.import_module("ocrs")
        ocr_cfg, env_cfg = slate_config()
        assert ocr_cfg.name == "SLATE"
        model = getattr(ocrs, ocr_cfg.name)(ocr_cfg, env_cfg)  # train_ocr.py:37
        module = getattr(ocrs, ocr_cfg.name + "_Module")(ocr_cfg, env_cfg)  # utils/tools.py:326-330
        assert type(model) is ocrl_b200.SLATE and type(module) is ocrl_b200.SLATE_Module
        assert model.num_slots == 6 and model.rep_dim == 192 and isinstance(model._module, ocrl_b200.SLATE_Module)
        with pytest.raises(AttributeError):
            getattr(ocrs, "NoSuchOCR")
        # sb3s/ocr_extractor.py:33-35: the pooling module by name, built from the OCR's rep_dim / num_slots
        for k in [k for k in sys.modules if k == "poolings" or k.startswith("poolings.")]:
            del sys.modules[k]
        poolings = importlib.import_module("poolings")
        from types import SimpleNamespace as NS

        pcfg = NS(name="Transformer", d_model=128, nhead=8, num_layers=1, pos_emb="None", norm_first=False,
                  use_mlp1=False, use_mlp2=False, cw_embedding=False, push_embedding=False)
        pool = getattr(poolings, pcfg.name + "_Module")(model.rep_dim, model.num_slots, pcfg)
        assert type(pool) is ocrl_b200.Transformer_Module and pool.rep_dim == 128
        for k in [k for k in sys.modules if k == "poolings" or k.startswith("poolings.")]:
            del sys.modules[k]
        from oracle import reference_bridge as rb

        if rb.available():  # with the reference further down the path its sub-packages stay reachable through the shim
            for k in [k for k in sys.modules if k == "ocrs" or k.startswith("ocrs.")]:
                del sys.modules[k]
            for name in ("h5py", "omegaconf"):
                sys.modules.setdefault(name, type(sys)(name))
            sys.path.insert(1, rb.REFERENCE_ROOT)
            try:
                ocrs = importlib.import_module("ocrs")
                assert ocrs.SLATE is ocrl_b200.SLATE
                ref_sa = importlib.import_module("ocrs.common.slot_attn")
                assert ref_sa.__file__.startswith(rb.REFERENCE_ROOT)
            finally:
                sys.path.remove(rb.REFERENCE_ROOT)
    finally:
        sys.path.remove(shim)
        for k in [k for k in sys.modules if k == "ocrs" or k.startswith("ocrs.")]:
            del sys.modules[k]
        sys.modules.update(saved)
