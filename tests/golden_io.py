"""Load the frozen reference outputs under tests/golden (made by oracle/make_golden.py)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name, device="cpu", dtype=None):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    groups = {"p": {}, "in": {}, "out": {}, "g.p": {}, "g.in": {}, "g.out": {}}
    for key in z.files:
        if key == "meta":
            continue
        t = torch.from_numpy(z[key])
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        t = t.to(device)
        for g in ("g.p.", "g.in.", "g.out.", "p.", "in.", "out."):
            if key.startswith(g):
                groups[g[:-1]][key[len(g):]] = t
                break
    return meta, groups


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def rel_err(a, b):
    """relative L2 error of a against reference b"""
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
