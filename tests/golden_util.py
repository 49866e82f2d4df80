"""Helpers to read the committed reference fixtures (tests/golden/*.npz)."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    # adapter cases only (f3_*.npz are the fixtures of tests/test_f3_heads.py)
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")) if not os.path.basename(p).startswith("f3_"))


def load_case(name):
    from oracle import dgsct_oracle as O
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    cfg = O.AdapterConfig(**meta["cfg"])
    t = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    params = {k[len("param/"):]: v for k, v in t.items() if k.startswith("param/")}
    return cfg, meta, params, t


def rel_l2(a, b):
    a = a.detach().double().reshape(-1)
    b = b.detach().double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))
