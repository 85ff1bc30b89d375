"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import os

import numpy as np
import torch

from oracle import contextunet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
NCF = 6


def load(name):
    return np.load(os.path.join(GOLD, name))


def T(a):
    return torch.from_numpy(np.asarray(a))


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


_cache = {}


def raw_sd():
    """Seeded random-init weights == torch.manual_seed(0); ContextUnet(1,128,6,64) of the reference."""
    if "raw" not in _cache:
        _cache["raw"] = O.init_state_dict(0, n_cfeat=NCF)
    return {k: v.clone() for k, v in _cache["raw"].items()}


def cal_sd():
    """'Calibrated' synthetic weights: raw init + the norm-layer tensors stored in unet_eval.npz."""
    sd = raw_sd()
    g = load("unet_eval.npz")
    for k in g.files:
        if k.startswith("sd/"):
            sd[k[3:]] = T(g[k]).clone()
    return sd


def split_shortcut(v):
    v = torch.as_tensor(v)
    n = v.numel() // 2
    return v[:n].clone(), v[n:].clone()


def make_model(sd, device="cuda"):
    import camels_diffusion_model_b200 as cdm
    m = cdm.ContextUnet(1, 128, NCF, 64)
    m.load_state_dict(sd)
    return m.to(device).eval()
