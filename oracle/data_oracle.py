"""ORACLE — test infrastructure only.

CPU restatement of the reference's data preparation (module-level code of code/train_diffusion_paper.py:232-262),
as functions over numpy / torch CPU tensors.  Pinned by tests/test_oracle_golden.py to tests/golden/data_prep.npz,
which oracle/make_golden_data.py produced by executing those reference statements themselves.
"""
import numpy as np
import torch
import torch.nn.functional as F


def synthetic_maps(seed, n=3, size=256, negative=False):
    """Deterministic stand-in for Maps_HI_IllustrisTNG_LH_z=0.00.npy (float32, log-normal column densities
    spanning ~6 decades); negative=True adds an offset so that min <= 0 (the shift branch, :256-257)."""
    rng = np.random.RandomState(seed)
    m = np.exp(rng.randn(n, size, size).astype(np.float32) * 2.5 + 3.0).astype(np.float32)
    return (m - np.float32(5.0)).astype(np.float32) if negative else m


def synthetic_params(seed, n_sets=7, n_cols=6):
    rng = np.random.RandomState(seed)
    return (rng.rand(n_sets, n_cols) * np.array([0.4, 0.4, 3.75, 3.75, 1.5, 1.5]) + 0.1).astype(np.float32)


def preprocess_maps(camels_data, size=64):
    """:255-262 -> float32 tensor [N,1,size,size]."""
    d = np.asarray(camels_data)
    lo = np.min(d)
    if lo <= 0:
        d = d - lo + 1e-8                      # :256-257
    d = d / np.max(d)                          # :258
    d = np.log10(d)                            # :259
    d = (d - d.min()) / (d.max() - d.min())    # :260
    t = torch.tensor(d, dtype=torch.float32).unsqueeze(1)
    return F.interpolate(t, size=(size, size), mode="bilinear")  # :262


def normalize_params(param_data, num_params, images_per_param=15):
    """:230-252 -> (float32 tensor [n*15, num_params], param_min [1,c], param_max [1,c])."""
    e = np.repeat(np.asarray(param_data), images_per_param, axis=0)
    pmin, pmax = e.min(axis=0, keepdims=True), e.max(axis=0, keepdims=True)
    nrm = (e - pmin) / (pmax - pmin + 1e-8)
    if nrm.shape[1] > num_params:
        nrm = nrm[:, :num_params]
    elif nrm.shape[1] < num_params:
        nrm = np.concatenate([nrm, np.zeros((nrm.shape[0], num_params - nrm.shape[1]))], axis=1)
    return torch.tensor(nrm, dtype=torch.float32), pmin, pmax
