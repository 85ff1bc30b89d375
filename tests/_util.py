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
PARITY = {}  # measured errors of this session, written to gpurun_out/parity.json at session end (conftest.py)


def record(name, value, tol=None):
    """Keep a measured parity error (and the tolerance it was asserted at) for profiles/rNN_parity.json."""
    PARITY[name] = {"measured": float(value), "tolerance": tol}
    return value


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


def replay_sampler_draws(seed, B, T, reps=2, n_feat=128, size=64):
    """Re-issue, in order, every draw the reference's sample_ddpm (train_diffusion_paper.py:578-609) takes
    from the global CPU generator after torch.manual_seed(seed): x_T, then per step i = T..1 the noise z
    (only if i > 1) followed by one fresh shortcut conv per forward (weight U(-1,1)[n_feat], bias
    U(-1,1)[n_feat]; conditional pass first).  Returns x_T [B,1,s,s], z [T,B,1,s,s] (row k <-> step T-k, last
    row zeros) and the step-indexed shortcut table [T+1,reps,2,n_feat]."""
    torch.manual_seed(seed)
    x_T = torch.randn(B, 1, size, size)
    z = torch.zeros(T, B, 1, size, size)
    tab = torch.zeros(T + 1, reps, 2, n_feat)
    for i in range(T, 0, -1):
        if i > 1:
            z[T - i] = torch.randn(B, 1, size, size)
        for r in range(reps):
            tab[i, r, 0].uniform_(-1, 1)
            tab[i, r, 1].uniform_(-1, 1)
    return x_T, z, tab


def check_replay(g, x_T, z, tab):
    """Hold a replayed draw sequence to the per-draw checksums stored with the golden trajectory."""
    T = int(g["T"])
    assert torch.equal(x_T, T_(g["x_T"]))
    zs = z[:T - 1].double().sum(dim=(1, 2, 3, 4)).numpy()
    np.testing.assert_allclose(zs, g["z_sum"], rtol=0, atol=1e-9)
    assert np.array_equal(z[:T - 1].reshape(T - 1, -1)[:, :4].numpy(), g["z_first"])
    sc = np.array([float(tab[i, r, 0].double().sum() + 2 * tab[i, r, 1].double().sum())
                   for i in range(T, 0, -1) for r in range(tab.shape[1])])
    np.testing.assert_allclose(sc, g["sc_sum"], rtol=0, atol=1e-9)


T_ = T
