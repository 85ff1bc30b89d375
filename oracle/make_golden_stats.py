"""ORACLE — test infrastructure only (build container only).

Full-length (T = 1500) classifier-free-guided trajectory of the UNMODIFIED reference sampler
(`sample_ddpm`, code/train_diffusion_paper.py:555-623, lifted by ref_harness) on CPU fp32, plus the
reference's own map statistics of the generated maps:

* radial power spectrum: `power_spectrum` imported unchanged from code/diffusion_utilities.py:302-368;
* pixel histograms: the numeric part of `compare_distributions` (code/train_diffusion_paper.py:861-876:
  bins = arange(min, max + 0.01, 0.01), np.histogram(density=True) per image, mean / std over images).

The per-step noise (1500 x B x 4096 floats) is not stored: the GPU test regenerates the reference's whole
draw sequence from the seed (tests/_util.replay_sampler_draws) and checks it against the per-step
checksums stored here before trusting it.

    python oracle/make_golden_stats.py [B] [T]      # rewrites tests/golden/sampler_stats.npz (~12 min at B=8)
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import make_golden as MG  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402

SEED = 50
KEEP_STEPS = (1500, 1400, 1200, 1000, 800, 600, 400, 200, 100, 40, 20, 7, 1)


def reference_histograms(a, b, delta=0.01):
    """compare_distributions (train_diffusion_paper.py:861-876), numeric part only."""
    bin_max = max(a.max(), b.max())
    bin_min = min(a.min(), b.min())
    bins = np.arange(bin_min, bin_max + delta, delta)
    pa = np.array([np.histogram(a[i].ravel(), bins, density=True)[0] for i in range(len(a))])
    pb = np.array([np.histogram(b[i].ravel(), bins, density=True)[0] for i in range(len(b))])
    return bins, pa, pb


def histogram_part(maps, other):
    """An untrained network cannot denoise: x grows by ~1/sqrt(ab_T) ~ 2e3 over the trajectory, so the maps are
    min-max normalised with the REFERENCE run's own extrema (stored) before the reference's 0.01-wide bins are
    applied — the same [0,1] range the reference's normalised training maps live in (train_diffusion_paper.py:260).
    `other` (the step-20 snapshot of the same run) is the second image set of the two-sided call."""
    lo, hi = np.float32(maps.min()), np.float32(maps.max())
    a, b = (maps - lo) / (hi - lo), (other - lo) / (hi - lo)
    bins, pa, pb = reference_histograms(a, b)
    return {"hist_lo": lo, "hist_hi": hi, "hist_other": other, "hist_bins": bins, "hist_a": pa, "hist_b": pb}


def rehist():
    """Recompute only the histogram part of an existing fixture (no 20-minute trajectory)."""
    path = os.path.join(MG.GOLD, "sampler_stats.npz")
    g = dict(np.load(path))
    g.update(histogram_part(g["x"][:, 0], g["hist_other"]))
    np.savez_compressed(path, **g)
    print("rehist:", g["hist_bins"].shape, g["hist_lo"], g["hist_hi"])


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
    torch.set_num_threads(int(os.environ.get("CDM_THREADS", "6")))
    _, raw, cal = MG.build_models()
    _, DU = RH.load_modules()
    ns = RH.paper_namespace(cal, T, MG.NCF)
    torch.manual_seed(SEED)
    params = torch.rand(B, MG.NCF)
    torch.manual_seed(SEED + 1)
    t0 = time.time()
    with RH.DrawRecorder() as rec:
        x, inter, _, _ = ns["sample_ddpm"](n_sample=B, size=64, device=torch.device("cpu"), params=params,
                                           guide_w=2.0)
    print(f"reference trajectory: {time.time() - t0:.1f} s", flush=True)
    snap_steps = [i for i in range(T, 0, -1) if i % 20 == 0 or i == T or i < 8]
    assert len(snap_steps) == inter.shape[0]
    keep = [k for k, i in enumerate(snap_steps) if i in KEEP_STEPS]
    out = {"T": np.int64(T), "B": np.int64(B), "seed": np.int64(SEED + 1), "guide_w": np.float32(2.0),
           "params": params.numpy(), "x_T": rec.randn[0].numpy(), "x": x.numpy(),
           "inter_steps": np.array([snap_steps[k] for k in keep]), "inter": inter[keep]}
    # checksums of every draw, in order, for the replay check
    z = rec.randn[1:]
    assert len(z) == T - 1 and len(rec.shortcuts) == 2 * T
    out["z_sum"] = np.array([float(v.double().sum()) for v in z])
    out["z_first"] = np.stack([v.reshape(-1)[:4].numpy() for v in z])
    out["sc_sum"] = np.array([float(w.double().sum() + 2 * b.double().sum()) for w, b in rec.shortcuts])
    # the reference's statistics of the final maps
    maps = x.numpy()[:, 0]
    pk = []
    for i in range(B):
        k_bins, p = DU.power_spectrum(maps[i], dl=1.0)
        pk.append(p)
    out["pk_k"], out["pk"] = k_bins, np.stack(pk)
    out.update(histogram_part(maps, inter[snap_steps.index(20)][:, 0] if T >= 20 else maps))
    name = "sampler_stats.npz" if T == 1500 else f"sampler_stats_T{T}.npz"
    np.savez_compressed(os.path.join(MG.GOLD, name), **out)
    print(name, {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    assert RH.available(), "reference checkout not found"
    if len(sys.argv) > 1 and sys.argv[1] == "rehist":
        rehist()
    else:
        main()
