"""Map statistics of generated samples on the device: radial power spectrum and pixel histograms.

Mirrors `power_spectrum` / `compare_power_spectra` (code/diffusion_utilities.py:302-431) and the numeric
part of `compare_distributions` (code/train_diffusion_paper.py:861-876) — same names, argument meaning
and return values (numpy arrays), but the maps stay in HBM: one kernel launch per batch instead of a
Python loop over every Fourier mode of every map.  Plotting is not reproduced (the reference's
`plt.savefig` calls); the functions return the numbers the plots are drawn from.
"""
import functools

import numpy as np
import torch

from . import _lib as L


@functools.lru_cache(maxsize=8)
def _radial_bins(n, dl):
    """The reference's binning rule (diffusion_utilities.py:325-356) evaluated once per (N, dl):
    k = 2 pi fftfreq(N, dl) on both axes, dk = 2 pi/(N dl), bin = int(round(|k|/dk)) (Python's round),
    n_bins = ceil(max|k|/dk) + 1.  Returns (k_bins fp64 [n_bins], CSR start int32, CSR items int32)."""
    kc = 2 * np.pi * np.fft.fftfreq(n, dl)
    kx, ky = np.meshgrid(kc, kc, indexing="ij")
    kgrid = np.sqrt(kx ** 2 + ky ** 2)
    dk = 2 * np.pi / (n * dl)
    n_bins = int(np.ceil(np.max(kgrid) / dk)) + 1
    flat = kgrid.flatten()
    idx = np.array([int(round(v / dk)) for v in flat], dtype=np.int64)
    assert idx.max() < n_bins
    order = np.argsort(idx, kind="stable").astype(np.int32)
    start = np.zeros(n_bins + 1, np.int32)
    np.cumsum(np.bincount(idx, minlength=n_bins), out=start[1:])
    return np.arange(n_bins) * dk, start, order


def _as_maps(x, dev):
    """[B,1,H,W] / [B,H,W] / [H,W] numpy or tensor -> contiguous fp32 device tensor [B,H,W]."""
    t = torch.as_tensor(x)
    if t.dim() == 4:
        t = t[:, 0]
    if t.dim() == 2:
        t = t[None]
    return t.to(dev, torch.float32).contiguous()


def power_spectra(maps, dl=1.0, device=None):
    """Batched form: maps [B,(1,)N,N] -> (k_bins [n_bins], pk [B, n_bins]) as float64 numpy arrays."""
    dev = torch.device("cuda") if device is None else torch.device(device)
    m = _as_maps(maps, dev)
    if m.shape[1] != m.shape[2]:
        raise L.CdmError("power_spectra: square maps only")
    k_bins, start, items = _radial_bins(int(m.shape[1]), float(dl))
    pk = torch.empty(m.shape[0], len(k_bins), device=dev, dtype=torch.float64)
    L.power_spectrum(m, torch.from_numpy(start).to(dev), torch.from_numpy(items).to(dev), float(dl) ** 2, pk)
    return k_bins, pk.cpu().numpy()


def power_spectrum(box, dl=1.0):
    """Drop-in for diffusion_utilities.power_spectrum on one 2-D map -> (k_bins, pk)."""
    if np.ndim(box) != 2:
        raise ValueError("Input box must be 2D")  # the 3-D branch of the reference is never reached by its callers
    k, pk = power_spectra(np.asarray(box)[None], dl)
    return k, pk[0]


def compare_power_spectra(original_images, generated_images, output_dir=None, dl=1.0, title=None):
    """diffusion_utilities.py:370-431 without the plot -> (k, orig_pk_mean, gen_pk_mean)."""
    n = min(len(original_images), len(generated_images))
    k, po = power_spectra(original_images[:n], dl)
    _, pg = power_spectra(generated_images[:n], dl)
    return k, po.mean(0), pg.mean(0)


def pixel_histograms(images, bins, density=True, device=None):
    """np.histogram(images[i].ravel(), bins, density=density)[0] for every i -> float64 [B, len(bins)-1]."""
    dev = torch.device("cuda") if device is None else torch.device(device)
    m = _as_maps(images, dev)
    m = m.reshape(m.shape[0], -1)
    edges = np.asarray(bins, dtype=np.float64)
    counts = torch.empty(m.shape[0], len(edges) - 1, device=dev, dtype=torch.int32)
    L.pixel_histogram(m, torch.from_numpy(edges).to(dev), counts)
    c = counts.cpu().numpy().astype(np.float64)
    if not density:
        return c
    return c / np.diff(edges)[None, :] / c.sum(1, keepdims=True)


def compare_distributions(camels_images, diffusion_images, output_dir=None, bin_delta=0.01):
    """train_diffusion_paper.py:861-876 without the plot -> dict(bins, train_pdf_mean, train_pdf_std,
    test_pdf_mean, test_pdf_std): common bins arange(min, max + 0.01, 0.01), density histogram per image,
    mean / std over the first len(camels_images) images of each set."""
    a, b = torch.as_tensor(camels_images), torch.as_tensor(diffusion_images)
    bin_max = max(float(a.max()), float(b.max()))
    bin_min = min(float(a.min()), float(b.min()))
    # the reference's scalars are np.float32 (numpy .max()/.min() of float32 arrays): keep that rounding
    bins = np.arange(np.float32(bin_min), np.float32(bin_max) + bin_delta, bin_delta)
    n = len(a)
    pa, pb = pixel_histograms(a[:n], bins), pixel_histograms(b[:n], bins)
    return {"bins": bins, "train_pdf_mean": pa.mean(0), "train_pdf_std": pa.std(0),
            "test_pdf_mean": pb.mean(0), "test_pdf_std": pb.std(0)}
