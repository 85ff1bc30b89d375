"""ORACLE — test infrastructure only.

CPU restatement of the reference's map statistics, pinned by tests/test_oracle_golden.py to the values the
UNMODIFIED reference functions produced (tests/golden/sampler_stats.npz, written by oracle/make_golden_stats.py):

* power_spectrum   — code/diffusion_utilities.py:302-368 (2-D branch), vectorised instead of the per-mode loop;
* histograms       — the numeric part of compare_distributions, code/train_diffusion_paper.py:861-876.
"""
import numpy as np


def power_spectrum(box, dl=1.0):
    box = np.asarray(box)
    n0, n1 = box.shape
    ft = np.fft.fftn(box, norm="ortho")                                   # :321
    kx, ky = np.meshgrid(2 * np.pi * np.fft.fftfreq(n0, dl), 2 * np.pi * np.fft.fftfreq(n1, dl), indexing="ij")
    kgrid = np.sqrt(kx ** 2 + ky ** 2)                                    # :329-332
    dk = 2 * np.pi / (min(n0, n1) * dl)                                   # :339
    n_bins = int(np.ceil(np.max(kgrid) / dk)) + 1                         # :340-341
    idx = np.rint(kgrid.flatten() / dk).astype(np.int64)                  # :353 int(round(.)): half-to-even, as rint
    power = (np.abs(ft) ** 2).flatten()
    ok = idx < n_bins
    pk = np.bincount(idx[ok], weights=power[ok].astype(np.float64), minlength=n_bins)
    count = np.bincount(idx[ok], minlength=n_bins)
    pk[count > 0] /= count[count > 0]                                     # :358-360
    return np.arange(n_bins) * dk, pk * dl ** 2                           # :363-366


def histograms(a, b, delta=0.01):
    """-> (bins, per-image density histograms of a, of b)  (train_diffusion_paper.py:862-873)."""
    bin_max = max(a.max(), b.max())
    bin_min = min(a.min(), b.min())
    bins = np.arange(bin_min, bin_max + delta, delta)
    pa = np.array([np.histogram(a[i].ravel(), bins, density=True)[0] for i in range(len(a))])
    pb = np.array([np.histogram(b[i].ravel(), bins, density=True)[0] for i in range(len(b))])
    return bins, pa, pb
