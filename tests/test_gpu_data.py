"""Device data preparation against the reference's statements (tests/golden/data_prep.npz)."""
import numpy as np
import pytest
import torch

from oracle import data_oracle as DO
from tests._util import load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,neg", [("pos", False), ("neg", True)])
def test_preprocess_maps(tag, neg):
    from camels_diffusion_model_b200 import data
    g = load("data_prep.npz")
    maps = DO.synthetic_maps(11, n=30, size=256, negative=neg)
    out = data.preprocess_maps(maps)
    assert out.shape == (30, 1, 64, 64) and out.dtype == torch.float32
    got, ref = out[:4].cpu().numpy(), g[f"{tag}/maps"]
    # fp32 on both sides; CUDA log10f vs numpy's log10 differ by <= 1 ulp before the min-max division
    assert np.abs(got - ref).max() < 2e-6, np.abs(got - ref).max()
    assert out.min() >= 0 and out.max() <= 1
    # full set against the oracle restatement (pinned to the same reference statements)
    assert (out.cpu() - DO.preprocess_maps(maps)).abs().max() < 2e-6


def test_preprocess_maps_other_sizes():
    from camels_diffusion_model_b200 import data
    maps = DO.synthetic_maps(5, n=3, size=96)
    for size in (64, 32, 96, 128):  # down-, same- and up-sampling taps of F.interpolate(bilinear)
        assert (data.preprocess_maps(maps, size=size).cpu() - DO.preprocess_maps(maps, size=size)).abs().max() < 2e-6


@pytest.mark.parametrize("k", [6, 2, 8])
def test_normalize_params(k):
    from camels_diffusion_model_b200 import data
    g = load("data_prep.npz")
    params = DO.synthetic_params(12, n_sets=2)
    tab, pmin, pmax = data.normalize_params(params, k)
    assert np.array_equal(pmin, g["pos/pmin"]) and np.array_equal(pmax, g["pos/pmax"])
    # the reference divides in float32 numpy, as the kernel does: bit-exact
    assert np.array_equal(tab.cpu().numpy(), g[f"pos/params{k}"])


def test_minmax_large_and_ragged():
    from camels_diffusion_model_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(0)
    for n in (1, 3, 1000, 4 * 1024 * 1024 + 3):
        x = torch.randn(n, device="cuda", generator=g)
        ws, out = torch.zeros(2 * 148 * 8 + 1, device="cuda"), torch.empty(2, device="cuda")
        for _ in range(2):  # the ticket re-arms itself
            L.minmax(x, ws, out)
            assert out[0] == x.min() and out[1] == x.max()
