"""ORACLE — test infrastructure only (build container only).

tests/golden/driver_contexts.npz: the context tensors (and guidance strengths) the reference's three sampling drivers
hand to `sample_ddpm` — parameter grid (code/train_diffusion_paper.py:917-941), guidance sweep (:1009-1019) and
per-parameter sensitivity (:1114-1127).  They are module-level statements, so they are cut out of the reference
file BY LINE RANGE and exec'd with `sample_ddpm` replaced by a recorder: the text that runs is the reference's,
not a restatement.  One record per context width num_params = 1..6 (BASELINE config 4).

    python oracle/make_golden_drivers.py
"""
import os
import sys
import textwrap

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as RH  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def ref_lines(first, last):
    with open(os.path.join(RH.REF, "code", "train_diffusion_paper.py")) as fh:
        lines = fh.readlines()
    return textwrap.dedent("".join(lines[first - 1:last]))


def base_params(num_params):
    """The `selected_params` stand-in: three seeded rows in [0, 1) (the drivers only read row 0)."""
    return torch.rand(3, num_params, generator=torch.Generator().manual_seed(100 + num_params))


def run_reference(num_params):
    calls = []

    def sample_ddpm(n_sample, size, device, params=None, guide_w=0.0):  # recorder with the closure's signature (:556)
        calls.append((n_sample, params.clone(), float(guide_w)))
        z = torch.zeros(n_sample, 1, size, size)
        return z, None, 0.0, None

    ns = dict(torch=torch, num_params=num_params, selected_params=base_params(num_params), sample_ddpm=sample_ddpm,
              height=64, device="cpu")
    src_grid = ref_lines(917, 941)
    assert src_grid.startswith("if num_params >= 2:") and src_grid.rstrip().endswith("torch.stack(grid_params)")
    exec(src_grid, ns)
    grid = ns["grid_params"]
    src_guid = ref_lines(1009, 1019)
    assert src_guid.startswith("guidance_strengths = [") and src_guid.rstrip().endswith("append(samples_guided)")
    exec(src_guid, ns)
    guid = list(calls)
    del calls[:]
    src_sens = ref_lines(1114, 1127)
    assert src_sens.startswith("for param_idx in range(num_params):") and "sample_ddpm(n_sample=1" in src_sens
    exec(src_sens, ns)
    sens = list(calls)
    return grid, guid, sens


def main():
    out = {}
    for n in range(1, 7):
        grid, guid, sens = run_reference(n)
        out[f"base/{n}"] = base_params(n).numpy()
        out[f"grid/{n}"] = grid.numpy()
        assert all(c[0] == 5 for c in guid) and all(c[0] == 1 for c in sens)
        out[f"guidance_w/{n}"] = np.array([c[2] for c in guid], dtype=np.float64)
        out[f"guidance_params/{n}"] = torch.stack([c[1] for c in guid]).numpy()     # [5 strengths, 5, n]
        out[f"sensitivity/{n}"] = torch.cat([c[1] for c in sens]).numpy()          # [n*5, n], call order
        assert all(c[2] == 0.0 for c in sens)  # the sensitivity calls use sample_ddpm's default guide_w
    np.savez_compressed(os.path.join(GOLD, "driver_contexts.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    assert RH.available(), "reference checkout not found"
    main()
