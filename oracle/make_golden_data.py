"""ORACLE — test infrastructure only (build container only).

tests/golden/data_prep.npz: outputs of the reference's data-preparation statements.  They are module-level code
(code/train_diffusion_paper.py:230-262 sits between an np.load of a dataset that is not in the repo and the
DataLoader construction), so the statements are cut out of the reference file BY LINE RANGE and exec'd on
synthetic arrays — the text that runs is the reference's, not a restatement.

    python oracle/make_golden_data.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import data_oracle as DO  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def ref_lines(first, last):
    with open(os.path.join(RH.REF, "code", "train_diffusion_paper.py")) as fh:
        lines = fh.readlines()
    return "".join(lines[first - 1:last])


def run_reference(camels_data, param_data, num_params):
    src_params = ref_lines(232, 252)   # expanded_param_data = np.repeat(...) ... param_data_tensor = torch.tensor(...)
    src_maps = ref_lines(255, 262)     # min_value = np.min(camels_data) ... camels_data_resized = F.interpolate(...)
    assert src_params.lstrip().startswith("expanded_param_data") and "param_data_tensor" in src_params
    assert src_maps.startswith("min_value") and "camels_data_resized" in src_maps
    # the np.save of param_min/param_max and the shape assert need an output_dir / matching counts: provide them
    import tempfile
    ns = dict(np=np, torch=torch, F=F, os=os, camels_data=camels_data, param_data=param_data, num_params=num_params,
              output_dir=tempfile.mkdtemp())
    exec(src_params, ns)
    exec(src_maps, ns)
    return ns["camels_data_resized"], ns["param_data_tensor"], ns["param_min"], ns["param_max"]


def main():
    out = {}
    for tag, neg in (("pos", False), ("neg", True)):
        n_sets = 2
        maps = DO.synthetic_maps(11, n=15 * n_sets, size=256, negative=neg)[:, ::1]
        params = DO.synthetic_params(12, n_sets=n_sets)
        for num_params in (6, 2, 8) if tag == "pos" else (6,):
            resized, ptab, pmin, pmax = run_reference(maps.copy(), params.copy(), num_params)
            out[f"{tag}/params{num_params}"] = ptab.numpy()
        out[f"{tag}/maps"] = resized.numpy()[:4]  # the first 4 maps (the extrema are over all 30)
        out[f"{tag}/pmin"], out[f"{tag}/pmax"] = pmin, pmax
        # the oracle restatement must reproduce the reference statements exactly
        assert torch.equal(DO.preprocess_maps(maps), resized)
        assert torch.equal(DO.normalize_params(params, 6)[0], torch.tensor(out[f"{tag}/params6"]))
    np.savez_compressed(os.path.join(GOLD, "data_prep.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    assert RH.available(), "reference checkout not found"
    main()
