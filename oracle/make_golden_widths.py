"""ORACLE — test infrastructure only (build container only).

tests/golden/unet_widths.npz: forward vectors of the UNMODIFIED reference `ContextUnet` at the other context widths
of BASELINE config 4 (n_cfeat = 1..5; 6 is in unet_eval.npz), so that the oracle's `init_state_dict` /
`unet_forward` are pinned to the reference there too.  Weights: `torch.manual_seed(5); ContextUnet(1,128,n,64)` with
the seeded norm-affine randomisation of `contextunet_oracle.calibrate_state_dict` (running statistics at their
defaults) — reproducible from the seed, so only a checksum is stored.

    python oracle/make_golden_widths.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import contextunet_oracle as O  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SEED = 5


def main():
    CU, _ = RH.load_modules()
    out = {}
    for n in (1, 2, 3, 4, 5):
        torch.manual_seed(SEED)
        m = CU(1, 128, n, 64)
        out[f"{n}/checksum_init"] = np.float64(O.state_dict_checksum(m.state_dict()))
        m.load_state_dict(O.calibrate_state_dict(m.state_dict()))
        m.eval()
        g = torch.Generator().manual_seed(40 + n)
        x, c = torch.randn(2, 1, 64, 64, generator=g), torch.rand(2, n, generator=g)
        for tn, t in (("t1", torch.tensor([0.25])), ("tB", torch.tensor([0.8, 0.1]))):
            with RH.DrawRecorder() as rec, torch.no_grad():
                y = m(x, t, c)
            out[f"{n}/{tn}/eps"] = y.numpy()
            out[f"{n}/{tn}/shortcut"] = torch.cat(rec.shortcuts[0]).numpy()
            out[f"{n}/{tn}/t"] = t.numpy()
        out[f"{n}/x"], out[f"{n}/c"] = x.numpy(), c.numpy()
        # the oracle must reproduce the reference here, or the fixture is not written
        sd = O.calibrate_state_dict(O.init_state_dict(SEED, n_cfeat=n))
        sc = torch.from_numpy(out[f"{n}/t1/shortcut"])
        with torch.no_grad():
            eps = O.unet_forward(sd, x, torch.tensor([0.25]), c, (sc[:128], sc[128:]), n_cfeat=n)
        err = float((eps - torch.from_numpy(out[f"{n}/t1/eps"])).norm() / torch.from_numpy(out[f"{n}/t1/eps"]).norm())
        assert err < 2e-5, (n, err)
    np.savez_compressed(os.path.join(GOLD, "unet_widths.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("eps")})


if __name__ == "__main__":
    assert RH.available(), "reference checkout not found"
    main()
