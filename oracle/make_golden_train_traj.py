"""ORACLE — test infrastructure only.  Golden 20-step TRAINING TRAJECTORY at batch 32 (fp32, CPU):
tests/golden/train_traj.npz.

The loop body of code/train_diffusion_paper.py:349-366 (noise, t, perturb_input, train-mode forward, MSE, backward,
Adam with torch defaults) through the oracle restatement (contextunet_oracle.train_step / adam_step, themselves
pinned to the unmodified reference by tests/golden/train_step.npz).  Every draw comes from ONE seeded generator in a
fixed order, so the GPU test regenerates the inputs instead of storing them; the fixture holds the loss curve, the
final values of the parameters nearest the loss and two layers' BatchNorm running statistics.

    python oracle/make_golden_train_traj.py      (about 8 CPU-minutes on 8 cores: the fp32 trajectory and an emulated-bf16 one)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import contextunet_oracle as O  # noqa: E402

SEED, B, STEPS, LR, T, NCF = 4321, 32, 20, 1e-4, 1500, 6
KEEP = ["out.3.weight", "out.3.bias", "out.1.weight", "out.1.bias", "out.0.weight", "out.0.bias",
        "up2.model.2.conv2.1.weight", "up2.model.2.conv2.1.bias", "up2.model.2.conv2.0.weight",
        "contextembed2.model.2.weight", "timeembed2.model.2.bias"]
KEEP_BN = ["up2.model.2.conv2.1", "init_conv.conv1.1", "down2.model.1.conv2.1"]


def draws(seed=SEED, steps=STEPS, batch=B):
    """The trajectory's inputs, in draw order: the fixed batch (maps, params), then per step noise, t, shortcut."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 1, 64, 64, generator=g)
    prm = torch.rand(batch, NCF, generator=g)
    per_step = []
    for _ in range(steps):
        noise = torch.randn(batch, 1, 64, 64, generator=g)
        t = torch.randint(1, T + 1, (batch,), generator=g)
        sc = torch.rand(256, generator=g) * 2 - 1
        per_step.append((noise, t, sc))
    return x, prm, per_step


def run(emulate_bf16=False):
    """-> (losses, final state dict).  emulate_bf16: the same trajectory with bf16 rounding emulated at the device path's
    storage points (pure torch): what ANY bf16 implementation of this step does to the fp32 trajectory."""
    sd = O.init_state_dict(0, n_cfeat=NCF)
    _, _, ab_t = O.make_schedule(T)
    x, prm, per_step = draws()
    names = [k for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k]
    m = {k: torch.zeros_like(sd[k]) for k in names}
    v = {k: torch.zeros_like(sd[k]) for k in names}
    losses = []
    for s, (noise, t, sc) in enumerate(per_step, 1):
        loss, grads, stats = O.train_step(sd, x, prm, t, noise, (sc[:128], sc[128:]), T, ab_t, n_cfeat=NCF,
                                          emulate_bf16=emulate_bf16)
        for k in names:
            sd[k], m[k], v[k] = O.adam_step(sd[k], grads[k], m[k], v[k], s, LR)
        for pre, (mean, uvar) in stats.items():  # nn.BatchNorm2d momentum 0.1, unbiased variance
            sd[pre + ".running_mean"] = 0.9 * sd[pre + ".running_mean"] + 0.1 * mean
            sd[pre + ".running_var"] = 0.9 * sd[pre + ".running_var"] + 0.1 * uvar
        losses.append(float(loss))
        print(f"{'emul ' if emulate_bf16 else ''}step {s:2d}  loss {float(loss):.6f}", flush=True)
    return losses, sd


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    losses, sd = run(False)
    losses_e, sd_e = run(True)
    out = {"seed": SEED, "batch": B, "steps": STEPS, "lr": LR, "losses": np.array(losses, np.float64),
           "emul/losses": np.array(losses_e, np.float64)}
    sd0 = O.init_state_dict(0, n_cfeat=NCF)
    for k in KEEP:
        out["final/" + k] = sd[k].numpy()
        out["delta_norm/" + k] = float((sd[k] - sd0[k]).norm())
        # how far the emulated-bf16 trajectory ends from the fp32 one (relative L2 of the weights)
        out["emul_dev/" + k] = float((sd_e[k] - sd[k]).norm() / sd[k].norm().clamp_min(1e-30))
    for pre in KEEP_BN:
        out["bn/" + pre + ".running_mean"] = sd[pre + ".running_mean"].numpy()
        out["bn/" + pre + ".running_var"] = sd[pre + ".running_var"].numpy()
        v, ve = sd[pre + ".running_var"], sd_e[pre + ".running_var"]
        out["emul_dev/" + pre + ".running_var"] = float((ve - v).norm() / v.norm())
        out["emul_dev/" + pre + ".running_mean"] = float(
            ((sd_e[pre + ".running_mean"] - sd[pre + ".running_mean"]).abs() / v.sqrt()).max())
    for k in sorted(out):
        if k.startswith("emul_dev/"):
            print(k, f"{out[k]:.3e}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "train_traj.npz"), **out)


if __name__ == "__main__":
    main()
