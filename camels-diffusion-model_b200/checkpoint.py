"""Checkpoint I/O.

* `.pth` compatibility: the reference saves `torch.save(nn_model.state_dict(), save_dir + f"model_epoch_{ep}.pth")`
  (code/train_diffusion_paper.py:477-478) and loads it with `load_state_dict(torch.load(path, map_location=device))`
  (code/sample_power_spectra.py:187-189).  ContextUnet here keeps all 156 state_dict keys with PyTorch-native shapes,
  so `save_model` / `load_model` files are interchangeable with the reference's in both directions.
* resume: the reference cannot resume (weights only).  `save_checkpoint` additionally stores the optimizer state
  (torch.optim.Adam layout), the epoch / step counters and the CPU generator state that drives the per-forward
  shortcut draws and the `t ~ randint` draws, so a resumed run continues bit-for-bit.
* data parallel: rank 0 writes; every rank loads the same file (`load_checkpoint`), or rank 0 loads and
  `broadcast_model` ships parameters and BatchNorm buffers to the other ranks.
"""
import os

import torch
import torch.distributed as dist

from .parallel import world


def save_model(model, path):
    """The reference's weights-only `.pth` (rank 0 only under torch.distributed)."""
    rank, _ = world()
    if rank == 0:
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        tmp = path + ".tmp"
        torch.save(sd, tmp)
        os.replace(tmp, path)  # never leave a truncated checkpoint behind


def load_model(model, path, map_location=None):
    sd = torch.load(path, map_location=map_location or "cpu")
    if isinstance(sd, dict) and "model" in sd and "format" in sd:  # a resume checkpoint: take its weights
        sd = sd["model"]
    model.load_state_dict(sd)
    return model


def save_checkpoint(path, model, optim, epoch, step=0, extra=None):
    """Full training state -> one file.  `optim`: FusedAdam, torch.optim.Adam or GraphedTrainStep."""
    rank, _ = world()
    if rank != 0:
        return
    ck = {"format": "cdm_b200/1", "epoch": int(epoch), "step": int(step),
          "model": {k: v.detach().cpu() for k, v in model.state_dict().items()},
          "optim": _to_cpu(optim.state_dict()), "cpu_rng_state": torch.get_rng_state(), "extra": _plain(extra or {})}
    tmp = path + ".tmp"
    torch.save(ck, tmp)
    os.replace(tmp, path)


def load_checkpoint(path, model, optim=None, restore_rng=True):
    """-> (epoch, step, extra).  Restores weights + BatchNorm buffers, optimizer moments / step count and (by default)
    the CPU generator state."""
    ck = torch.load(path, map_location="cpu", weights_only=True)  # tensors and plain containers only: no pickle code
    if not (isinstance(ck, dict) and ck.get("format", "").startswith("cdm_b200/")):
        raise ValueError(f"{path} is not a resume checkpoint (weights-only .pth files go through load_model)")
    model.load_state_dict(ck["model"])
    if optim is not None:
        optim.load_state_dict(ck["optim"])
    if restore_rng:
        torch.set_rng_state(ck["cpu_rng_state"])
    return ck["epoch"], ck["step"], ck["extra"]


def broadcast_model(model, src=0):
    """Parameters and buffers of rank `src` -> every rank (after rank 0 alone loaded a checkpoint)."""
    _, ws = world()
    if ws == 1:
        return
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=src)
    if hasattr(model, "invalidate"):
        model.invalidate()  # .data writes do not bump version counters: re-pack the eval weights on next use


def _plain(o):
    """`extra` as tensors + plain Python containers / scalars only (numpy scalars and arrays become floats / lists), so
    that the file loads with torch.load(weights_only=True): a resume file never needs arbitrary pickle code."""
    import numpy as np
    if torch.is_tensor(o):
        return o.detach().cpu()
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, np.generic):
        return o.item()
    if isinstance(o, dict):
        return {(k if isinstance(k, (str, int, float, bool)) else str(k)): _plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_plain(v) for v in o]
    if o is None or isinstance(o, (str, int, float, bool)):
        return o
    return str(o)


def _to_cpu(o):
    if torch.is_tensor(o):
        return o.detach().cpu()
    if isinstance(o, dict):
        return {k: _to_cpu(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return type(o)(_to_cpu(v) for v in o)
    return o
