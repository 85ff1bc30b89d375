"""Data preparation of the reference's training / evaluation scripts, on the device.

Mirrors the module-level code of code/train_diffusion_paper.py:232-262: the CAMELS maps [N,256,256] are shifted to
positive values, divided by their maximum, log10'd, min-max normalised to [0,1] and resized to 64x64 with
F.interpolate(mode='bilinear'); the parameter table [N/15, 6] is repeated 15x, min-max normalised per column and
cut / padded to `num_params` columns.  One reduction + one fused transform/resize kernel for the maps (the
full-resolution intermediate tensors of the reference are never materialised), one kernel for the parameters.
"""
import numpy as np
import torch

from . import _lib as L


def preprocess_maps(camels_data, size=64, device=None):
    """camels_data: numpy / tensor [N,H,W] (any float dtype; computed in fp32 as the reference's float32 maps)
    -> fp32 device tensor [N,1,size,size] == `camels_data_resized` of train_diffusion_paper.py:261."""
    dev = torch.device("cuda") if device is None else torch.device(device)
    raw = torch.as_tensor(camels_data).to(dev, torch.float32).contiguous()
    if raw.dim() != 3:
        raise L.CdmError("preprocess_maps: expected [N,H,W] maps")
    ws = torch.zeros(2 * L.num_sms() * 8 + 1, device=dev)
    mm = torch.empty(2, device=dev)
    L.minmax(raw, ws, mm)
    out = torch.empty(raw.shape[0], size, size, device=dev)
    L.preprocess_maps(raw, mm, out)
    return out.unsqueeze(1)


def normalize_params(param_data, num_params, images_per_param=15, device=None):
    """param_data [n_sets, n_cols] -> (param_data_tensor fp32 [n_sets*images_per_param, num_params] on the device,
    param_min [1,n_cols], param_max [1,n_cols] as numpy — the arrays the reference saves for generation,
    train_diffusion_paper.py:236-252)."""
    dev = torch.device("cuda") if device is None else torch.device(device)
    x = torch.as_tensor(np.asarray(param_data)).to(dev, torch.float32).contiguous()
    rows, cols = x.shape
    out = torch.empty(rows * images_per_param, num_params, device=dev)
    cmin, cmax = torch.empty(cols, device=dev), torch.empty(cols, device=dev)
    L.normalize_params(x, images_per_param, out, cmin, cmax)
    return out, cmin.cpu().numpy()[None, :], cmax.cpu().numpy()[None, :]
