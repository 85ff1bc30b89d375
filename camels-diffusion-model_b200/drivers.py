"""Callers of the sampler: parameter grid, guidance-strength sweep and per-parameter sensitivity
(the module-level driver code of code/train_diffusion_paper.py:913-954, 1009-1075, 1109-1182),
as library functions over a `DDPM` process.  Context construction follows the reference line by line;
the sampling itself is the CUDA-graph sampler of diffusion.py.
"""
import torch

from .diffusion import DDPM


def parameter_grid_contexts(base_param, num_params):
    """25 contexts: 5x5 linspace(0,1,5) over parameters 0 and 1 (or a 25-point sweep of parameter 0 if there
    is only one), every other entry copied from `base_param` (train_diffusion_paper.py:917-942)."""
    grid = []
    if num_params >= 2:
        for p1 in torch.linspace(0.0, 1.0, 5):
            for p2 in torch.linspace(0.0, 1.0, 5):
                q = base_param.clone()
                q[0], q[1] = p1, p2
                grid.append(q)
    else:
        for p1 in torch.linspace(0.0, 1.0, 25):
            q = base_param.clone()
            q[0] = p1
            grid.append(q)
    return torch.stack(grid)


def sample_parameter_grid(ddpm: DDPM, base_param, **kw):
    """One batched sample_ddpm over the 25 grid contexts (paper.py:944-947) -> (samples, intermediate, time, ctx)."""
    ctx = parameter_grid_contexts(base_param.cpu(), ddpm.n_cfeat)
    x, inter, dt, _ = ddpm.sample_ddpm(n_sample=len(ctx), size=ddpm.nn_model.h, params=ctx, **kw)
    return x, inter, dt, ctx


def guidance_sweep(ddpm: DDPM, base_param, strengths=(0.0, 1.0, 2.0, 3.0, 5.0), n_sample=5, **kw):
    """sample_ddpm(n_sample=5, params=base repeated, guide_w=w) for each strength (paper.py:1009-1019).
    Returns {w: (samples, sampling_time)}.  w == 0 runs ONE conditional pass per step (G5)."""
    out = {}
    params = base_param.cpu().unsqueeze(0).repeat(n_sample, 1)
    for w in strengths:
        x, _, dt, _ = ddpm.sample_ddpm(n_sample=n_sample, size=ddpm.nn_model.h, params=params, guide_w=w, **kw)
        out[float(w)] = (x, dt)
    return out


def sensitivity_contexts(base_param, num_params):
    """num_params x 5 contexts: parameter k swept over linspace(0,1,5), the rest from base (paper.py:1114-1124)."""
    ctx = []
    for k in range(num_params):
        for val in torch.linspace(0.0, 1.0, 5):
            q = base_param.clone()
            q[k] = val
            ctx.append(q)
    return torch.stack(ctx)


def parameter_sensitivity(ddpm: DDPM, base_param, batched=True, **kw):
    """The reference runs num_params*5 separate batch-1 sample_ddpm calls (1500 sequential launch-bound steps each,
    paper.py:1126-1127).  batched=True draws the same contexts as ONE batch (independent samples, so the result
    distribution is identical and the GPU is actually filled); batched=False keeps the reference's call pattern
    (one CUDA-graph replay chain per context)."""
    ctx = sensitivity_contexts(base_param.cpu(), ddpm.n_cfeat)
    if batched:
        x, _, dt, _ = ddpm.sample_ddpm(n_sample=len(ctx), size=ddpm.nn_model.h, params=ctx, **kw)
        return x, ctx, dt
    xs, total = [], 0.0
    for q in ctx:
        x, _, dt, _ = ddpm.sample_ddpm(n_sample=1, size=ddpm.nn_model.h, params=q.unsqueeze(0), **kw)
        xs.append(x)
        total += dt
    return torch.cat(xs), ctx, total
