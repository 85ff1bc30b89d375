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


# --------------------------------------------------------------------------- the training script's epoch loop
def train_diffusion(nn_model, train_dataloader, n_epoch, lrate, timesteps, *, test_dataloader=None, save_dir=None,
                    eval_every=5, save_every=25, likelihood_subset=200, elbo_subset=2000, seed=0, log=None,
                    resume_from=None, use_graph=True):
    """The epoch loop of code/train_diffusion_paper.py:338-478 as a library entry point, on the B200 path:

    * `optim.param_groups[0]['lr'] = lrate * (1 - ep / n_epoch)` per epoch (:343) -> GraphedTrainStep.set_lr;
    * per batch: noise, `t ~ randint(1, T+1)`, perturb_input, train-mode forward, MSE, backward, Adam (:349-366) =
      one replay of the captured step (the ragged last batch of an epoch gets its own graph); the loss stays on the
      device and is read once per epoch instead of `loss.item()` every step;
    * every `eval_every` epochs and at the end (:386-470): validation MSE, ELBO/BPD on (a subset of) the training set and
      on the validation set, all-timestep NLL on `likelihood_subset` random samples of each;
    * `model_epoch_{ep+1}.pth` every `save_every` epochs and at the end (:473-474, the reference's weights-only
      format) plus `resume.pt` (optimizer, counters, generator state) so that an interrupted run continues.

    Data parallel: one process per GPU, each with its own dataloader shard (e.g. DistributedSampler); BatchNorm
    statistics and gradients are exchanged inside the step, rank 0 writes the files.  Returns the logs as a dict of
    lists with the reference's names."""
    import os
    import random
    import time

    from torch.utils.data import DataLoader, Subset

    from . import checkpoint as CK
    from . import diffusion as D
    from . import train as TR
    from .parallel import world

    dev = nn_model._check_supported()
    rank, _ = world()
    b_t, a_t, ab_t = D.make_schedule(timesteps, device=dev)
    batch_size = max(x.shape[0] for x, _ in [next(iter(train_dataloader))])
    step = TR.GraphedTrainStep(nn_model.train(), batch_size, timesteps, ab_t, lr=lrate, seed=seed, use_graph=use_graph)
    logs = {k: [] for k in ("loss_log", "val_loss_log", "likelihood_log", "val_likelihood_log", "elbo_log", "bpd_log",
                            "val_elbo_log", "val_bpd_log", "epoch_times")}
    first_epoch = 0
    if resume_from is not None:
        first_epoch, _, extra = CK.load_checkpoint(resume_from, nn_model, step)
        logs.update(extra.get("logs", {}))
    say = log if log is not None else (print if rank == 0 else (lambda *_: None))

    def subset_loader(loader, k):
        ds = loader.dataset
        idx = random.sample(range(len(ds)), min(len(ds), k))  # :405, :437, :446
        return DataLoader(Subset(ds, idx), batch_size=loader.batch_size or batch_size, shuffle=False)

    for ep in range(first_epoch, n_epoch):
        t0 = time.time()
        nn_model.train()
        step.set_lr(lrate * (1 - ep / n_epoch))  # :343
        loss_acc = torch.zeros((), device=dev)
        n_batches = 0
        for x, param in train_dataloader:
            loss_acc += step(x.to(dev, non_blocking=True), param.to(dev, non_blocking=True))
            n_batches += 1
        logs["loss_log"].append(float(loss_acc) / max(n_batches, 1))  # the epoch's only host sync
        logs["epoch_times"].append(time.time() - t0)
        say(f"Epoch {ep + 1}/{n_epoch} completed in {logs['epoch_times'][-1]:.2f} seconds; "
            f"Training Loss: {logs['loss_log'][-1]:.6f}")
        if test_dataloader is not None and (ep % eval_every == 0 or ep == n_epoch - 1):
            nn_model.eval()
            val, nb = torch.zeros((), device=dev), 0
            with torch.no_grad():
                for x, param in test_dataloader:  # :392-404
                    x = x.to(dev)
                    noise = torch.randn(x.shape).to(dev)
                    t = torch.randint(1, timesteps + 1, (x.shape[0],))
                    x_pert = D.perturb_input(x, t, noise, ab_t)
                    pred = nn_model(x_pert, (t / timesteps).to(dev), param.to(dev))
                    val += torch.nn.functional.mse_loss(pred, noise)
                    nb += 1
            logs["val_loss_log"].append(float(val) / max(nb, 1))
            tl = subset_loader(train_dataloader, elbo_subset) if len(train_dataloader.dataset) > elbo_subset \
                else train_dataloader
            e, b = D.calculate_elbo_and_bpd(nn_model, tl, timesteps, dev, ab_t, b_t, a_t)
            ve, vb = D.calculate_elbo_and_bpd(nn_model, test_dataloader, timesteps, dev, ab_t, b_t, a_t)
            logs["elbo_log"].append(e), logs["bpd_log"].append(b)
            logs["val_elbo_log"].append(ve), logs["val_bpd_log"].append(vb)
            logs["likelihood_log"].append(D.calculate_likelihood(
                nn_model, subset_loader(train_dataloader, likelihood_subset), timesteps, dev, ab_t, b_t, a_t))
            logs["val_likelihood_log"].append(D.calculate_likelihood(
                nn_model, subset_loader(test_dataloader, likelihood_subset), timesteps, dev, ab_t, b_t, a_t))
            say(f"  Val Loss: {logs['val_loss_log'][-1]:.6f}  Train BPD: {b:.6f}  Val BPD: {vb:.6f}  "
                f"Train NLL: {logs['likelihood_log'][-1]:.6f}  Val NLL: {logs['val_likelihood_log'][-1]:.6f}")
            nn_model.train()
        if save_dir is not None and ((ep + 1) % save_every == 0 or ep == n_epoch - 1):
            os.makedirs(save_dir, exist_ok=True)
            CK.save_model(nn_model, os.path.join(save_dir, f"model_epoch_{ep + 1}.pth"))  # :473-474
            CK.save_checkpoint(os.path.join(save_dir, "resume.pt"), nn_model, step, epoch=ep + 1, extra={"logs": logs})
    return logs
