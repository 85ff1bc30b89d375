"""ORACLE — test infrastructure only.

CPU fp32 restatement of the reference's ContextUnet / DDPM hot path, written
from the reference's *behaviour* (file:line cited per function, relative to
/root/reference).  It is a functional restatement over plain `state_dict`
tensors with torch.nn.functional primitives: no reference module is imported.

Pinning: `oracle/make_golden.py` runs the UNMODIFIED reference modules in the
build container and stores their inputs/outputs under tests/golden/;
tests/test_oracle_golden.py checks this restatement against those vectors
(fp32, tolerance 2e-5 relative L2).  The reference has no tests or golden
vectors of its own, so reference-generated fixtures are the pin.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product path never does.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
GN_EPS = 1e-5


# --------------------------------------------------------------------------- schedule
def make_schedule(timesteps, beta1=1e-4, beta2=0.02, device="cpu"):
    """train_diffusion_paper.py:205-217 — linear beta, T+1 entries, ab_t[0] forced to 1."""
    b_t = (beta2 - beta1) * torch.linspace(0, 1, timesteps + 1, device=device) + beta1
    a_t = 1 - b_t
    ab_t = torch.cumsum(a_t.log(), dim=0).exp()
    ab_t[0] = 1
    return b_t, a_t, ab_t


def draw_shortcut(n_out=128, generator=None):
    """The fresh nn.Conv2d(1, n_out, 1) the reference builds on every forward
    (diffusion_utilities.py:54): kaiming_uniform(a=sqrt(5)) with fan_in 1 is U(-1,1)
    for the weight, then U(-1/sqrt(fan_in), ..) = U(-1,1) for the bias, drawn in that
    order from the global CPU generator."""
    w = torch.empty(n_out).uniform_(-1, 1, generator=generator)
    b = torch.empty(n_out).uniform_(-1, 1, generator=generator)
    return w, b


# --------------------------------------------------------------------------- blocks
def _rb(v, on):
    """bf16 rounding with a straight-through gradient.  `emulate_bf16=True` places it exactly where the
    sm_100a path stores bf16 (tensor-core weights, every conv / norm output), so that a test can separate
    rounding noise (ReLU masks flipping near zero) from real backward-pass errors.  Off by default: the
    oracle proper is the reference's fp32 arithmetic."""
    return v + (v.to(torch.bfloat16).float() - v).detach() if on else v


def _bn(x, sd, prefix, training, stats_out=None):
    """nn.BatchNorm2d (diffusion_utilities.py:28,35): eval = running stats; train = batch stats
    (biased var for normalisation; running stats get momentum 0.1 and the unbiased var)."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if not training:
        rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
        return (x - rm.view(1, -1, 1, 1)) / torch.sqrt(rv.view(1, -1, 1, 1) + BN_EPS) * w.view(1, -1, 1, 1) \
            + b.view(1, -1, 1, 1)
    mean = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=False)
    if stats_out is not None:
        n = x.numel() // x.shape[1]
        stats_out[prefix] = (mean.detach().clone(), (var * n / max(n - 1, 1)).detach().clone())
    return (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + BN_EPS) * w.view(1, -1, 1, 1) \
        + b.view(1, -1, 1, 1)


def _cbr(x, sd, prefix, training, stats_out, rb=False, round_out=True):
    """Conv3x3(s1,p1) -> BatchNorm2d -> ReLU (diffusion_utilities.py:26-37; G2: ReLU, not GELU)."""
    w = sd[prefix + ".0.weight"]
    y = _rb(F.conv2d(x, _rb(w, rb and w.shape[1] > 1), sd[prefix + ".0.bias"], padding=1), rb)
    return _rb(F.relu(_bn(y, sd, prefix + ".1", training, stats_out)), rb and round_out)


def _rcb(x, sd, prefix, training, stats_out, rb=False):
    """ResidualConvBlock with is_res=False (diffusion_utilities.py:62-65)."""
    return _cbr(_cbr(x, sd, prefix + ".conv1", training, stats_out, rb), sd, prefix + ".conv2", training, stats_out,
                rb)


def _embed(v, sd, prefix, in_dim):
    """EmbedFC (diffusion_utilities.py:137-145): view(-1,in) -> Linear -> GELU(erf) -> Linear."""
    v = v.reshape(-1, in_dim).float()
    h = F.gelu(F.linear(v, sd[prefix + ".model.0.weight"], sd[prefix + ".model.0.bias"]))
    return F.linear(h, sd[prefix + ".model.2.weight"], sd[prefix + ".model.2.bias"])


def unet_forward(sd, x, t, c, shortcut, *, n_feat=128, n_cfeat=6, height=64, training=False, stats_out=None,
                 taps=None, emulate_bf16=False):
    """ContextUnet.forward (ContextUnet.py:42-60) on a state_dict `sd`.

    shortcut = (w_c[n_feat], b_c[n_feat]) — the fresh 1x1 conv of this call (G1).
    t has numel 1 or B (G4); c may be None -> zeros (ContextUnet.py:48-49).
    taps: optional dict that receives named intermediate activations.
    """
    B = x.shape[0]
    w_c, b_c = shortcut
    rb = emulate_bf16
    # init_conv: is_res=True, in!=out channels -> conv2(conv1(x)) + fresh 1x1 conv(x), no /1.414
    x1 = _cbr(x, sd, "init_conv.conv1", training, stats_out, rb)
    x2 = _cbr(x1, sd, "init_conv.conv2", training, stats_out, rb, round_out=False)
    x0 = _rb(x2 + (x * w_c.view(1, -1, 1, 1) + b_c.view(1, -1, 1, 1)), rb)
    # UnetDown = RCB, RCB, MaxPool2d(2)  (diffusion_utilities.py:109)
    d1 = F.max_pool2d(_rcb(_rcb(x0, sd, "down1.model.0", training, stats_out, rb), sd, "down1.model.1", training,
                           stats_out, rb), 2)
    d2 = F.max_pool2d(_rcb(_rcb(d1, sd, "down2.model.0", training, stats_out, rb), sd, "down2.model.1", training,
                           stats_out, rb), 2)
    hidden = _rb(F.gelu(F.avg_pool2d(d2, height // 4)), rb)  # to_vec (ContextUnet.py:17)
    if c is None:
        c = torch.zeros(B, n_cfeat)
    cemb1 = _embed(c, sd, "contextembed1", n_cfeat).view(-1, 2 * n_feat, 1, 1)
    temb1 = _embed(t, sd, "timeembed1", 1).view(-1, 2 * n_feat, 1, 1)
    cemb2 = _embed(c, sd, "contextembed2", n_cfeat).view(-1, n_feat, 1, 1)
    temb2 = _embed(t, sd, "timeembed2", 1).view(-1, n_feat, 1, 1)
    # up0: ConvT(k=s=h/4) on a 1x1 map == GEMM; GroupNorm(8); ReLU  (ContextUnet.py:26-30)
    u0 = torch.einsum("ni,iokl->nokl", hidden.view(B, -1), _rb(sd["up0.0.weight"], rb)) \
        + sd["up0.0.bias"].view(1, -1, 1, 1)
    u0 = F.relu(F.group_norm(_rb(u0, rb), 8, sd["up0.1.weight"], sd["up0.1.bias"], GN_EPS))

    def unet_up(a, skip, prefix):
        # UnetUp: cat(x, skip) -> ConvT 2x2 s2 -> RCB, RCB  (diffusion_utilities.py:86-100)
        z = torch.cat((a, skip), 1)
        v = _rb(F.conv_transpose2d(z, _rb(sd[prefix + ".model.0.weight"], rb), sd[prefix + ".model.0.bias"],
                                   stride=2), rb)
        return _rcb(_rcb(v, sd, prefix + ".model.1", training, stats_out, rb), sd, prefix + ".model.2", training,
                    stats_out, rb)

    u1 = unet_up(_rb(cemb1 * u0 + temb1, rb), d2, "up1")
    u2 = unet_up(_rb(cemb2 * u1 + temb2, rb), d1, "up2")
    o_raw = _rb(F.conv2d(torch.cat((u2, x0), 1), _rb(sd["out.0.weight"], rb), sd["out.0.bias"], padding=1), rb)
    o = _rb(F.relu(F.group_norm(o_raw, 8, sd["out.1.weight"], sd["out.1.bias"], GN_EPS)), rb)
    eps = F.conv2d(o, sd["out.3.weight"], sd["out.3.bias"], padding=1)
    if taps is not None:
        taps.update(x1=x1, x0=x0, d1=d1, d2=d2, hidden=hidden, cemb1=cemb1, temb1=temb1, cemb2=cemb2, temb2=temb2,
                    u0=u0, u1=u1, u2=u2, o=o, o_raw=o_raw)
    return eps


# --------------------------------------------------------------------------- diffusion process
def perturb_input(x, t, noise, ab_t):
    """train_diffusion_paper.py:320-321 — note (1 - ab_t), NOT sqrt(1 - ab_t) (G3)."""
    return ab_t.sqrt()[t, None, None, None] * x + (1 - ab_t[t, None, None, None]) * noise


def denoise_add_noise(x, t, pred_noise, z, b_t, a_t, ab_t):
    """train_diffusion_paper.py:548-553."""
    noise = b_t.sqrt()[t] * z
    mean = (x - pred_noise * ((1 - a_t[t]) / (1 - ab_t[t]).sqrt())) / a_t[t].sqrt()
    return mean + noise


def snapshot_steps(timesteps, save_rate=20):
    """Steps whose x is appended to `intermediate` (train_diffusion_paper.py:617-618)."""
    return [i for i in range(timesteps, 0, -1) if i % save_rate == 0 or i == timesteps or i < 8]


def sample_ddpm(sd, x_T, params, guide_w, timesteps, sched, z_all, shortcuts, *, n_cfeat=6, save_rate=20):
    """train_diffusion_paper.py:555-623 with every random draw made explicit.

    z_all[k] is the z of loop iteration k (i = T-k; unused at i == 1, where z = 0);
    shortcuts[k] is a list of (w_c, b_c), one per forward of that iteration in call order
    (conditional first, then unconditional — G5).  params=None -> c=None (from_noise variant :646-672).
    Returns (x, intermediate[n_snap,B,1,64,64])."""
    b_t, a_t, ab_t = sched
    x = x_T.clone()
    inter = []
    for k, i in enumerate(range(timesteps, 0, -1)):
        t = torch.tensor([i / timesteps])
        z = z_all[k] if i > 1 else 0
        if guide_w > 0 and params is not None:
            e_c = unet_forward(sd, x, t, params, shortcuts[k][0], n_cfeat=n_cfeat)
            e_u = unet_forward(sd, x, t, torch.zeros_like(params), shortcuts[k][1], n_cfeat=n_cfeat)
            eps = e_u + guide_w * (e_c - e_u)
        else:
            eps = unet_forward(sd, x, t, params, shortcuts[k][0], n_cfeat=n_cfeat)
        x = denoise_add_noise(x, i, eps, z, b_t, a_t, ab_t)
        if i % save_rate == 0 or i == timesteps or i < 8:
            inter.append(x.numpy().copy())
    return x, np.stack(inter)


def per_sample_mse(pred, noise):
    """F.mse_loss(reduction='none').mean(dim=[1,2,3]) (train_diffusion_paper.py:119,173)."""
    return ((pred - noise) ** 2).mean(dim=(1, 2, 3))


def likelihood_batch(sd, x, param, timesteps, sched, noises, shortcuts, *, n_cfeat=6):
    """Inner loop of calculate_likelihood (train_diffusion_paper.py:163-178) for one batch:
    returns batch_nll[B] = sum_t mse_t / (2 b_t).  noises[t-1], shortcuts[t-1] are the draws of step t."""
    b_t, a_t, ab_t = sched
    nll = torch.zeros(x.shape[0])
    for t in range(1, timesteps + 1):
        noise = noises[t - 1]
        x_t = ab_t.sqrt()[t, None, None, None] * x + (1 - ab_t[t, None, None, None]) * noise
        pred = unet_forward(sd, x_t, torch.tensor([t / timesteps]), param, shortcuts[t - 1], n_cfeat=n_cfeat)
        nll += per_sample_mse(pred, noise) / (2 * b_t[t])
    return nll


def elbo_bpd_batch(pred_noise, noise, t, ab_t, dims):
    """Per-batch ELBO of train_diffusion_elbo.py:74-105."""
    mse = per_sample_mse(pred_noise, noise)
    weight = 0.5 * (1.0 / (1.0 - ab_t[t]) - 1.0)
    elbo = (weight * mse).mean()
    return elbo, elbo / (dims * np.log(2))


def elbo_paper_batch(sd, x, param, timesteps, sched, noises, shortcuts, *, n_cfeat=6):
    """Inner loop of the dataloader ELBO (train_diffusion_paper.py:107-127): 10 timesteps
    linspace(1,T,10).long(), sqrt(1-ab_t) perturbation, weight 0.5 b_t/(1-ab_t), t<=1 skipped, /10."""
    b_t, a_t, ab_t = sched
    out = torch.zeros(x.shape[0])
    for k, t in enumerate(torch.linspace(1, timesteps, 10).long()):
        noise = noises[k]
        x_t = ab_t.sqrt()[t] * x + torch.sqrt(1 - ab_t[t]) * noise
        pred = unet_forward(sd, x_t, torch.tensor([t / timesteps]), param, shortcuts[k], n_cfeat=n_cfeat)
        if t > 1:
            out += 0.5 * (b_t[t] / (1.0 - ab_t[t])) * per_sample_mse(pred, noise) / 10.0
    return out


def train_step(sd, x, param, t, noise, shortcut, timesteps, ab_t, *, n_cfeat=6, emulate_bf16=False):
    """Loss + gradients of one training step (train_diffusion_paper.py:351-363), train-mode BN.
    Returns (loss, grads{name: tensor}, bn_batch_stats{prefix: (mean, unbiased_var)})."""
    params = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k)
              for k, v in sd.items()}
    x_pert = perturb_input(x, t, noise, ab_t)
    stats = {}
    pred = unet_forward(params, x_pert, t / timesteps, param, shortcut, n_cfeat=n_cfeat, training=True,
                        stats_out=stats, emulate_bf16=emulate_bf16)
    loss = F.mse_loss(pred, noise)
    names = [k for k, v in params.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [params[k] for k in names])
    return loss.detach(), dict(zip(names, grads)), stats


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (train_diffusion_paper.py:318): returns updated (p, m, v)."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    mhat = m / (1 - beta1 ** step)
    vhat = v / (1 - beta2 ** step)
    return p - lr * mhat / (vhat.sqrt() + eps), m, v


def calibrate_state_dict(sd, seed=1234):
    """Synthetic 'calibrated' weights: randomised BN/GN affine (seeded) so that
    normalisation folding, FiLM and context paths carry signal.  Running stats are filled by
    make_golden.py with train-mode forwards of the reference and shipped in the fixture."""
    g = torch.Generator().manual_seed(seed)
    out = {k: v.clone() for k, v in sd.items()}
    for k in out:
        is_norm_w = (k.endswith(".1.weight") and out[k].dim() == 1)
        is_norm_b = (k.endswith(".1.bias") and out[k].dim() == 1)
        if is_norm_w:
            out[k] = torch.empty_like(out[k]).uniform_(0.5, 1.5, generator=g)
        elif is_norm_b:
            out[k] = torch.empty_like(out[k]).normal_(0, 0.2, generator=g)
    return out


# --------------------------------------------------------------------------- synthetic weights
def init_state_dict(seed, n_feat=128, n_cfeat=6, height=64, in_channels=1):
    """Random-init weights identical to `torch.manual_seed(seed); ContextUnet(...)` of the
    reference: the same torch.nn layers are constructed in the same order as
    ContextUnet.__init__ (ContextUnet.py:6-40) / the block constructors
    (diffusion_utilities.py:14-37,80-92,104-112,119-135), so the global generator is consumed
    identically.  Keys follow the reference state_dict."""
    import torch.nn as nn
    torch.manual_seed(seed)
    sd = {}

    def put(prefix, mod):
        for k, v in mod.state_dict().items():
            sd[f"{prefix}.{k}"] = v.detach().clone()

    def rcb(prefix, cin, cout):
        put(prefix + ".conv1.0", nn.Conv2d(cin, cout, 3, 1, 1))
        put(prefix + ".conv1.1", nn.BatchNorm2d(cout))
        put(prefix + ".conv2.0", nn.Conv2d(cout, cout, 3, 1, 1))
        put(prefix + ".conv2.1", nn.BatchNorm2d(cout))

    def embed(prefix, din, demb):
        put(prefix + ".model.0", nn.Linear(din, demb))
        put(prefix + ".model.2", nn.Linear(demb, demb))

    rcb("init_conv", in_channels, n_feat)
    rcb("down1.model.0", n_feat, n_feat)
    rcb("down1.model.1", n_feat, n_feat)
    rcb("down2.model.0", n_feat, 2 * n_feat)
    rcb("down2.model.1", 2 * n_feat, 2 * n_feat)
    embed("timeembed1", 1, 2 * n_feat)
    embed("timeembed2", 1, n_feat)
    embed("contextembed1", n_cfeat, 2 * n_feat)
    embed("contextembed2", n_cfeat, n_feat)
    put("up0.0", nn.ConvTranspose2d(2 * n_feat, 2 * n_feat, height // 4, height // 4))
    put("up0.1", nn.GroupNorm(8, 2 * n_feat))
    put("up1.model.0", nn.ConvTranspose2d(4 * n_feat, n_feat, 2, 2))
    rcb("up1.model.1", n_feat, n_feat)
    rcb("up1.model.2", n_feat, n_feat)
    put("up2.model.0", nn.ConvTranspose2d(2 * n_feat, n_feat, 2, 2))
    rcb("up2.model.1", n_feat, n_feat)
    rcb("up2.model.2", n_feat, n_feat)
    put("out.0", nn.Conv2d(2 * n_feat, n_feat, 3, 1, 1))
    put("out.1", nn.GroupNorm(8, n_feat))
    put("out.3", nn.Conv2d(n_feat, in_channels, 3, 1, 1))
    return sd


def state_dict_checksum(sd):
    """Order-independent fingerprint used to check that seeded weights reproduce on another box."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].double()
        tot += float(v.sum()) + 0.5 * float(v.abs().sum())
    return tot
