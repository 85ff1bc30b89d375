"""DDPM process around ContextUnet: schedule, perturb_input, denoise_add_noise, the
1500-step ancestral sampler with classifier-free guidance, the all-timestep NLL and the
two ELBO/BPD variants.  Mirrors the closures / functions of the reference scripts
(code/train_diffusion_paper.py:77-183,205-217,320-321,548-686; code/sample_power_spectra.py:64-110;
code/train_diffusion_elbo.py:74-105) — same names, argument meaning and return values.

The sampling step (1 or 2 U-Net passes + CFG mix + x_{t-1} update) is captured once as a CUDA
graph and replayed `timesteps` times; the step index lives in a device int that the kernels
read (time-embedding row, shortcut row, schedule coefficients, noise offset, snapshot slot).
"""
import time

import numpy as np
import torch

from . import _lib as L


# --------------------------------------------------------------------------- schedule
def make_schedule(timesteps, beta1=1e-4, beta2=0.02, device="cuda"):
    """b_t, a_t, ab_t with T+1 entries (train_diffusion_paper.py:205-217)."""
    b_t = (beta2 - beta1) * torch.linspace(0, 1, timesteps + 1, device=device) + beta1
    a_t = 1 - b_t
    ab_t = torch.cumsum(a_t.log(), dim=0).exp()
    ab_t[0] = 1
    return b_t, a_t, ab_t


def _coef_table(b_t, a_t, ab_t):
    """[T+1][4] fp32: (1-a_t)/sqrt(1-ab_t), sqrt(a_t), sqrt(b_t), 0 — the scalars of
    denoise_add_noise evaluated with the reference's own torch expressions (paper.py:551-552)."""
    k2 = (1 - a_t) / (1 - ab_t).sqrt()
    return torch.stack([k2, a_t.sqrt(), b_t.sqrt(), torch.zeros_like(b_t)], 1).float().contiguous()


def _time_table(model, timesteps, dev):
    """temb1/temb2 for every t = i/T, i = 0..T: [T+1, 2nf], [T+1, nf] (row i <-> tensor([i / timesteps]))."""
    tv = (torch.arange(timesteps + 1, dtype=torch.float64) / timesteps).float().to(dev).view(-1, 1)
    return model.timeembed1(tv), model.timeembed2(tv)


def draw_shortcut_table(timesteps, reps, n_feat=128):
    """Shortcut (w_c, b_c) draws for a whole sampling run, in the order the reference's loop consumes the
    global CPU generator (one fresh nn.Conv2d(1,n_feat,1) per forward: weight U(-1,1)[n_feat] then
    bias U(-1,1)[n_feat]; conditional pass before the unconditional one).  Returns [T+1][reps][2][n_feat]
    indexed by the step i (row 0 unused)."""
    tab = torch.zeros(timesteps + 1, reps, 2, n_feat)
    for i in range(timesteps, 0, -1):
        for r in range(reps):
            tab[i, r, 0].uniform_(-1, 1)
            tab[i, r, 1].uniform_(-1, 1)
    return tab


def shortcut_table_from_list(shortcuts, timesteps, reps, order="descending"):
    """Recorded per-forward (w_c, b_c) pairs -> step-indexed table.  order='descending': iteration k is
    step T-k (sampler); 'ascending': iteration k is step k+1 (likelihood loop)."""
    n_feat = shortcuts[0][0].numel()
    tab = torch.zeros(timesteps + 1, reps, 2, n_feat)
    for k in range(timesteps):
        i = timesteps - k if order == "descending" else k + 1
        for r in range(reps):
            w, b = shortcuts[k * reps + r]
            tab[i, r, 0], tab[i, r, 1] = w.view(-1), b.view(-1)
    return tab


def snapshot_steps(timesteps, save_rate=20):
    return [i for i in range(timesteps, 0, -1) if i % save_rate == 0 or i == timesteps or i < 8]


# --------------------------------------------------------------------------- elementwise API
def perturb_input(x, t, noise, ab_t):
    """sqrt(ab_t[t]) * x + (1 - ab_t[t]) * noise  (train_diffusion_paper.py:320-321; not sqrt(1-ab_t))."""
    dev = ab_t.device
    x = x.to(dev, torch.float32).contiguous()
    noise = noise.to(dev, torch.float32).contiguous()
    out = torch.empty_like(x)
    ca, cb = ab_t.sqrt().contiguous(), (1 - ab_t).contiguous()
    if torch.is_tensor(t) and t.numel() > 1:
        L.perturb(x, out, ca, cb, noise=noise, t_idx=t.to(dev, torch.int64).contiguous())
    else:
        L.perturb(x, out, ca, cb, noise=noise, t_shared=int(t))
    return out


def denoise_add_noise(x, t, pred_noise, z=None, b_t=None, a_t=None, ab_t=None):
    """(x - pred_noise*(1-a_t)/sqrt(1-ab_t))/sqrt(a_t) + sqrt(b_t)*z  (sample_power_spectra.py:64-69)."""
    dev = b_t.device
    out = x.to(dev, torch.float32).clone().contiguous()
    eps = pred_noise.to(dev, torch.float32).contiguous()
    if z is None:
        z = torch.randn_like(out)
    elif not torch.is_tensor(z):
        z = torch.full_like(out, float(z))
    t = int(t)
    # the kernel zeroes z at step 1 (the sampler's convention); apply the explicit z through a 2-step table
    coef = _coef_table(b_t, a_t, ab_t)
    tab = torch.stack([coef[t], coef[t], coef[t]], 0).contiguous()
    L.ddpm_step(out, eps, tab, 2, reps=1, step=2, z=z.to(dev, torch.float32).contiguous(), z_iter_stride=0)
    return out


# --------------------------------------------------------------------------- sampler
class _SamplerRun:
    """Device state of one sampling run + the captured one-step CUDA graph."""

    def __init__(self, model, *a, **k):
        self.dev = model._check_supported()  # every launch below runs with the model's device current (L.on_device)
        self._init(model, *a, **k)

    @L.on_device
    def _init(self, model, x_T, params, guide_w, timesteps, sched, *, z_all=None, shortcut_tab=None,
              save_rate=20, seed=None, use_graph=True, snapshots=True, sample_offset=0):
        dev = model._check_supported()
        if model.training:
            raise L.CdmError("sampling needs eval mode (BatchNorm running statistics): call model.eval()")
        self.model, self.T, self.dev = model, timesteps, dev
        b_t, a_t, ab_t = (s.to(dev, torch.float32) for s in sched)
        B = x_T.shape[0]
        self.B = B
        self.cfg = bool(guide_w > 0 and params is not None)
        self.reps = 2 if self.cfg else 1
        self.guide_w = float(guide_w)
        hw = model.h * model.h
        self.x = x_T.detach().to(dev, torch.float32).reshape(B, 1, model.h, model.h).clone().contiguous()
        c = torch.zeros(B, model.n_cfeat, device=dev) if params is None else params.to(dev, torch.float32)
        c_all = torch.cat([c, torch.zeros_like(c)], 0) if self.cfg else c
        self.cemb1, self.cemb2 = model.contextembed1(c_all.contiguous()), model.contextembed2(c_all.contiguous())
        self.temb1, self.temb2 = _time_table(model, timesteps, dev)
        if shortcut_tab is None:
            shortcut_tab = draw_shortcut_table(timesteps, self.reps, model.n_feat)
        assert tuple(shortcut_tab.shape) == (timesteps + 1, self.reps, 2, model.n_feat)
        self.sc_tab = shortcut_tab.to(dev, torch.float32).contiguous()
        self.coef = _coef_table(b_t, a_t, ab_t)
        self.step = torch.full((1,), timesteps, device=dev, dtype=torch.int32)
        self.snap_steps = snapshot_steps(timesteps, save_rate) if snapshots else []
        if self.snap_steps:
            slot = torch.full((timesteps + 1,), -1, dtype=torch.int32)
            for k, i in enumerate(self.snap_steps):
                slot[i] = k
            self.snap_slot = slot.to(dev)
            self.snap = torch.empty(len(self.snap_steps), B, 1, model.h, model.h, device=dev)
        else:
            self.snap_slot, self.snap = None, None
        if z_all is not None:
            self.z = z_all.to(dev, torch.float32).reshape(timesteps, B * hw).contiguous()
            self.z_stride = B * hw
        else:
            self.z, self.z_stride = None, 0
        self.seed = int(seed) if seed is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
        # in-kernel noise is keyed by the GLOBAL sample index: a shard [offset, offset + B) of a larger batch draws
        # exactly what the unsharded run draws for those samples, and two ranks never share noise
        self.sample_offset = int(sample_offset)
        self.graph = None
        self.use_graph = use_graph
        self.x_T = self.x.clone()
        self.remaining = timesteps  # host mirror of the device step counter: steps left before i reaches 0

    @L.on_device
    def reset(self):
        """Back to x_T at step T (a new trajectory over the same graph; snapshots are overwritten)."""
        self.x.copy_(self.x_T)
        self.step.fill_(self.T)
        self.remaining = self.T

    @L.on_device
    def _one_step(self):
        """One reverse-diffusion step = ONE C call (cdm_sample_step: the 26 launches of the reps*B-image forward,
        the CFG mix + x_{t-1} update, the step counter)."""
        m = self.model
        pl = m.plan(self.B, self.reps)[0]
        pl.sample_step(self.x, self.sc_tab, self.cemb1, self.temb1, self.cemb2, self.temb2, self.step, self.coef,
                       self.T, guide_w=self.guide_w, z=self.z, z_iter_stride=self.z_stride, seed=self.seed,
                       sample_offset=self.sample_offset, snap=self.snap, snap_slot=self.snap_slot)

    @L.on_device
    def capture(self):
        """Warm up once on scratch state (sets kernel attributes, fills caches), then capture one step."""
        x_keep = self.x.clone()
        self._one_step()
        torch.cuda.synchronize()
        self.x.copy_(x_keep)
        self.step.fill_(self.T)
        if self.use_graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._one_step()
            self.x.copy_(x_keep)
            self.step.fill_(self.T)
            self.graph = g
        torch.cuda.synchronize()

    @L.on_device
    def run(self, n_steps=None):
        n_steps = self.remaining if n_steps is None else n_steps
        if n_steps > self.remaining:  # the step index selects table rows on the device: never let it pass 0
            raise L.CdmError(f"{n_steps} steps requested but only {self.remaining} of {self.T} remain: reset() first")
        self.remaining -= n_steps
        for _ in range(n_steps):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._one_step()

    def intermediate(self):
        if self.snap is None:
            return np.zeros((0,) + tuple(self.x.shape), np.float32)
        return self.snap.cpu().numpy()


def _sample(model, x_T, params, guide_w, timesteps, sched, **kw):
    run = _SamplerRun(model, x_T, params, guide_w, timesteps, sched, **kw)
    run.capture()
    t0 = time.time()
    run.run()
    torch.cuda.synchronize()
    dt = time.time() - t0
    return run.x, run.intermediate(), dt


@torch.no_grad()
def sample_ddpm(model, n_sample=1, size=64, device=None, params=None, guide_w=0.0, timesteps=1000, b_t=None,
                a_t=None, ab_t=None, **kw):
    """Pure-function sampler of code/sample_power_spectra.py:71-110 -> x [n,1,size,size]."""
    dev = model._check_supported() if device is None else torch.device(device)
    x = torch.randn(n_sample, 1, size, size)  # CPU generator, as the reference (:79)
    if params is None:
        params = torch.rand(n_sample, 6)  # :83 (hard-coded 6 in the reference's pure-function form)
    x, _, _ = _sample(model, x.to(dev), params.to(dev), guide_w, timesteps, (b_t, a_t, ab_t), snapshots=False, **kw)
    return x


class DDPM:
    """The closures of code/train_diffusion_paper.py bound to one model + schedule."""

    def __init__(self, nn_model, timesteps, beta1=1e-4, beta2=0.02, device=None):
        self.nn_model = nn_model
        self.timesteps = timesteps
        self.device = torch.device(device) if device is not None else nn_model.out[3].weight.device
        self.b_t, self.a_t, self.ab_t = make_schedule(timesteps, beta1, beta2, self.device)
        self.n_cfeat = nn_model.n_cfeat

    @property
    def sched(self):
        return self.b_t, self.a_t, self.ab_t

    def perturb_input(self, x, t, noise):
        return perturb_input(x, t, noise, self.ab_t)

    def denoise_add_noise(self, x, t, pred_noise, z=None):
        return denoise_add_noise(x, t, pred_noise, z, self.b_t, self.a_t, self.ab_t)

    @torch.no_grad()
    def sample_ddpm(self, n_sample=1, size=64, device=None, params=None, guide_w=0.0, **kw):
        """train_diffusion_paper.py:555-623 -> (x, intermediate[n_snap,B,1,64,64], sampling_time, timestep_times)."""
        t0 = time.time()
        x = torch.randn(n_sample, 1, size, size)  # :578
        if params is None:
            params = torch.rand(n_sample, self.n_cfeat)  # :580-582
        x, inter, dt = _sample(self.nn_model, x.to(self.device), params.to(self.device), guide_w, self.timesteps,
                               self.sched, **kw)
        return x, inter, time.time() - t0, [dt / self.timesteps] * self.timesteps

    @torch.no_grad()
    def sample_ddpm_from_noise(self, noise_images, params=None, save_rate=20, guide_w=0.0, **kw):
        """train_diffusion_paper.py:625-686: params=None -> unconditional (c=None -> zeros), no CFG."""
        t0 = time.time()
        x, inter, dt = _sample(self.nn_model, noise_images.clone().to(self.device), params, guide_w, self.timesteps,
                               self.sched, save_rate=save_rate, **kw)
        return x, inter, time.time() - t0, [dt / self.timesteps] * self.timesteps

    @torch.no_grad()
    def open_sampler(self, noise_images, params=None, guide_w=0.0, save_rate=20, shortcut_tab=None, seed=None):
        """Step-wise form of `sample_ddpm_from_noise` for callers that drive the loop themselves (progress bars,
        early stopping, feeding their own noise): returns a `SamplerSession`; `session.step(z)` advances one
        reverse-diffusion step (z: the step's noise [B,1,H,W] as a pinned host or device tensor; callers that do not
        feed their own noise use `sample_ddpm` / `sample_ddpm_from_noise`, which draw it in-kernel),
        `session.result()` returns (x, intermediate) on the host."""
        return SamplerSession(self, noise_images, params, guide_w, save_rate, shortcut_tab, seed)

    def calculate_likelihood(self, dataloader, **kw):
        return calculate_likelihood(self.nn_model, dataloader, self.timesteps, self.device, self.ab_t, self.b_t,
                                    self.a_t, **kw)

    def calculate_elbo_and_bpd(self, dataloader, **kw):
        return calculate_elbo_and_bpd(self.nn_model, dataloader, self.timesteps, self.device, self.ab_t, self.b_t,
                                      self.a_t, **kw)


class SamplerSession:
    """One sampling run driven step by step through the captured CUDA graph (see DDPM.open_sampler)."""

    def __init__(self, ddpm, noise_images, params, guide_w, save_rate, shortcut_tab, seed):
        dev = self.dev = ddpm.device
        x_T = noise_images.to(dev, non_blocking=True)
        prm = None if params is None else params.to(dev, non_blocking=True)
        self.run = _SamplerRun(ddpm.nn_model, x_T, prm, guide_w, ddpm.timesteps, ddpm.sched, shortcut_tab=shortcut_tab,
                               save_rate=save_rate, seed=seed, snapshots=True)
        self.run.z = torch.empty(self.run.B * ddpm.nn_model.h * ddpm.nn_model.h, device=dev)  # this step's noise
        self.run.z_stride = 0
        self.run.capture()
        self.steps_done = 0
        # double-buffered noise upload (step(z, z_next=...)): z_next travels host -> device on a side stream
        # while the step's graph runs
        self._stage = torch.empty_like(self.run.z)
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._staged = torch.cuda.Event()       # the staged noise has arrived
        self._stage_free = torch.cuda.Event()   # the staged noise has been consumed (stage may be overwritten)
        self._stage_free.record()
        self._staged_key, self._staged_ref = None, None
        # per-step device -> host read of the step counter: copied into pinned memory behind each step's kernels and
        # collected ONE step later, so the read never drains the GPU's queue (the next graph is already enqueued)
        self._ctr_host = torch.zeros(2, dtype=torch.int32).pin_memory()
        self._ctr_ev = [torch.cuda.Event(), torch.cuda.Event()]
        # results leave through pinned host memory: each snapshot is copied out on the side stream as soon as the step
        # that takes it has run (the reference copies `x` to numpy at every 20th step, paper.py:617-618), the final x at
        # result().  If the host cannot pin that much memory the copies fall back to pageable memory at result().
        self._snap_slot_of = {i: k for k, i in enumerate(self.run.snap_steps)}
        try:
            self._x_host = torch.empty(self.run.x.shape, dtype=torch.float32).pin_memory()
            self._snap_host = (torch.empty(self.run.snap.shape, dtype=torch.float32).pin_memory()
                               if self.run.snap is not None else None)
        except RuntimeError:
            self._x_host, self._snap_host = None, None
        self._snap_done = torch.cuda.Event()

    @staticmethod
    def _key(t):
        return (t.data_ptr(), t.numel(), t.device)

    @L.on_device
    def step(self, z, z_next=None, sync=False):
        """One reverse-diffusion step with the caller's noise `z` [B,1,H,W] (host tensors are copied asynchronously;
        pin them to overlap the copy).  `z_next`, if given, is the NEXT step's noise: its host-to-device copy runs
        on a side stream underneath this step's kernels, and the next `step(z_next, ...)` finds it on the device
        (keep the tensor alive and unchanged until then).  Every step also reads the device step counter back (4 bytes,
        pinned host memory); the call returns the value read back by the PREVIOUS step (the current one is still in
        flight — waiting for it would drain the GPU's queue between steps); `sync=True` waits for this step's own value.
        A snapshot taken by this step starts its device-to-host copy (pinned memory, side stream) right away.
        `result()` synchronises."""
        if z is None or z.numel() != self.run.z.numel():
            raise L.CdmError(f"step(z): z must hold this step's noise, {self.run.z.numel()} values "
                             f"([B,1,H,W] = [{self.run.B},1,{self.run.model.h},{self.run.model.h}])")
        if self.steps_done >= self.run.T:
            raise L.CdmError(f"all {self.run.T} steps of this trajectory are done: open a new sampler session")
        main = torch.cuda.current_stream()
        if self._staged_key is not None and self._staged_key == self._key(z):
            main.wait_event(self._staged)
            self.run.z.copy_(self._stage, non_blocking=True)
            self._stage_free.record(main)
        else:
            self.run.z.copy_(z.reshape(-1), non_blocking=True)
        self._staged_key, self._staged_ref = None, None
        if self.run.graph is not None:
            self.run.graph.replay()
        else:
            self.run._one_step()
        if z_next is not None:
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(self._stage_free)
                self._stage.copy_(z_next.reshape(-1), non_blocking=True)
                self._staged.record(self._copy_stream)
            self._staged_key, self._staged_ref = self._key(z_next), z_next
        slot = self._snap_slot_of.get(self.run.T - self.steps_done)  # this step ran at i = T - steps_done
        if slot is not None and self._snap_host is not None:
            taken = torch.cuda.Event()
            taken.record(main)
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(taken)
                self._snap_host[slot].copy_(self.run.snap[slot], non_blocking=True)
                self._snap_done.record(self._copy_stream)
        k = self.steps_done & 1
        self._ctr_host[k:k + 1].copy_(self.run.step, non_blocking=True)
        self._ctr_ev[k].record(main)
        self.steps_done += 1
        if sync or self.steps_done == 1:
            self._ctr_ev[k].synchronize()
            return int(self._ctr_host[k])
        self._ctr_ev[1 - k].synchronize()  # the previous step's read-back: complete unless the host runs ahead
        return int(self._ctr_host[1 - k])

    def result(self):
        """(x [B,1,H,W], intermediate [n_snapshots_taken,B,1,H,W]) on the host, like the reference's return values.
        Both are views of the session's pinned result buffers: copy them if the session goes on stepping (later
        snapshots fill later rows; a later result() call overwrites x)."""
        n_snap = sum(1 for i in self.run.snap_steps if i > self.run.T - self.steps_done)
        if self._x_host is None:
            return self.run.x.cpu(), self.run.snap[:n_snap].cpu().numpy()
        self._x_host.copy_(self.run.x, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self._snap_done.synchronize()
        inter = self._snap_host[:n_snap].numpy() if self._snap_host is not None else \
            np.zeros((0,) + tuple(self.run.x.shape), np.float32)
        return self._x_host, inter


# --------------------------------------------------------------------------- likelihood / ELBO
class _EvalLoop:
    """perturb -> U-Net -> per-sample MSE accumulate for one (x, param) batch at a device-resident step t."""

    def __init__(self, model, *a, **k):
        self.dev = model._check_supported()
        self._init(model, *a, **k)

    @L.on_device
    def _init(self, model, x, param, timesteps, sched, cb_kind, weight_tab, *, shortcut_tab=None, seed=0,
              weight_tab2=None, sample_offset=0):
        dev = model._check_supported()
        if model.training:
            raise L.CdmError("likelihood / ELBO evaluation needs eval mode: call model.eval()")
        self.model, self.T, self.dev = model, timesteps, dev
        b_t, a_t, ab_t = (s.to(dev, torch.float32) for s in sched)
        self.B = x.shape[0]
        self.x = x.detach().to(dev, torch.float32).reshape(self.B, 1, model.h, model.h).contiguous()
        c = torch.zeros(self.B, model.n_cfeat, device=dev) if param is None else param.to(dev, torch.float32)
        self.cemb1, self.cemb2 = model.contextembed1(c.contiguous()), model.contextembed2(c.contiguous())
        self.temb1, self.temb2 = _time_table(model, timesteps, dev)
        self.ca = ab_t.sqrt().contiguous()
        self.cb = ((1 - ab_t) if cb_kind == "one_minus" else torch.sqrt(1 - ab_t)).contiguous()
        self.weight = weight_tab.to(dev, torch.float32).contiguous()
        if shortcut_tab is None:
            shortcut_tab = torch.zeros(timesteps + 1, 1, 2, model.n_feat)
            for i in range(1, timesteps + 1):
                shortcut_tab[i, 0, 0].uniform_(-1, 1)
                shortcut_tab[i, 0, 1].uniform_(-1, 1)
        self.sc_tab = shortcut_tab.to(dev, torch.float32).contiguous()
        self.step = torch.zeros(1, device=dev, dtype=torch.int32)
        self.xt = torch.empty_like(self.x)
        self.noise = torch.empty_like(self.x)
        self.acc = torch.zeros(self.B, device=dev)
        self.mse = torch.zeros(self.B, device=dev)
        # optional second weighted sum over the SAME forwards (BASELINE config 5: NLL and ELBO weights in one sweep)
        self.weight2 = None if weight_tab2 is None else weight_tab2.to(dev, torch.float32).contiguous()
        self.acc2 = None if weight_tab2 is None else torch.zeros(self.B, device=dev)
        self.seed = seed
        self.sample_offset = int(sample_offset)  # global index of x[0]: in-kernel noise is keyed by the global sample
        self.graph = None

    @L.on_device
    def one(self, noise=None):
        m = self.model
        if noise is not None:
            self.noise.copy_(noise.reshape(self.noise.shape))
            L.perturb(self.x, self.xt, self.ca, self.cb, noise=self.noise, step_ptr=self.step)
        else:
            L.perturb(self.x, self.xt, self.ca, self.cb, step_ptr=self.step, seed=self.seed, noise_out=self.noise,
                      sample_offset=self.sample_offset)
        eps = m.forward_eval_into(self.xt.view(self.B, m.h, m.h), self.sc_tab, self.cemb1, self.temb1, self.cemb2,
                                  self.temb2, 1, reps=1, step_ptr=self.step)
        L.mse_accum(eps, self.noise, weight_tab=self.weight, step_ptr=self.step, mse_out=self.mse, acc=self.acc,
                    weight_tab2=self.weight2, acc2=self.acc2)

    @L.on_device
    def sweep_all(self):
        """t = 1..T with in-kernel noise, one captured graph replayed T times."""
        self.step.fill_(1)
        self.one()
        torch.cuda.synchronize()
        self.acc.zero_()
        if self.acc2 is not None:
            self.acc2.zero_()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.one()
            L.step_advance(self.step, 1)
        self.acc.zero_()
        if self.acc2 is not None:
            self.acc2.zero_()
        self.step.fill_(1)
        for _ in range(self.T):
            g.replay()
        return self.acc


@torch.no_grad()
def calculate_likelihood(model, dataloader, timesteps, device, ab_t, b_t, a_t, *, noises=None, shortcuts=None,
                         seed=0, sample_offset=0, with_elbo=False):
    """Mean over the dataset of sum_{t=1..T} mse_t / (2 b_t)  (train_diffusion_paper.py:142-183).
    noises / shortcuts: optional per-batch lists of recorded draws (tests); default: in-kernel Philox noise
    and shortcut draws from the global CPU generator.  sample_offset: global index of the loader's first map when
    the dataset is sharded over ranks (the in-kernel noise is keyed by the global map index, so two ranks never
    share noise).  with_elbo=True (BASELINE config 5, train_diffusion_elbo.py:91-103 over all timesteps): the same
    forwards also accumulate 0.5 (1/(1-ab_t) - 1) mse and the call returns (nll, elbo, bpd)."""
    model.eval()
    total_nll, total_elbo, num_samples = 0.0, 0.0, 0
    w = 1.0 / (2 * b_t.float())
    w2 = 0.5 * (1.0 / (1.0 - ab_t.float()) - 1.0) if with_elbo else None
    if w2 is not None:
        w2[0] = 0.0  # ab_t[0] = 1: t = 0 is never evaluated
    for bi, (x, param) in enumerate(dataloader):
        sc = None
        if shortcuts is not None:
            sc = shortcut_table_from_list(shortcuts[bi], timesteps, 1, order="ascending")
        loop = _EvalLoop(model, x, param, timesteps, (b_t, a_t, ab_t), "one_minus", w, shortcut_tab=sc,
                         seed=seed + 7919 * bi, weight_tab2=w2, sample_offset=sample_offset + num_samples)
        if noises is None:
            acc = loop.sweep_all()
        else:
            for t in range(1, timesteps + 1):
                loop.step.fill_(t)
                loop.one(noises[bi][t - 1].to(loop.dev))
            acc = loop.acc
        total_nll += acc.sum().item()
        if with_elbo:
            total_elbo += loop.acc2.sum().item()
        num_samples += x.shape[0]
    if with_elbo:
        elbo = total_elbo / num_samples
        return total_nll / num_samples, elbo, elbo / (64 * 64 * np.log(2))
    return total_nll / num_samples


def calculate_likelihood_and_elbo(model, dataloader, timesteps, device, ab_t, b_t, a_t, **kw):
    """BASELINE config 5 in one sweep: (NLL, ELBO, BPD) over all `timesteps` per map, one forward per (map, t)."""
    return calculate_likelihood(model, dataloader, timesteps, device, ab_t, b_t, a_t, with_elbo=True, **kw)


@torch.no_grad()
def calculate_elbo_and_bpd(model, dataloader, timesteps, device, ab_t, b_t, a_t, *, noises=None, shortcuts=None,
                           seed=0, sample_offset=0):
    """Dataloader ELBO/BPD of train_diffusion_paper.py:77-139: 10 timesteps linspace(1,T,10).long(),
    x_t = sqrt(ab) x + sqrt(1-ab) noise, weight 0.5 b_t/(1-ab_t), t<=1 skipped, /10."""
    model.eval()
    total, num = 0.0, 0
    w = 0.5 * (b_t.float() / (1.0 - ab_t.float())) / 10.0
    w = torch.where(torch.arange(timesteps + 1, device=w.device) > 1, w, torch.zeros_like(w))
    ts = [int(v) for v in torch.linspace(1, timesteps, 10).long()]
    for bi, (x, param) in enumerate(dataloader):
        sc = None
        if shortcuts is not None:
            sc = torch.zeros(timesteps + 1, 1, 2, model.n_feat)
            for k, t in enumerate(ts):  # later duplicates of t (tiny T) overwrite: replay needs distinct t
                sc[t, 0, 0], sc[t, 0, 1] = shortcuts[bi][k][0].view(-1), shortcuts[bi][k][1].view(-1)
        loop = _EvalLoop(model, x, param, timesteps, (b_t, a_t, ab_t), "sqrt", w, shortcut_tab=sc,
                         seed=seed + 104729 * bi, sample_offset=sample_offset + num)
        for k, t in enumerate(ts):
            loop.step.fill_(t)
            loop.seed = seed + 104729 * bi + 31 * k
            loop.one(None if noises is None else noises[bi][k].to(loop.dev))
        total += loop.acc.sum().item()
        num += x.shape[0]
    avg = total / num
    return avg, avg / (64 * 64 * np.log(2))


@torch.no_grad()
def calculate_elbo_and_bpd_batch(x, pred_noise, noise, t, b_t, a_t, ab_t, dims):
    """Per-batch ELBO/BPD of train_diffusion_elbo.py:74-105 -> (0-d tensor, 0-d tensor)."""
    dev = ab_t.device
    pred = pred_noise.detach().to(dev, torch.float32).contiguous()
    tgt = noise.detach().to(dev, torch.float32).contiguous()
    B = pred.shape[0]
    w = (0.5 * (1.0 / (1.0 - ab_t.float()) - 1.0)).contiguous()
    acc = torch.zeros(B, device=dev)
    L.mse_accum(pred, tgt, weight_tab=w, t_idx=t.to(dev, torch.int64).contiguous(), acc=acc)
    elbo = acc.mean()
    return elbo, elbo / (dims * np.log(2))
