"""Drop-in ContextUnet for the reference's `ContextUnet.py` / `diffusion_utilities.py`.

Same constructor, attribute names, parameter shapes and `state_dict` keys as the
reference (ContextUnet.py:5-40, code/diffusion_utilities.py:13-145), so weights
move both ways with `load_state_dict`.  The nn.Conv2d / BatchNorm2d / ... children
are parameter containers only: `forward` never calls them.  All arithmetic runs in
the sm_100a kernels behind include/cdm_b200.h; on a CPU tensor / non-sm_100 device
`forward` raises (there is no fallback path).
"""
import os

import torch
import torch.nn as nn

from . import _lib as L

BN_EPS = 1e-5
GN_EPS = 1e-5


class ResidualConvBlock(nn.Module):
    """Parameter layout of diffusion_utilities.py:13-37 (Conv3x3-BN-ReLU twice)."""

    def __init__(self, in_channels, out_channels, is_res=False):
        super().__init__()
        self.same_channels = in_channels == out_channels
        self.is_res = is_res
        self.conv1 = nn.Sequential(nn.Conv2d(in_channels, out_channels, 3, 1, 1), nn.BatchNorm2d(out_channels),
                                   nn.ReLU())
        self.conv2 = nn.Sequential(nn.Conv2d(out_channels, out_channels, 3, 1, 1), nn.BatchNorm2d(out_channels),
                                   nn.ReLU())

    def get_out_channels(self):
        return self.conv2[0].out_channels

    def _run_nhwc(self, a, pool=False, shortcut=None):
        """Eval-mode block on an NHWC bf16 tensor (or fp32 [n,H,W] if in_channels == 1)."""
        if self.training:
            raise L.CdmError("stand-alone blocks run in eval mode; training goes through ContextUnet.forward")
        dev = self.conv1[0].weight.device
        cout = self.conv2[0].out_channels
        s1, b1 = _fold_bn(self.conv1[0], self.conv1[1])
        s2, b2 = _fold_bn(self.conv2[0], self.conv2[1])
        if self.conv1[0].in_channels == 1:
            n, H, W = a.shape
            y1 = torch.empty(n, H, W, cout, device=dev, dtype=torch.bfloat16)
            L.conv_in(a, self.conv1[0].weight.detach().float().reshape(cout, 9).t().contiguous(), s1, b1, y1)
        else:
            n, H, W, _ = a.shape
            y1 = torch.empty(n, H, W, cout, device=dev, dtype=torch.bfloat16)
            L.conv3x3(a, _pack_conv3(self.conv1[0]), s1, b1, y1)
        Ho, Wo = (H // 2, W // 2) if pool else (H, W)
        y2 = torch.empty(n, Ho, Wo, cout, device=dev, dtype=torch.bfloat16)
        kw = {}
        flags = L.EPI_RELU | (L.EPI_POOL if pool else 0)
        if shortcut is not None:  # is_res with in_channels != out_channels: the fresh random 1x1 conv (G1)
            flags |= L.EPI_SHORTCUT
            kw = dict(sc_x=a, sc_tab=shortcut.to(dev, torch.float32).reshape(1, 1, 2, cout).contiguous(), sc_reps=1)
        L.conv3x3(y1, _pack_conv3(self.conv2[0]), s2, b2, y2, flags=flags, **kw)
        return y2

    def forward(self, x, shortcut=None):
        """diffusion_utilities.py:39-65, eval mode, NCHW fp32 in / out."""
        dev = self.conv1[0].weight.device
        if dev.type != "cuda":
            raise L.CdmError("no CPU path: move the block to an sm_100 device")
        x = x.detach().to(dev, torch.float32)
        cin, cout = self.conv1[0].in_channels, self.conv2[0].out_channels
        a = x[:, 0].contiguous() if cin == 1 else x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        sc = None
        if self.is_res and not self.same_channels:
            if cin != 1:
                raise L.CdmError("the fresh 1x1 shortcut is only built for in_channels == 1 (init_conv)")
            if shortcut is None:
                conv = nn.Conv2d(cin, cout, kernel_size=1)  # consumes the global CPU generator like the reference
                shortcut = torch.cat([conv.weight.detach().view(-1), conv.bias.detach().view(-1)])
            sc = shortcut
        y = self._run_nhwc(a, shortcut=sc)
        if self.is_res and self.same_channels:
            L.add_bf16(y, cout, a, cout, y.shape[0] * y.shape[1] * y.shape[2], cout)
        return y.float().permute(0, 3, 1, 2).contiguous()


class UnetUp(nn.Module):
    """diffusion_utilities.py:79-92."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.model = nn.Sequential(nn.ConvTranspose2d(in_channels, out_channels, 2, 2),
                                   ResidualConvBlock(out_channels, out_channels),
                                   ResidualConvBlock(out_channels, out_channels))

    def forward(self, x, skip):
        """cat(x, skip) -> ConvTranspose2d(2,2) -> RCB -> RCB (diffusion_utilities.py:94-100), eval, NCHW fp32."""
        ct = self.model[0]
        dev = ct.weight.device
        if dev.type != "cuda":
            raise L.CdmError("no CPU path: move the block to an sm_100 device")
        n, c0, H, W = x.shape
        a0 = x.detach().to(dev, torch.float32).permute(0, 2, 3, 1).reshape(n * H * W, c0).contiguous().to(torch.bfloat16)
        a1 = skip.detach().to(dev, torch.float32).permute(0, 2, 3, 1).reshape(n * H * W, -1).contiguous().to(torch.bfloat16)
        cout = ct.out_channels
        if cout != 128 or (H & (H - 1)) or (W & (W - 1)):
            raise L.CdmError("UnetUp kernel path is specialised for 128 output channels and power-of-two maps")
        v = torch.empty(n, 2 * H, 2 * W, cout, device=dev, dtype=torch.bfloat16)
        L.gemm(a0, _pack_convT(ct), ct.bias.detach().float().contiguous(), v, a1=a1, out_mode=1, H=H, W=W,
               shift_mod=cout)
        y = self.model[2]._run_nhwc(self.model[1]._run_nhwc(v))
        return y.float().permute(0, 3, 1, 2).contiguous()


class UnetDown(nn.Module):
    """diffusion_utilities.py:103-112."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.model = nn.Sequential(ResidualConvBlock(in_channels, out_channels),
                                   ResidualConvBlock(out_channels, out_channels), nn.MaxPool2d(2))

    def forward(self, x):
        """RCB -> RCB -> MaxPool2d(2) (diffusion_utilities.py:114-116), eval mode, NCHW fp32 in / out."""
        dev = self.model[0].conv1[0].weight.device
        if dev.type != "cuda":
            raise L.CdmError("no CPU path: move the block to an sm_100 device")
        a = x.detach().to(dev, torch.float32).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        y = self.model[1]._run_nhwc(self.model[0]._run_nhwc(a), pool=True)
        return y.float().permute(0, 3, 1, 2).contiguous()


class EmbedFC(nn.Module):
    """diffusion_utilities.py:118-145."""

    def __init__(self, input_dim, emb_dim):
        super().__init__()
        self.input_dim = input_dim
        self.model = nn.Sequential(nn.Linear(input_dim, emb_dim), nn.GELU(), nn.Linear(emb_dim, emb_dim))

    def forward(self, x):
        p = self.model[0].weight
        x = x.to(p.device).reshape(-1, self.input_dim).float().contiguous()
        out = torch.empty(x.shape[0], self.model[2].weight.shape[0], device=p.device, dtype=torch.float32)
        return L.embed_fc(x, self.model[0].weight.detach().contiguous(), self.model[0].bias.detach().contiguous(),
                          self.model[2].weight.detach().contiguous(), self.model[2].bias.detach().contiguous(), out)


def _fold_bn(conv, bn):
    """Eval-mode BatchNorm folded into a per-channel scale/shift of the conv accumulator."""
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() + (conv.bias.detach().float() - bn.running_mean.detach().float()) * scale
    return scale.contiguous(), shift.contiguous()


def _pack_conv3(conv):
    """OIHW fp32 -> [cout][kh][kw][cin] bf16 (K-major rows for the implicit GEMM)."""
    return conv.weight.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _pack_convT(convt):
    """ConvTranspose2d IOHW -> GEMM B operand [(kh*k+kw)*cout + co][ci] bf16."""
    w = convt.weight.detach()
    ci, co, kh, kw = w.shape
    return w.permute(2, 3, 1, 0).reshape(kh * kw * co, ci).contiguous().to(torch.bfloat16)


class _Workspace:
    """Named views of a plan's activation workspace (NHWC bf16) for `n` = reps * batch images in flight."""

    def __init__(self, plan, nf, h):
        n, batch = plan.n, plan.batch
        bf, f32 = torch.bfloat16, torch.float32
        h2, h4 = h // 2, h // 4
        self.n, self.batch, self.reps, self.plan = n, batch, plan.reps, plan
        self.x0 = plan.buffer("x0", bf, (n, h, h, nf))
        self.p64 = plan.buffer("p64", bf, (n, h, h, nf))
        self.q64 = plan.buffer("q64", bf, (n, h, h, nf))
        self.a1 = self.q64[:batch]  # init_conv.conv1 output, dead before q64 is first written
        self.d1 = plan.buffer("d1", bf, (n, h2, h2, nf))
        self.p32w = plan.buffer("p32w", bf, (n, h2, h2, 2 * nf))
        self.q32w = plan.buffer("q32w", bf, (n, h2, h2, 2 * nf))
        # the nf-wide h/2 buffers of up1 alias the (dead by then) 2nf-wide ones of down2
        self.p32 = self.p32w.view(-1)[: n * h2 * h2 * nf].view(n, h2, h2, nf)
        self.q32 = self.q32w.view(-1)[: n * h2 * h2 * nf].view(n, h2, h2, nf)
        self.u1f = plan.buffer("u1f", bf, (n, h2, h2, nf))
        self.d2 = plan.buffer("d2", bf, (n, h4, h4, 2 * nf))
        self.hidden = plan.buffer("hidden", bf, (n, 2 * nf))
        self.u0raw = plan.buffer("u0raw", bf, (n, h4 * h4, 2 * nf))
        self.u0f = plan.buffer("u0f", bf, (n, h4 * h4, 2 * nf))
        self.gn_partial = plan.buffer("gn_partial", f32, (n, (h // 16) * (h // 16) * 8, 8, 2))
        self.gn_mr = plan.buffer("gn_mr", f32, (n, 8, 2))
        self.eps = plan.buffer("eps", f32, (n, 1, h, h))


class ContextUnet(nn.Module):
    """ContextUnet(in_channels=1, n_feat=128, n_cfeat, height=64) — ContextUnet.py:5-60."""

    def __init__(self, in_channels, n_feat=128, n_cfeat=10, height=64):
        super().__init__()
        self.in_channels = in_channels
        self.n_feat = n_feat
        self.n_cfeat = n_cfeat
        self.h = height
        # construction order == reference (ContextUnet.py:14-40): seeded init gives identical weights
        self.init_conv = ResidualConvBlock(in_channels, n_feat, is_res=True)
        self.down1 = UnetDown(n_feat, n_feat)
        self.down2 = UnetDown(n_feat, 2 * n_feat)
        self.to_vec = nn.Sequential(nn.AvgPool2d((self.h // 4)), nn.GELU())
        self.timeembed1 = EmbedFC(1, 2 * n_feat)
        self.timeembed2 = EmbedFC(1, n_feat)
        self.contextembed1 = EmbedFC(n_cfeat, 2 * n_feat)
        self.contextembed2 = EmbedFC(n_cfeat, n_feat)
        self.up0 = nn.Sequential(nn.ConvTranspose2d(2 * n_feat, 2 * n_feat, self.h // 4, self.h // 4),
                                 nn.GroupNorm(8, 2 * n_feat), nn.ReLU())
        self.up1 = UnetUp(4 * n_feat, n_feat)
        self.up2 = UnetUp(2 * n_feat, n_feat)
        self.out = nn.Sequential(nn.Conv2d(2 * n_feat, n_feat, 3, 1, 1), nn.GroupNorm(8, n_feat), nn.ReLU(),
                                 nn.Conv2d(n_feat, self.in_channels, 3, 1, 1))
        self._plans = {}   # (batch, reps) -> (_lib.Plan, _Workspace, pointer key, version key)
        self._dirty = False
        self.conv_mode = int(os.environ.get("CDM_CONV_MODE", L.CONV_MODE_SWAPPED_TMA))

    # ------------------------------------------------------------------ weights
    def _check_supported(self):
        if self.in_channels != 1 or self.n_feat != 128 or self.h != 64:
            raise L.CdmError("the sm_100a kernels are specialised for in_channels=1, n_feat=128, height=64 "
                             "(the configuration BASELINE.json names)")
        dev = self.out[3].weight.device
        if dev.type != "cuda":
            raise L.CdmError("ContextUnet parameters are on the CPU: the hot path has no CPU fallback; "
                             "move the module to an sm_100 device with .to('cuda')")
        return dev

    def _plan_tensors(self):
        sd = dict(self.named_parameters())
        sd.update(dict(self.named_buffers()))
        return [sd[k].detach() for k in L.plan_tensor_names()]

    def invalidate(self):
        """Parameters / buffers were changed behind autograd's back (raw-pointer kernels, .data writes): re-pack the
        bf16 weights and folded BatchNorm vectors on the next eval forward."""
        self._dirty = True

    def plan(self, batch, reps):
        """The cdm_plan (packed weights, workspace, prepared launches) for `reps` passes over `batch` inputs: built on
        first use, re-packed when a parameter / buffer changed (version counters, or invalidate()), rebuilt when the
        tensors moved (other device, load_state_dict(assign=True))."""
        ts = self._plan_tensors()
        pkey = tuple((t.data_ptr(), str(t.device)) for t in ts)
        vkey = tuple(t._version for t in ts)
        if self._dirty:  # applies to every cached plan
            for k, (pl, ws, pk, vk) in list(self._plans.items()):
                self._plans[k] = (pl, ws, pk, None)
            self._dirty = False
        ent = self._plans.get((batch, reps))
        if ent is not None and ent[2] == pkey:
            if ent[3] != vkey:
                ent[0].refresh()
                self._plans[(batch, reps)] = (ent[0], ent[1], pkey, vkey)
            return self._plans[(batch, reps)]
        if ent is not None:
            ent[0].close()
        if len(self._plans) > 4:
            for pl, *_ in self._plans.values():
                pl.close()
            self._plans.clear()
        dev = self._check_supported()
        with torch.cuda.device(dev):
            pl = L.Plan([t if (t.is_contiguous() and t.dtype == torch.float32) else t.float().contiguous() for t in ts],
                        self.n_cfeat, batch, reps, dev, conv_mode=self.conv_mode)
        ws = _Workspace(pl, self.n_feat, self.h)
        self._plans[(batch, reps)] = (pl, ws, pkey, vkey)
        return self._plans[(batch, reps)]

    def workspace(self, batch, reps):
        return self.plan(batch, reps)[1]

    # ------------------------------------------------------------------ embeddings
    def embed(self, t, c):
        """The four EmbedFC outputs (ContextUnet.py:51-54): cemb1[B,2nf], temb1[nt,2nf], cemb2[B,nf], temb2[nt,nf]."""
        return self.contextembed1(c), self.timeembed1(t), self.contextembed2(c), self.timeembed2(t)

    # ------------------------------------------------------------------ eval forward
    def forward_eval_into(self, x, sc_tab, cemb1, temb1, cemb2, temb2, temb_rows, *, reps=1, step_ptr=None):
        """Eval-mode forward of `reps` passes that share x (reps=2: the conditional and unconditional
        classifier-free-guidance passes).  x fp32 [B,64,64]; sc_tab fp32 [steps][reps][2][128];
        cemb* fp32 [reps*B,C]; temb* fp32 [steps][temb_rows][C] (row picked by *step_ptr).
        Returns the workspace's eps buffer, fp32 [reps*B,1,64,64] (overwritten by the next call)."""
        B = x.shape[0]
        pl, ws = self.plan(B, reps)[:2]
        # ONE C call: 26 launches over the plan's prepared tensor maps (cdm_forward_eval, csrc/plan.cu)
        pl.forward_eval(x, sc_tab, cemb1, temb1, cemb2, temb2, temb_rows, step_ptr=step_ptr)
        return ws.eps

    def draw_shortcut(self):
        """The reference builds `nn.Conv2d(C_in, n_feat, 1)` afresh on every forward
        (diffusion_utilities.py:54): random, unregistered, drawn from the global CPU generator.
        Constructing the same layer consumes the generator identically."""
        sc = nn.Conv2d(self.in_channels, self.n_feat, kernel_size=1, stride=1, padding=0)
        return torch.cat([sc.weight.detach().view(-1), sc.bias.detach().view(-1)])  # [2*n_feat]: w_c then b_c

    @L.on_device
    def forward(self, x, t, c=None, shortcut=None):
        """ContextUnet.forward (ContextUnet.py:42-60).  x [B,1,64,64]; t numel 1 or B; c [B,n_cfeat] or None.
        `shortcut` ([2*n_feat] = w_c,b_c) overrides the per-call random 1x1 shortcut (tests / replay)."""
        dev = self._check_supported()
        if self.training:
            from .train import forward_train
            return forward_train(self, x, t, c, shortcut)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and x.requires_grad:
            raise L.CdmError("eval-mode forward does not build an autograd graph; call .train() for training")
        B = x.shape[0]
        x3 = x.detach().to(dev, torch.float32).reshape(B, self.h, self.h).contiguous()
        if c is None:
            c = torch.zeros(B, self.n_cfeat, device=dev)  # ContextUnet.py:48-49
        t = torch.as_tensor(t).to(dev, torch.float32).reshape(-1, 1)
        if t.shape[0] not in (1, B):
            raise L.CdmError(f"t must have 1 or {B} elements, got {t.shape[0]}")
        cemb1, temb1, cemb2, temb2 = self.embed(t, c.to(dev, torch.float32))
        sc = self.draw_shortcut() if shortcut is None else shortcut
        sc_tab = sc.detach().to(dev, torch.float32).reshape(1, 1, 2, self.n_feat).contiguous()
        eps = self.forward_eval_into(x3, sc_tab, cemb1, temb1, cemb2, temb2, t.shape[0], reps=1)
        return eps.clone()

    def train(self, mode=True):
        self.invalidate()  # a mode switch brackets raw-pointer updates (optimizer, BatchNorm running statistics)
        return super().train(mode)
