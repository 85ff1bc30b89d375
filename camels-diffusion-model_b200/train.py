"""Training path of ContextUnet: train-mode BatchNorm forward, full backward and a fused Adam, all through
the sm_100a kernels behind include/cdm_b200.h.

Mirrors the reference training step (code/train_diffusion_paper.py:349-366):
    pred = nn_model(x_pert, t / timesteps, param); loss = F.mse_loss(pred, noise); loss.backward(); optim.step()
`ContextUnet.forward` in `.train()` mode returns `pred` attached to autograd through `_UnetFn`, whose backward runs
the hand-written backward pass and hands every parameter its gradient, so `loss.backward()` and any torch optimizer
work unchanged; `FusedAdam` is the in-house optimizer (one multi-tensor kernel).

Data parallelism (one process per GPU): if torch.distributed is initialised, BatchNorm batch statistics and their
backward sums are all-reduced across ranks (count = global N*H*W, so results match the single-device reference at
the same global batch) and parameter gradients are all-reduced once, as one flat buffer, at the end of backward.

Tensor-core work: forward convs and data gradients (a conv with flipped / transposed weights) use cdm_conv3x3 /
cdm_gemm; every weight gradient is cdm_gemm_tn (MN-major operands straight from the NHWC activations).
"""
import torch
import torch.distributed as dist

from . import _lib as L

BN_EPS, GN_EPS, BN_MOMENTUM = 1e-5, 1e-5, 0.1
DATA_PARALLEL = True  # set False to run a rank-local step inside an initialised process group (tests)


def _world():
    if DATA_PARALLEL and dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


def _rank():
    if DATA_PARALLEL and dist.is_available() and dist.is_initialized():
        return dist.get_rank()
    return 0


def _allreduce(t):
    if _world() > 1:
        dist.all_reduce(t)
    return t


# Backward BatchNorm sums in the data-gradient convolution's epilogue (CDM_EPI_BNBWD) instead of a cdm_chan_reduce pass
# over dy and z: up to this many images per GPU.  Measured: it removes 13 launches from the latency-bound chain of the
# 32-images-per-GPU step (3.91 -> 3.85 ms), but at batch 256 the extra z loads + mask arithmetic (~200 instructions per
# 32-pixel chunk on top of ~130) make the epilogue the pacing role of the convolution and cost what the separate pass —
# which runs at the HBM roofline — would have cost (18.7 vs 18.5-18.9 ms): no gain there, so it stays off.
FUSE_BN_BWD_MAX_IMAGES = 64
PEER = None            # parallel.PeerExchange of this process, created on the first data-parallel forward
PEER_EXCHANGE = True   # False: all-reduce the BatchNorm statistics with NCCL instead (the baseline path)


def enable_peer_exchange(device=None, group=None):
    """Route the cross-rank BatchNorm statistics (forward and backward, 36 exchanges of [2C] floats per step)
    through the fused reduce + exchange kernel over NVLink peer memory instead of NCCL (7.31 -> 6.70 ms per step at
    8 x 32 images).  Collective: every rank calls it at the same point (the first data-parallel forward does, unless
    PEER_EXCHANGE is False); a no-op for a single rank."""
    global PEER
    if _world() > 1 and PEER is None:
        from .parallel import PeerExchange
        PEER = PeerExchange(torch.device("cuda", torch.cuda.current_device()) if device is None else device, group)
    return PEER


def disable_peer_exchange():
    global PEER
    PEER = None


_SIDE = {}


def _side_stream(dev, which=0):
    """Per-device auxiliary streams: 0 = weight gradients, 1 = gradient all-reduce."""
    idx = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    key = (idx, which)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


def _xr():
    global PEER_EXCHANGE
    if _world() <= 1 or not PEER_EXCHANGE:
        return None
    if PEER is None:
        err = None
        try:
            enable_peer_exchange()
        except Exception as ex:  # noqa: BLE001  (no NVLink peer access / symmetric memory on this system)
            err = ex
        # the ranks must AGREE on the path: one rank falling back to NCCL on its own would leave the others spinning
        # on its flag.  (A rank whose rendezvous itself failed cannot be rescued here; symmetric-memory rendezvous is
        # collective, so in practice it fails or succeeds everywhere — this guards the allocation / mapping steps.)
        ok = torch.tensor([0 if err is not None else 1], device=torch.cuda.current_device())
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            import warnings
            PEER_EXCHANGE = False
            disable_peer_exchange()
            warnings.warn(f"peer-memory exchange unavailable on at least one rank ({type(err).__name__ if err else 'peer'}"
                          f": {err}); BatchNorm statistics use NCCL all-reduce on EVERY rank (still on the GPUs)")
            return None
    return PEER.args


class _Ctx:
    """Saved tensors of one training forward."""
    pass


def _bf(*shape, dev):
    return torch.empty(*shape, device=dev, dtype=torch.bfloat16)


def _f32(*shape, dev, zero=False):
    return (torch.zeros if zero else torch.empty)(*shape, device=dev, dtype=torch.float32)


def _rcb_list(m):
    return [("init_conv", m.init_conv), ("down1.0", m.down1.model[0]), ("down1.1", m.down1.model[1]),
            ("down2.0", m.down2.model[0]), ("down2.1", m.down2.model[1]), ("up1.1", m.up1.model[1]),
            ("up1.2", m.up1.model[2]), ("up2.1", m.up2.model[1]), ("up2.2", m.up2.model[2])]


class _PackPlan:
    """Persistent bf16 operand buffers + the device table of cdm_pack_bf16 for one model: every tensor-core layout of
    the step (forward [co][kh][kw][ci] and data-gradient [ci][2-kh][2-kw][co] forms of the 3x3 convolutions, the
    (kh,kw,co) x ci and ci x (kh,kw,co) forms of the transposed convolutions) is refreshed by ONE launch per step."""

    def __init__(self, m):
        self.P, rows, vec = {}, [], 0
        self.key = self.key_of(m)

        def add(key, w, perm, flips=()):
            nonlocal vec
            shape = [w.shape[d] for d in perm]
            stride = [w.stride(d) for d in perm]
            off = 0
            for d in flips:  # a flipped source dim is walked backwards from its last element
                k = perm.index(d)
                off += (w.shape[d] - 1) * w.stride(d)
                stride[k] = -stride[k]
            assert w.is_contiguous() and w.dtype == torch.float32 and shape[3] % 8 == 0
            dst = torch.empty(shape, device=w.device, dtype=torch.bfloat16)
            rows.append([w.data_ptr(), dst.data_ptr(), shape[1], shape[2], shape[3], *stride, off, vec, 0])
            vec += dst.numel() // 8
            self.P[key] = dst

        for name, blk in _rcb_list(m):
            for cn, seq in (("c1", blk.conv1), ("c2", blk.conv2)):
                w = seq[0].weight.detach()
                if w.shape[1] != 1:
                    add(f"{name}.{cn}.f", w, (0, 2, 3, 1))               # [co][kh][kw][ci]
                    add(f"{name}.{cn}.d", w, (1, 2, 3, 0), flips=(2, 3))  # [ci][2-kh][2-kw][co]
        w = m.out[0].weight.detach()
        add("out0.f", w, (0, 2, 3, 1))
        add("out0.d", w, (1, 2, 3, 0), flips=(2, 3))
        self.transposes = []  # (src, dst, batches, R, C, sb, sr, db, dc): cdm_pack_transpose_bf16 launches
        for nm, mod in (("up0", m.up0[0]), ("up1", m.up1.model[0]), ("up2", m.up2.model[0])):
            w = mod.weight.detach()  # IOHW
            ci, co, kh, kw = w.shape
            khw = kh * kw
            if khw % 64 == 0 and ci % 64 == 0 and co % 64 == 0:
                # up0 (78 % of all parameters): both layouts are batched 2-D transposes of [ci][co][khw] with contiguous
                # innermost dimensions -> tiled through shared memory instead of the strided gather of the table
                assert w.is_contiguous() and w.dtype == torch.float32
                f = torch.empty(khw * co, ci, device=w.device, dtype=torch.bfloat16)   # [(khw, co)][ci]
                d = torch.empty(ci, khw * co, device=w.device, dtype=torch.bfloat16)   # [ci][(khw, co)]
                self.transposes.append((w, f, co, ci, khw, khw, co * khw, ci, co * ci))   # b = co, r = ci, c = khw
                self.transposes.append((w, d, ci, co, khw, co * khw, khw, co * khw, co))  # b = ci, r = co, c = khw
                self.P[nm + ".f"], self.P[nm + ".d"] = f, d
                continue
            add(nm + ".f", w, (2, 3, 1, 0))
            add(nm + ".d", w, (0, 2, 3, 1))
            self.P[nm + ".f"] = self.P[nm + ".f"].view(kh * kw * co, ci)
            self.P[nm + ".d"] = self.P[nm + ".d"].view(ci, kh * kw * co)
        self.n_rows, self.total_vec = len(rows), vec
        self.table = torch.tensor(rows, dtype=torch.int64).to(m.out[0].weight.device)

    @staticmethod
    def key_of(m):
        return tuple(p.data_ptr() for p in m.parameters())

    def refresh(self):
        L.pack_bf16(self.table, self.n_rows, self.total_vec)
        for t in self.transposes:
            L.pack_transpose_bf16(*t)


def _pack_train(m):
    """bf16 operand packs for this step (weights change every optimizer step): forward and data-gradient forms."""
    plan = getattr(m, "_train_pack_plan", None)
    if plan is None or plan.key != _PackPlan.key_of(m):
        plan = _PackPlan(m)
        object.__setattr__(m, "_train_pack_plan", plan)
    plan.refresh()
    P = dict(plan.P)
    blk = m.init_conv.conv1[0].weight.detach()
    P["init_conv.c1.f"] = blk.float().reshape(blk.shape[0], 9).t().contiguous()  # [9][cout] fp32 (K = 9: no tensor core)
    w3 = m.out[3].weight.detach().float()[0].permute(1, 2, 0).reshape(9, -1)  # [tap][c]
    P["out3.f"] = w3.contiguous()
    P["out3.d"] = w3.flip(0).contiguous()  # flipped taps: dgrad of a 1-output-channel conv
    return P


class _Layer:
    """One Conv3x3 + train-mode BatchNorm + ReLU, with everything its backward needs."""
    __slots__ = ("name", "conv", "bn", "x_in", "z", "y", "scale", "shift", "mean", "rstd", "count", "H", "cin", "cout")


def _conv_bn_relu(S, P, name, seq, src, n, H, *, first=False, **apply_kw):
    """z = conv(src)+bias; batch statistics (all-reduced over ranks); y = relu(bn(z)) [+ extras]."""
    dev = S.dev
    conv, bn = seq[0], seq[1]
    cout = conv.out_channels
    z = _bf(n, H, H, cout, dev=dev)
    bias = conv.bias.detach().float().contiguous()
    Pn = n * H * H
    sums = _f32(2, cout, dev=dev)
    xr = _xr()
    if not first and S.mode >= L.CONV_MODE_SWAPPED and H % 32 == 0:
        # the batch statistics come out of the convolution's epilogue (+ the cross-rank exchange): one C call
        L.conv3x3(src, P[name + ".f"], S.ones[:cout], bias, z, flags=L.EPI_BNSTATS, mode=S.mode, bn_partial=S.bnp,
                  bn_sums=sums, xr=xr)
    else:
        if first:
            L.conv_in(src, P[name + ".f"], S.ones[:cout], bias, z, relu=False)
        else:
            L.conv3x3(src, P[name + ".f"], S.ones[:cout], bias, z, flags=0, mode=S.mode)
        L.chan_reduce(z, cout, Pn, cout, sums, S.ws, mode=0, xr=xr)
    if xr is None:
        _allreduce(sums)
    count = float(Pn * _world())
    ly = _Layer()
    ly.name, ly.conv, ly.bn, ly.x_in, ly.z, ly.count, ly.H = name, conv, bn, src, z, count, H
    ly.cin, ly.cout = conv.in_channels, cout
    ly.scale, ly.shift, ly.mean, ly.rstd = (_f32(cout, dev=dev) for _ in range(4))
    S.nbt.append(bn.num_batches_tracked)  # += 1 for all 18 layers in one multi-tensor launch (end of forward)
    y = _bf(n, H, H, cout, dev=dev)
    # one launch: batch statistics -> scale / shift (+ mean, rstd, running statistics), then y = relu(z*scale+shift)
    L.bn_apply(z, Pn, cout, ly.scale, ly.shift, y, relu=1,
               finalize=(sums, bn.weight.detach(), bn.bias.detach(), count, bn.eps, bn.momentum, bn.running_mean,
                         bn.running_var, ly.mean, ly.rstd), **apply_kw)
    ly.y = y
    S.layers[name] = ly
    return y


def forward_train(model, x, t, c, shortcut=None):
    dev = model._check_supported()
    params = [p for p in model.parameters()]
    return _UnetFn.apply(model, x, t, c, shortcut, *params)


class _UnetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, t, c, shortcut, *params):
        dev = model._check_supported()
        m = model
        nf, h, n = m.n_feat, m.h, x.shape[0]
        S = _Ctx()
        S.dev, S.mode, S.layers, S.n, S.nbt = dev, m.conv_mode, {}, n, []
        S.ones = torch.ones(256, device=dev)
        S.zeros = torch.zeros(65536, device=dev)
        S.ws = _f32(L.num_sms() * 8, 9 * 256, dev=dev)  # reduction workspace (partials), main stream
        S.ws2 = _f32(L.num_sms() * 8, 9 * 256, dev=dev)  # the same for the leaf reductions on the side stream
        S.bnp = _f32(L.num_sms(), 2 * 256, dev=dev)  # per-CTA BatchNorm partial sums of the conv epilogue
        S.wgws = _f32(160 * 128 * 384, dev=dev)  # split-K tiles of the 3x3 weight gradients (used on the SIDE stream only)
        S.skws = _f32(160 * 128 * 384, dev=dev)  # split-K tiles of the main-stream GEMMs (transposed convs, up0)
        P = _pack_train(m)
        S.P = P
        x3 = x.detach().to(dev, torch.float32).reshape(n, h, h).contiguous()
        if c is None:
            c = torch.zeros(n, m.n_cfeat, device=dev)
        c = c.detach().to(dev, torch.float32).contiguous()
        tt = torch.as_tensor(t).detach().to(dev, torch.float32).reshape(-1, 1).contiguous()
        if tt.shape[0] not in (1, n):
            raise L.CdmError(f"t must have 1 or {n} elements")
        S.x, S.c, S.t = x3, c, tt
        sc = (m.draw_shortcut() if shortcut is None else shortcut).detach().to(dev, torch.float32).contiguous()
        # the four EmbedFC forwards (latency-bound: ~20 us each) are first needed by up0's FiLM, after the whole down
        # path: they run on the side stream underneath it
        main_s, side_s = torch.cuda.current_stream(), _side_stream(dev)
        fork = torch.cuda.Event()
        fork.record(main_s)
        with torch.cuda.stream(side_s):
            side_s.wait_event(fork)
            S.cemb1, S.temb1, S.cemb2, S.temb2 = m.embed(tt, c)
            emb_done = torch.cuda.Event()
            emb_done.record(side_s)
        trows = tt.shape[0]
        # ---- init_conv (ContextUnet.py:43): conv1, conv2, + fresh 1x1 shortcut
        y1 = _conv_bn_relu(S, P, "init_conv.c1", m.init_conv.conv1, x3, n, h, first=True)
        x0 = _conv_bn_relu(S, P, "init_conv.c2", m.init_conv.conv2, y1, n, h, sc_x=x3, sc_w=sc[:nf], sc_b=sc[nf:])
        # ---- down1 / down2 (RCB, RCB, MaxPool2d)
        a = x0
        for nm, blk in (("down1.0", m.down1.model[0]), ("down1.1", m.down1.model[1])):
            a = _conv_bn_relu(S, P, nm + ".c1", blk.conv1, a, n, h)
            a = _conv_bn_relu(S, P, nm + ".c2", blk.conv2, a, n, h)
        d1 = _bf(n, h // 2, h // 2, nf, dev=dev)
        L.maxpool2_fwd(a, d1)
        a = d1
        for nm, blk in (("down2.0", m.down2.model[0]), ("down2.1", m.down2.model[1])):
            a = _conv_bn_relu(S, P, nm + ".c1", blk.conv1, a, n, h // 2)
            a = _conv_bn_relu(S, P, nm + ".c2", blk.conv2, a, n, h // 2)
        h4 = h // 4
        d2 = _bf(n, h4, h4, 2 * nf, dev=dev)
        L.maxpool2_fwd(a, d2)
        # ---- to_vec, up0 (+GroupNorm, ReLU, FiLM)
        S.hid_pre = _f32(n, 2 * nf, dev=dev)
        hidden = _bf(n, 2 * nf, dev=dev)
        L.avgpool_gelu_train(d2.view(n, h4 * h4, 2 * nf), S.hid_pre, hidden)
        u0raw = _bf(n, h4 * h4, 2 * nf, dev=dev)
        L.gemm(hidden, P["up0.f"], m.up0[0].bias.detach().float().contiguous(), u0raw, shift_mod=2 * nf)
        u0f = _bf(n, h4 * h4, 2 * nf, dev=dev)
        S.gn0_mr = _f32(n, 8, 2, dev=dev)
        main_s.wait_event(emb_done)  # join: the embeddings are ready
        L.gn_relu_film(u0raw, m.up0[1].weight.detach(), m.up0[1].bias.detach(), u0f, groups=8, eps=GN_EPS,
                       film_scale=S.cemb1, film_shift=S.temb1.view(1, trows, 2 * nf), film_rows=trows,
                       mean_rstd_out=S.gn0_mr)
        # ---- up1
        v1 = _bf(n, h // 2, h // 2, nf, dev=dev)
        L.gemm(u0f.view(n * h4 * h4, 2 * nf), P["up1.f"], m.up1.model[0].bias.detach().float().contiguous(), v1,
               a1=d2.view(n * h4 * h4, 2 * nf), out_mode=1, H=h4, W=h4, shift_mod=nf)
        a = _conv_bn_relu(S, P, "up1.1.c1", m.up1.model[1].conv1, v1, n, h // 2)
        a = _conv_bn_relu(S, P, "up1.1.c2", m.up1.model[1].conv2, a, n, h // 2)
        a = _conv_bn_relu(S, P, "up1.2.c1", m.up1.model[2].conv1, a, n, h // 2)
        u1f = _bf(n, h // 2, h // 2, nf, dev=dev)
        u1 = _conv_bn_relu(S, P, "up1.2.c2", m.up1.model[2].conv2, a, n, h // 2, film_scale=S.cemb2,
                           film_shift=S.temb2, film_rows=trows, px_per_img=(h // 2) ** 2, yf=u1f)
        # ---- up2
        h2 = h // 2
        v2 = _bf(n, h, h, nf, dev=dev)
        L.gemm(u1f.view(n * h2 * h2, nf), P["up2.f"], m.up2.model[0].bias.detach().float().contiguous(), v2,
               a1=d1.view(n * h2 * h2, nf), out_mode=1, H=h2, W=h2, shift_mod=nf)
        a = _conv_bn_relu(S, P, "up2.1.c1", m.up2.model[1].conv1, v2, n, h)
        a = _conv_bn_relu(S, P, "up2.1.c2", m.up2.model[1].conv2, a, n, h)
        a = _conv_bn_relu(S, P, "up2.2.c1", m.up2.model[2].conv1, a, n, h)
        u2 = _conv_bn_relu(S, P, "up2.2.c2", m.up2.model[2].conv2, a, n, h)
        # ---- out: conv(cat(u2, x0)) -> GroupNorm -> ReLU -> conv
        o = _bf(n, h, h, nf, dev=dev)
        gnp = _f32(n, (h // 16) ** 2 * 8, 8, 2, dev=dev)
        L.conv3x3(u2, P["out0.f"], S.ones[:nf], m.out[0].bias.detach().float().contiguous(), o, src1=x0,
                  flags=L.EPI_GNSTATS, gn_partial=gnp, mode=S.mode)
        S.gn1_mr = _f32(n, 8, 2, dev=dev)
        L.gn_finalize(gnp, float((nf // 8) * h * h), S.gn1_mr, GN_EPS)
        eps = _f32(n, 1, h, h, dev=dev)
        L.conv_out(o, S.gn1_mr, m.out[1].weight.detach(), m.out[1].bias.detach(), P["out3.f"],
                   m.out[3].bias.detach().float().contiguous(), eps.view(n, h, h))
        S.x0, S.d1, S.d2, S.hidden, S.u0raw, S.u0f, S.u1, S.u1f, S.u2, S.o = x0, d1, d2, hidden, u0raw, u0f, u1, u1f, u2, o
        S.pool1_in, S.pool2_in = S.layers["down1.1.c2"].y, S.layers["down2.1.c2"].y
        torch._foreach_add_(S.nbt, 1)
        ctx.S, ctx.model = S, m
        ctx.names = [nme for nme, _ in m.named_parameters()]
        return eps

    @staticmethod
    def backward(ctx, deps):
        S, m = ctx.S, ctx.model
        dev, P, n, nf, h = S.dev, S.P, S.n, m.n_feat, m.h
        G = {}  # parameter name -> gradient (torch layout, fp32)
        bn_affine = set()
        deps = deps.detach().to(dev, torch.float32).contiguous().view(n, h, h)

        # The 3x3 weight gradients are leaves of the backward graph (nothing downstream reads them before the
        # optimizer), so they run on a second stream next to the dz -> dx chain: at 32 images per GPU every kernel of
        # the step is too small to fill the SMs and the two chains overlap (fork / join are CUDA-graph capturable).
        main, side = torch.cuda.current_stream(), _side_stream(dev)
        keep = []  # operands stay referenced until the join: the allocator must not recycle them under the side stream

        def leaf(fn):
            """Run fn() on the side stream once everything issued so far on the main stream is done.  For the LEAVES
            of the backward graph (parameter gradients: nothing downstream reads them before the join below): the
            latency-bound small kernels among them (EmbedFC backward: 16 launches of a few blocks; bias / affine
            reductions; the K = 9 and N = 1 weight gradients) then run underneath the main dz -> dx chain."""
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                out = fn()
            keep.append((fn, out))  # the closure keeps its operands referenced until the join
            return out

        def conv_wgrad(dz, srcs, cout):
            cin = sum(s.shape[3] for s in srcs)
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                dw = _f32(cout, 9 * cin, dev=dev, zero=True)
                off = 0
                Hh = dz.shape[1]
                for s in srcs:
                    L.gemm_tn(dz, s, dw.view(-1)[off:], n_img=n, H=Hh, W=Hh, a_c=cout, b_c=s.shape[3], M=cout,
                              N=s.shape[3], ldc=9 * cin, taps=9, tap_stride=cin, workspace=S.wgws)
                    off += s.shape[3]
                out = dw.view(cout, 3, 3, cin).permute(0, 3, 1, 2).contiguous()
            keep.append((dz, srcs, dw, out))
            return out

        def cbr_bwd(name, prefix, dy, lddy, need_dx=True, sums=None, fuse_next=None):
            """Backward of one Conv-BatchNorm-ReLU layer.  `sums` = (sum g, sum g*xhat) of this layer if the producer
            of `dy` has already computed them; `fuse_next` = the layer whose dy is this layer's dx (its input layer):
            the data-gradient convolution then accumulates THAT layer's sums in its epilogue (CDM_EPI_BNBWD), which
            saves the separate pass over dy and z (cdm_chan_reduce mode 1).  Returns (dx, sums of fuse_next or None)."""
            ly = S.layers[name]
            Pn, C = n * ly.H * ly.H, ly.cout
            xr = _xr()
            if sums is None:
                sums = _f32(2, C, dev=dev)
                L.chan_reduce(dy, lddy, Pn, C, sums, S.ws, mode=1, z=ly.z, ldz=C, scale=ly.scale, shift=ly.shift,
                              mean=ly.mean, rstd=ly.rstd, relu=1, xr=xr)
                if xr is None:
                    _allreduce(sums)
            G[prefix + ".1.weight"], G[prefix + ".1.bias"] = sums[1], sums[0]  # `sums` is this layer's own buffer
            bn_affine.update((prefix + ".1.weight", prefix + ".1.bias"))
            dz = _bf(n, ly.H, ly.H, C, dev=dev)
            L.bn_bwd_apply(dy, lddy, ly.z, Pn, C, ly.scale, ly.shift, ly.mean, ly.rstd, sums, ly.count, dz, relu=1)
            G[prefix + ".0.bias"] = S.zeros[:C]  # a bias in front of train-mode BN has zero gradient
            if ly.cin == 1:
                def w_first():
                    w9 = _f32(9, C, dev=dev)
                    L.outer_wgrad(ly.x_in, dz, n, ly.H, ly.H, C, w9, S.ws2, flip=0)
                    return w9.t().reshape(C, 1, 3, 3).contiguous()
                G[prefix + ".0.weight"] = leaf(w_first)
                return None, None
            G[prefix + ".0.weight"] = conv_wgrad(dz, [ly.x_in], C)
            if not need_dx:
                return None, None
            dx = _bf(n, ly.H, ly.H, ly.cin, dev=dev)
            nxt = fuse_next
            if nxt is not None and n <= FUSE_BN_BWD_MAX_IMAGES and S.mode >= L.CONV_MODE_SWAPPED and ly.H % 32 == 0 \
                    and nxt.cout == ly.cin and nxt.H == ly.H:
                sums_next = _f32(2, ly.cin, dev=dev)
                L.conv3x3(dz, P[name + ".d"], S.ones[:ly.cin], S.zeros[:ly.cin], dx, flags=L.EPI_BNBWD, mode=S.mode,
                          bn_partial=S.bnp, bn_sums=sums_next, xr=xr,
                          bwd=(nxt.z, nxt.scale, nxt.shift, nxt.mean, nxt.rstd))
                if xr is None:
                    _allreduce(sums_next)
                return dx, sums_next
            L.conv3x3(dz, P[name + ".d"], S.ones[:ly.cin], S.zeros[:ly.cin], dx, flags=0, mode=S.mode)
            return dx, None

        def rcb_bwd(name, prefix, dy, lddy, fuse_next=None):
            """ResidualConvBlock (is_res=False) backward; `fuse_next`: the Conv-BN-ReLU layer that produced this block's
            input (its backward statistics ride on conv1's data-gradient launch).  Returns (dx, sums for fuse_next)."""
            d, s1 = cbr_bwd(name + ".c2", prefix + ".conv2", dy, lddy, fuse_next=S.layers[name + ".c1"])
            return cbr_bwd(name + ".c1", prefix + ".conv1", d, d.shape[3], sums=s1, fuse_next=fuse_next)

        def convT_bwd(nm, prefix, dv, srcs):
            """ConvTranspose2d(cin, 128, 2, 2) backward: dv [n,2H,2W,128]; srcs = the two K-split inputs."""
            Hh = dv.shape[1] // 2
            Mrows = n * Hh * Hh
            s2d = _bf(Mrows, 4 * nf, dev=dev)
            L.space_to_depth(dv, s2d)
            cin = sum(s.shape[-1] for s in srcs)

            def leaves():
                sums = _f32(2, nf, dev=dev)
                L.chan_reduce(dv, nf, Mrows * 4, nf, sums, S.ws2, mode=2)
                dw = _f32(cin, 4 * nf, dev=dev, zero=True)
                off = 0
                for s in srcs:
                    cs = s.shape[-1]
                    L.gemm_tn(s, s2d, dw[off:], n_img=1, H=1, W=Mrows, a_c=cs, b_c=4 * nf, M=cs, N=4 * nf, ldc=4 * nf,
                              workspace=S.wgws)
                    off += cs
                return sums[0], dw.view(cin, 2, 2, nf).permute(0, 3, 1, 2).contiguous()
            G[prefix + ".bias"], G[prefix + ".weight"] = leaf(leaves)
            da = _bf(Mrows, cin, dev=dev)
            L.gemm(s2d, P[nm + ".d"], S.zeros[:cin], da, shift_mod=cin)
            return da

        def embed_bwd(mod, prefix, inp, dout):
            rows, emb = inp.shape[0], mod.model[2].weight.shape[0]
            w1, b1, w2 = (mod.model[0].weight.detach().contiguous(), mod.model[0].bias.detach().contiguous(),
                          mod.model[2].weight.detach().contiguous())
            pre, hh, dpre = (_f32(rows, emb, dev=dev) for _ in range(3))
            dw1, db1, dw2, db2 = _f32(emb, inp.shape[1], dev=dev), _f32(emb, dev=dev), _f32(emb, emb, dev=dev), _f32(emb, dev=dev)
            L.embed_bwd(inp, w1, b1, w2, dout.contiguous(), pre, hh, dpre, dw1, db1, dw2, db2)
            G[prefix + ".model.0.weight"], G[prefix + ".model.0.bias"] = dw1, db1
            G[prefix + ".model.2.weight"], G[prefix + ".model.2.bias"] = dw2, db2

        # ---- out.3 (Conv 128->1), out.1 (GroupNorm) + ReLU
        def out3_leaves():
            w9 = _f32(9, nf, dev=dev)
            L.outer_wgrad(deps, S.o, n, h, h, nf, w9, S.ws2, flip=1, mean_rstd=S.gn1_mr, gamma=m.out[1].weight.detach(),
                          beta=m.out[1].bias.detach())
            return deps.sum().reshape(1), w9.view(3, 3, nf).permute(2, 0, 1).reshape(1, nf, 3, 3).contiguous()
        G["out.3.bias"], G["out.3.weight"] = leaf(out3_leaves)
        d_a = _bf(n, h, h, nf, dev=dev)
        L.conv_in(deps, P["out3.d"], S.ones[:nf], S.zeros[:nf], d_a, relu=False)
        do = _bf(n, h, h, nf, dev=dev)
        dg_nc, db_nc = _f32(n, nf, dev=dev), _f32(n, nf, dev=dev)
        L.gn_bwd(S.o, d_a, nf, n, h * h, nf, 8, S.gn1_mr, m.out[1].weight.detach(), m.out[1].bias.detach(), do, dg_nc,
                 db_nc)
        def out1_leaves(dg_nc=dg_nc, db_nc=db_nc):
            gw, gb, sums = _f32(nf, dev=dev), _f32(nf, dev=dev), _f32(2, nf, dev=dev)
            L.rows_sum(dg_nc, n, nf, gw)
            L.rows_sum(db_nc, n, nf, gb)
            L.chan_reduce(do, nf, n * h * h, nf, sums, S.ws2, mode=2)  # out.0 (Conv 256->128 on cat(u2, x0)): bias
            return gw, gb, sums[0]
        G["out.1.weight"], G["out.1.bias"], G["out.0.bias"] = leaf(out1_leaves)
        G["out.0.weight"] = conv_wgrad(do, [S.u2, S.x0], nf)
        d_cat = _bf(n, h, h, 2 * nf, dev=dev)
        L.conv3x3(do, P["out0.d"], S.ones[:2 * nf], S.zeros[:2 * nf], d_cat, flags=0, mode=S.mode)
        # ---- up2
        d, sm = rcb_bwd("up2.2", "up2.model.2", d_cat, 2 * nf, fuse_next=S.layers["up2.1.c2"])  # channels [0,128) of d_cat
        d, sm = cbr_bwd("up2.1.c2", "up2.model.1.conv2", d, nf, sums=sm, fuse_next=S.layers["up2.1.c1"])
        d_v2, _ = cbr_bwd("up2.1.c1", "up2.model.1.conv1", d, nf, sums=sm)
        h2 = h // 2
        da2 = convT_bwd("up2", "up2.model.0", d_v2, [S.u1f.view(-1, nf), S.d1.view(-1, nf)])  # [n*32*32, 256]
        # FiLM2: u1f = cemb2*u1 + temb2
        d_u1 = _bf(n, h2, h2, nf, dev=dev)
        dcemb2, dtemb2 = _f32(n, nf, dev=dev), _f32(n, nf, dev=dev)
        L.film_bwd(da2, 2 * nf, S.u1, n, h2 * h2, nf, S.cemb2, d_u1, dcemb2, dtemb2)
        trows = S.t.shape[0]
        leaf(lambda: (embed_bwd(m.contextembed2, "contextembed2", S.c, dcemb2),
                      embed_bwd(m.timeembed2, "timeembed2", S.t, dtemb2 if trows == n else dtemb2.sum(0, keepdim=True))))
        # ---- up1
        d, sm = rcb_bwd("up1.2", "up1.model.2", d_u1, nf, fuse_next=S.layers["up1.1.c2"])
        d, sm = cbr_bwd("up1.1.c2", "up1.model.1.conv2", d, nf, sums=sm, fuse_next=S.layers["up1.1.c1"])
        d_v1, _ = cbr_bwd("up1.1.c1", "up1.model.1.conv1", d, nf, sums=sm)
        h4 = h // 4
        da1 = convT_bwd("up1", "up1.model.0", d_v1, [S.u0f.view(-1, 2 * nf), S.d2.view(-1, 2 * nf)])  # [n*256, 512]
        # ---- up0: GroupNorm + ReLU + FiLM backward, then the [B,256]x[256,65536] GEMM
        d_u0raw = _bf(n, h4 * h4, 2 * nf, dev=dev)
        dg_nc, db_nc = _f32(n, 2 * nf, dev=dev), _f32(n, 2 * nf, dev=dev)
        dcemb1, dtemb1 = _f32(n, 2 * nf, dev=dev), _f32(n, 2 * nf, dev=dev)
        L.gn_bwd(S.u0raw, da1, 4 * nf, n, h4 * h4, 2 * nf, 8, S.gn0_mr, m.up0[1].weight.detach(),
                 m.up0[1].bias.detach(), d_u0raw, dg_nc, db_nc, film_scale=S.cemb1, dfs=dcemb1, dfb=dtemb1)
        def up0_leaves(dg_nc=dg_nc, db_nc=db_nc):
            gw, gb, sums = _f32(2 * nf, dev=dev), _f32(2 * nf, dev=dev), _f32(2, 2 * nf, dev=dev)
            L.rows_sum(dg_nc, n, 2 * nf, gw)
            L.rows_sum(db_nc, n, 2 * nf, gb)
            L.chan_reduce(d_u0raw, 2 * nf, n * h4 * h4, 2 * nf, sums, S.ws2, mode=2)
            embed_bwd(m.contextembed1, "contextembed1", S.c, dcemb1)
            embed_bwd(m.timeembed1, "timeembed1", S.t, dtemb1 if trows == n else dtemb1.sum(0, keepdim=True))
            return gw, gb, sums[0]
        G["up0.1.weight"], G["up0.1.bias"], G["up0.0.bias"] = leaf(up0_leaves)
        K0 = h4 * h4 * 2 * nf
        d_hid = _bf(n, 2 * nf, dev=dev)
        L.gemm(d_u0raw.view(n, K0), P["up0.d"], S.zeros[:2 * nf], d_hid, shift_mod=2 * nf, workspace=S.skws)
        dw0 = _f32(2 * nf, K0, dev=dev, zero=True)
        L.gemm_tn(S.hidden, d_u0raw, dw0, n_img=1, H=1, W=n, a_c=2 * nf, b_c=K0, M=2 * nf, N=K0, ldc=K0,
                  workspace=S.skws)  # K = batch rows: a single K slice unless the batch exceeds 128
        dw0_iohw = dw0.view(2 * nf, h4, h4, 2 * nf).permute(0, 3, 1, 2)  # IOHW view of the [ci][(h,w,co)] GEMM output
        direct = getattr(ctx, "flat_out", None)
        if direct is not None:  # 78 % of all gradient bytes: permuted straight into its slot of the flat buffer
            direct["direct"]["up0.0.weight"].copy_(dw0_iohw)
            if _world() > 1:
                # ... and all-reduced right away on a communication stream: the down-path backward (a third of the
                # step) hides the 67 MB transfer; the small tensors follow in one all-reduce at the end
                comm = _side_stream(dev, 1)
                ready = torch.cuda.Event()
                ready.record(main)
                with torch.cuda.stream(comm):
                    comm.wait_event(ready)
                    dist.all_reduce(direct["direct"]["up0.0.weight"])
        else:
            G["up0.0.weight"] = dw0_iohw.contiguous()
        # ---- to_vec backward joins the skip gradient of d2
        d_d2 = da1[:, 2 * nf:].contiguous().view(n, h4, h4, 2 * nf)
        L.avgpool_gelu_bwd(S.hid_pre, d_hid.float().contiguous(), n, h4 * h4, 2 * nf, d_d2)
        # ---- down2
        dy = _bf(n, h2, h2, 2 * nf, dev=dev)
        L.maxpool2_bwd(d_d2, 2 * nf, S.pool2_in, dy)
        d, sm = rcb_bwd("down2.1", "down2.model.1", dy, 2 * nf, fuse_next=S.layers["down2.0.c2"])
        d, sm = cbr_bwd("down2.0.c2", "down2.model.0.conv2", d, 2 * nf, sums=sm, fuse_next=S.layers["down2.0.c1"])
        d_d1, _ = cbr_bwd("down2.0.c1", "down2.model.0.conv1", d, 2 * nf, sums=sm)
        L.add_bf16(d_d1, nf, da2[:, nf:], 2 * nf, n * h2 * h2, nf)
        # ---- down1
        dy = _bf(n, h, h, nf, dev=dev)
        L.maxpool2_bwd(d_d1, nf, S.pool1_in, dy)
        d, sm = rcb_bwd("down1.1", "down1.model.1", dy, nf, fuse_next=S.layers["down1.0.c2"])
        d, sm = cbr_bwd("down1.0.c2", "down1.model.0.conv2", d, nf, sums=sm, fuse_next=S.layers["down1.0.c1"])
        d_x0, _ = cbr_bwd("down1.0.c1", "down1.model.0.conv1", d, nf, sums=sm)
        L.add_bf16(d_x0, nf, d_cat.view(-1, 2 * nf)[:, nf:], 2 * nf, n * h * h, nf)
        # ---- init_conv: x0 = y2 + shortcut(x) (the shortcut is not a parameter)
        d_y1, sm = cbr_bwd("init_conv.c2", "init_conv.conv2", d_x0, nf, fuse_next=S.layers["init_conv.c1"])
        cbr_bwd("init_conv.c1", "init_conv.conv1", d_y1, nf, sums=sm)
        main.wait_stream(side)  # join: every weight gradient is complete before it is reduced / handed out
        keep.clear()
        W = _world()
        flat_out = getattr(ctx, "flat_out", None)
        if flat_out is not None:
            # captured step: gradients land in ONE flat buffer that the parameters' .grad views alias ([all-reduced
            # tensors | BatchNorm affine tensors, already global]): cat + all-reduce + scale = 3 launches, not ~300
            assert set(flat_out["order"][flat_out["n_reduce"]:]) == bn_affine
            G["__pad__"] = S.zeros[:flat_out["pad"]]  # alignment filler in front of up0.0.weight's slot
            for names, dst in flat_out["segments"]:  # the tensors around the directly written ones
                torch.cat([G[k].reshape(-1) for k in names], out=dst)
            if W > 1:
                dist.all_reduce(flat_out["segments"][0][1])  # everything in front of up0.0.weight
                main.wait_stream(_side_stream(dev, 1))       # ... whose own all-reduce has been running since
                flat_out["flat"].mul_(1.0 / W)
            ctx.S = None
            return None
        # ---- data parallel: one flat all-reduce, then the 1/world of the global-batch mean
        if W > 1:
            names = [k for k in ctx.names if k not in bn_affine]
            flat = torch.cat([G[k].reshape(-1) for k in names])
            dist.all_reduce(flat)
            o = 0
            for k in names:
                cnt = G[k].numel()
                G[k] = flat[o:o + cnt].view_as(G[k])
                o += cnt
            for k in ctx.names:
                G[k] = G[k] / W
        ctx.S = None
        return (None, None, None, None, None) + tuple(G[k].view_as(p) for k, p in zip(ctx.names, m.parameters()))


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr) defaults (betas 0.9/0.999, eps 1e-8, no weight decay) — the optimiser of
    code/train_diffusion_paper.py:318 — as ONE multi-tensor kernel launch per step."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._tab, self._key, self._step = None, None, 0

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None
        self._step += 1
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                st = self.state[p]
                if not st:
                    st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(p), torch.zeros_like(p)
                # same per-parameter keys as torch.optim.Adam, so optimizer state_dicts are interchangeable
                st["step"] = torch.tensor(float(self._step))
            grads = [p.grad.contiguous() for p in ps]
            key = tuple((p.data_ptr(), g.data_ptr()) for p, g in zip(ps, grads))
            if self._tab is None or self._key != (gi, key):
                rows = [[p.data_ptr(), g.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                         self.state[p]["exp_avg_sq"].data_ptr(), p.numel()] for p, g in zip(ps, grads)]
                self._tab = torch.tensor(rows, dtype=torch.int64).to(ps[0].device)
                self._key = (gi, key)
            self._grads_keepalive = grads
            b1, b2 = group["betas"]
            L.adam_step(self._tab, len(ps), max(p.numel() for p in ps), group["lr"], b1, b2, group["eps"], self._step)


    def load_state_dict(self, state_dict):
        """Accepts a FusedAdam or a torch.optim.Adam state_dict (code/train_diffusion_paper.py:318 optimiser)."""
        super().load_state_dict(state_dict)
        steps = [int(st["step"]) for st in self.state.values() if "step" in st]
        self._step = max(steps) if steps else 0
        self._tab = self._key = None  # exp_avg tensors were replaced: rebuild the pointer table


def _draw_t(timesteps, n):
    """t ~ randint(1, T+1) for this rank's n samples (train_diffusion_paper.py:353).  Under data parallelism every
    rank draws the GLOBAL batch's timesteps and keeps its own slice: with the CPU generators in step (same seed on
    every rank, which the shared shortcut draw requires anyway) the ranks hold distinct draws — together exactly what
    the single-process reference draws for the global batch — and the generators stay in step."""
    w, r = _world(), _rank()
    return torch.randint(1, timesteps + 1, (n * w,))[r * n:(r + 1) * n]


def training_step(model, optim, x, param, timesteps, ab_t, *, noise=None, t=None, shortcut=None):
    """The loop body of code/train_diffusion_paper.py:350-364 on device: noise, t, perturb_input, forward, MSE,
    backward, optimizer step.  Returns the (local) loss as a 0-d tensor; `loss.item()` is left to the caller."""
    dev = model._check_supported()
    n = x.shape[0]
    x = x.to(dev, torch.float32).contiguous()
    if t is None:
        t = _draw_t(timesteps, n)
    t = t.to(dev)
    x_pert = torch.empty_like(x)
    ca, cb = ab_t.sqrt().contiguous(), (1 - ab_t).contiguous()
    if noise is None:
        noise = torch.empty_like(x)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        L.perturb(x, x_pert, ca, cb, t_idx=t.to(torch.int64).contiguous(), seed=seed, noise_out=noise,
                  sample_offset=_rank() * n)
    else:
        noise = noise.to(dev, torch.float32).contiguous()
        L.perturb(x, x_pert, ca, cb, noise=noise, t_idx=t.to(torch.int64).contiguous())
    optim.zero_grad(set_to_none=True)
    pred = model(x_pert, t / timesteps, param, shortcut=shortcut)
    # F.mse_loss + its gradient in one kernel; the backward pass starts from d pred directly
    dpred = torch.empty_like(pred)
    partial = torch.empty(L.num_sms() * 8, device=dev)
    loss_sum = torch.empty(1, device=dev)
    L.mse_grad(pred.detach(), noise, 1.0 / pred.numel(), dpred, partial, loss_sum)
    pred.backward(dpred)
    optim.step()
    return loss_sum[0] / pred.numel()


class GraphedTrainStep:
    """The whole training step (in-kernel noise + perturb_input, train-mode forward, MSE + gradient, backward, cross-rank
    reductions, Adam) captured ONCE as a CUDA graph and replayed: ~400 kernel launches become one graph launch, which is
    what bounds the step at 32 images per GPU (BASELINE config 3).  Per step the host only copies the batch, the
    timesteps `t` and the fresh shortcut draw (both from torch's CPU generator, like the reference) into static buffers.
    Learning rate and Adam's step count live on the device so the captured launch stays valid across epochs."""

    def __init__(self, model, batch, timesteps, ab_t, lr, betas=(0.9, 0.999), eps=1e-8, seed=0, warmup=2,
                 use_graph=True):
        dev = model._check_supported()
        self.model, self.T, self.dev, self.B = model, timesteps, dev, batch
        h, ncf = model.h, model.n_cfeat
        self.x = torch.zeros(batch, 1, h, h, device=dev)
        self.param = torch.zeros(batch, ncf, device=dev)
        self.t = torch.ones(batch, device=dev, dtype=torch.int64)
        self.sc = torch.zeros(2 * model.n_feat, device=dev)
        self.ca, self.cb = ab_t.to(dev).sqrt().contiguous(), (1 - ab_t.to(dev)).contiguous()
        self.noise, self.x_pert = torch.empty_like(self.x), torch.empty_like(self.x)
        self.dpred = torch.empty_like(self.x)
        self.partial, self.loss_sum = torch.empty(L.num_sms() * 8, device=dev), torch.zeros(1, device=dev)
        self.lr = torch.full((1,), float(lr), device=dev)
        self.count = torch.zeros(1, device=dev, dtype=torch.int32)  # Adam step == Philox stream offset
        self.betas, self.eps, self.seed = betas, eps, seed
        self.params = [p for p in model.parameters()]
        # one flat gradient buffer, [tensors that are all-reduced | BatchNorm affine tensors (their gradients come from
        # already-global sums)]; every parameter's .grad is a view into it
        named = list(model.named_parameters())
        is_bn = lambda nme: ".conv1.1." in nme or ".conv2.1." in nme  # noqa: E731  (the BatchNorm2d of a Conv-BN-ReLU)
        big = "up0.0.weight"  # written by its producer directly (never copied through the concatenation)
        order = [k for k, _ in named if not is_bn(k) and k != big] + [big] + [k for k, _ in named if is_bn(k)]
        by_name = dict(named)
        # out.3.bias has ONE element: without padding everything behind it — 78 % of the buffer is up0.0.weight's
        # gradient — sits at an address that is not a multiple of 16 bytes and the optimizer's float4 path is lost.
        # "__pad__" is a pseudo-tensor of zeros in front of `big` (backward supplies it to the concatenation).
        i_big = order.index(big)
        pad = (-sum(by_name[k].numel() for k in order[:i_big])) % 4
        self.flat = torch.zeros(sum(p.numel() for p in self.params) + pad, device=dev)
        o, start = 0, {}
        for k in order:
            if k == big:
                o += pad
            p = by_name[k]
            start[k] = o
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()
        end_big = start[big] + by_name[big].numel()
        self.flat_out = {"flat": self.flat, "order": order, "n_reduce": sum(1 for k in order if not is_bn(k)),
                         "numel_reduce": sum(by_name[k].numel() for k in order if not is_bn(k)),
                         "direct": {big: by_name[big].grad}, "pad": pad,
                         "segments": [(order[:i_big] + (["__pad__"] if pad else []), self.flat[:start[big]]),
                                      (order[i_big + 1:], self.flat[end_big:])]}
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        rows = [[p.data_ptr(), p.grad.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()]
                for p, m, v in zip(self.params, self.m, self.v)]
        self.table = torch.tensor(rows, dtype=torch.int64).to(dev)
        self.max_numel = max(p.numel() for p in self.params)
        self.graphs, self.use_graph, self.warmup = {}, use_graph, warmup
        if use_graph:
            self._capture(batch)

    @property
    def graph(self):
        return self.graphs.get(self.B)

    @L.on_device
    def _capture(self, b):
        """Capture the step for batch size b <= B (the ragged last batch of an epoch gets its own graph, lazily).
        Warm-up and capture run with lr = 0, and everything a step mutates besides the parameters — BatchNorm
        buffers, Adam moments, the step counter — is saved and restored, so capturing mid-training is invisible."""
        model = self.model
        saved_buf = [t.detach().clone() for t in model.buffers()]
        saved_m, saved_v = [t.clone() for t in self.m], [t.clone() for t in self.v]
        saved_count, saved_lr = self.count.clone(), self.lr.clone()
        self.lr.zero_()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._step(b)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._step(b)
        torch.cuda.synchronize()
        with torch.no_grad():
            for t, sv in zip(model.buffers(), saved_buf):
                t.copy_(sv)
            for t, sv in zip(self.m + self.v, saved_m + saved_v):
                t.copy_(sv)
            self.count.copy_(saved_count)
            self.lr.copy_(saved_lr)
        self.graphs[b] = g

    def _step(self, b=None):
        m = self.model
        b = self.B if b is None else b
        x, x_pert, noise, dpred, t = self.x[:b], self.x_pert[:b], self.noise[:b], self.dpred[:b], self.t[:b]
        L.step_advance(self.count, 1)
        # the in-kernel noise is keyed by the GLOBAL sample index (rank * B + i): ranks of a data-parallel step never
        # share noise, as the reference's iid draws over the global batch (train_diffusion_paper.py:352)
        L.perturb(x, x_pert, self.ca, self.cb, t_idx=t, step_ptr=self.count, seed=self.seed, noise_out=noise,
                  sample_offset=_rank() * self.B)
        ctx = _Ctx()
        pred = _UnetFn.forward(ctx, m, x_pert, t / self.T, self.param[:b], self.sc, *self.params)
        L.mse_grad(pred, noise, 1.0 / pred.numel(), dpred, self.partial, self.loss_sum)
        ctx.flat_out = self.flat_out
        _UnetFn.backward(ctx, dpred)  # gradients land in self.flat (= every p.grad)
        b1, b2 = self.betas
        L.adam_step(self.table, len(self.params), self.max_numel, 0.0, b1, b2, self.eps, 0, lr_dev=self.lr,
                    step_dev=self.count)

    def set_lr(self, lr):
        self.lr.fill_(float(lr))

    def reset_optimizer(self):
        for t in self.m + self.v:
            t.zero_()
        self.count.zero_()

    def state_dict(self):
        """Optimizer state in torch.optim.Adam's state_dict layout (state[i] = {step, exp_avg, exp_avg_sq} in
        model.parameters() order), so a checkpoint moves between this step, FusedAdam and the reference's Adam."""
        step = float(self.count.item())
        state = {i: {"step": torch.tensor(step), "exp_avg": m.detach().clone(), "exp_avg_sq": v.detach().clone()}
                 for i, (m, v) in enumerate(zip(self.m, self.v))} if step > 0 else {}
        group = dict(lr=float(self.lr.item()), betas=self.betas, eps=self.eps, weight_decay=0, amsgrad=False,
                     params=list(range(len(self.params))))
        return {"state": state, "param_groups": [group], "seed": self.seed}

    def load_state_dict(self, sd):
        self.reset_optimizer()
        steps = [int(st["step"]) for st in sd["state"].values()]
        with torch.no_grad():
            for i, st in sd["state"].items():
                self.m[int(i)].copy_(st["exp_avg"])
                self.v[int(i)].copy_(st["exp_avg_sq"])
        self.count.fill_(max(steps) if steps else 0)
        self.lr.fill_(float(sd["param_groups"][0]["lr"]))
        self.seed = int(sd.get("seed", self.seed))

    @L.on_device
    def __call__(self, x, param, t=None, shortcut=None):
        """One optimisation step on a batch of b <= B samples; returns the mean-squared-error loss as a 0-d device
        tensor (no host sync).  t / shortcut default to fresh draws from torch's CPU generator, like the reference."""
        m = self.model
        b = x.shape[0]
        if b > self.B:
            raise L.CdmError(f"batch of {b} samples exceeds the captured capacity {self.B}")
        self.x[:b].copy_(x.reshape(b, *self.x.shape[1:]), non_blocking=True)
        self.param[:b].copy_(param, non_blocking=True)
        self.t[:b].copy_(_draw_t(self.T, b) if t is None else t, non_blocking=True)
        self.sc.copy_(m.draw_shortcut() if shortcut is None else shortcut, non_blocking=True)
        if self.use_graph:
            if b not in self.graphs:
                self._capture(b)
            self.graphs[b].replay()
        else:
            self._step(b)
        return self.loss_sum[0] / (b * self.x[0].numel())
