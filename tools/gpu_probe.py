"""GPU bring-up probe for the tcgen05 kernels: runs each case in its own process
(with a timeout, so a trapped / hung kernel cannot take the others down) and
prints one line per case.  Usage on the GPU box:  python tools/gpu_probe.py [case ...]
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the measurement probes (flag bits 26-30, cdm_gemm_tn_args.probe) exist only in the -DCDM_PROBES build
os.environ.setdefault("CDM_LIB", os.path.join(ROOT, "camels-diffusion-model_b200", "libcdm_b200_probes.so"))


def _load():
    import camels_diffusion_model_b200 as pkg  # noqa: F401
    from camels_diffusion_model_b200 import _lib
    return _lib


def rel_l2(a, b):
    import torch
    return (torch.linalg.vector_norm((a - b).float()) / (torch.linalg.vector_norm(b.float()) + 1e-30)).item()


def diag(name, out, ref, tol=1e-2):
    """out/ref: [n,H,W,C] float tensors."""
    import torch
    err = rel_l2(out, ref)
    ok = err < tol and bool(torch.isfinite(out).all())
    print(f"CASE {name}: {'PASS' if ok else 'FAIL'} rel_l2={err:.3e} max_abs={float((out - ref).abs().max()):.3e} "
          f"ref_rms={float(ref.pow(2).mean().sqrt()):.3e}", flush=True)
    if not ok and out.dim() == 4:
        bad = ((out - ref).abs() > 0.05 * ref.abs().max()).float()
        print("  bad frac by w%16:", [round(float(bad[:, :, i::16].mean()), 3) for i in range(16)])
        print("  bad frac by h%16:", [round(float(bad[:, i::16].mean()), 3) for i in range(16)])
        print("  bad frac by c//16:", [round(float(bad[..., i * 16:(i + 1) * 16].mean()), 3)
                                       for i in range(out.shape[-1] // 16)])
        print("  bad frac by img:", [round(float(bad[i].mean()), 3) for i in range(out.shape[0])])
        print("  out[0,0,0,:8]", out[0, 0, 0, :8].tolist())
        print("  ref[0,0,0,:8]", ref[0, 0, 0, :8].tolist())
        print("  out[0,1,1,:8]", out[0, 1, 1, :8].tolist())
        print("  ref[0,1,1,:8]", ref[0, 1, 1, :8].tolist())
    return ok


def conv_ref(x, w, scale, shift):
    """x: [n,H,W,cin] bf16, w: [cout,3,3,cin] bf16 -> fp32 NHWC."""
    import torch
    import torch.nn.functional as F
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), padding=1)
    y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    return y.permute(0, 2, 3, 1).contiguous()


def case_conv(name, mode, n, H, c0, c1, cout, extra=None):
    import torch
    L = _load()
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"
    cin = c0 + c1
    x = torch.randn(n, H, H, cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, 3, 3, cin, device=dev) / (3 * cin ** 0.5)).to(torch.bfloat16)
    scale = torch.rand(cout, device=dev) + 0.5
    shift = torch.randn(cout, device=dev) * 0.1
    s0 = x[..., :c0].contiguous()
    s1 = x[..., c0:].contiguous() if c1 else None
    ref = conv_ref(x, w, scale, shift)
    flags = L.EPI_RELU
    kw = {}
    ref = ref.clamp_min(0)
    Ho = H
    if extra == "pool":
        flags |= L.EPI_POOL
        ref = torch.nn.functional.max_pool2d(ref.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).contiguous()
        Ho = H // 2
    elif extra == "film":
        flags |= L.EPI_FILM
        fs = torch.randn(n, cout, device=dev)
        fsh = torch.randn(3, 1, cout, device=dev)
        step = torch.tensor([2], device=dev, dtype=torch.int32)
        kw = dict(film_scale=fs, film_shift=fsh, film_shift_rows=1, step_ptr=step)
        ref = ref * fs.view(n, 1, 1, cout) + fsh[2, 0].view(1, 1, 1, cout)
    elif extra == "shortcut":
        flags |= L.EPI_SHORTCUT
        nx = n // 2
        xs = torch.randn(nx, H, H, device=dev)
        tab = torch.rand(4, 2, 2, cout, device=dev) * 2 - 1
        step = torch.tensor([3], device=dev, dtype=torch.int32)
        kw = dict(sc_x=xs, sc_tab=tab, step_ptr=step)
        xx = torch.cat([xs, xs], 0).view(n, H, H, 1)
        half = torch.arange(n, device=dev) // nx
        ref = ref + xx * tab[3, half, 0].view(n, 1, 1, cout) + tab[3, half, 1].view(n, 1, 1, cout)
    elif extra == "gn":
        flags = L.EPI_GNSTATS  # raw conv output + statistics
        ref = conv_ref(x, w, scale, shift)
        slots = (H // 16) * (H // 16) * 8
        part = torch.zeros(n, slots, 8, 2, device=dev)
        kw = dict(gn_partial=part)
    out = torch.full((n, Ho, Ho, cout), float("nan"), device=dev).to(torch.bfloat16)
    L.conv3x3(s0, w, scale, shift, out, src1=s1, flags=flags, mode=mode, **kw)
    torch.cuda.synchronize()
    ok = diag(name, out.float(), ref)
    if extra == "gn":
        g = ref.view(n, H * H, 8, 16)
        s_ref = torch.stack([g.sum((1, 3)), g.pow(2).sum((1, 3))], -1)  # [n,8,2]
        s_out = part.sum(1)
        print(f"CASE {name}/stats: {'PASS' if rel_l2(s_out, s_ref) < 1e-4 else 'FAIL'} rel_l2={rel_l2(s_out, s_ref):.3e}")
    return ok


def case_gemm(name, M, k0, k1, N, out_mode=0, H=0, W=0):
    import torch
    L = _load()
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"
    K = k0 + k1
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    bw = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    smod = 128 if out_mode == 1 else N
    shift = torch.randn(smod, device=dev)
    ref = a.float() @ bw.float().t() + shift.repeat(N // smod).view(1, N)
    a0 = a[:, :k0].contiguous()
    a1 = a[:, k0:].contiguous() if k1 else None
    if out_mode == 0:
        out = torch.full((M, N), float("nan"), device=dev).to(torch.bfloat16)
        L.gemm(a0, bw, shift, out, a1=a1)
        torch.cuda.synchronize()
        return diag(name, out.float().view(1, 1, M, N), ref.view(1, 1, M, N))
    n_img = M // (H * W)
    out = torch.full((n_img, 2 * H, 2 * W, 128), float("nan"), device=dev).to(torch.bfloat16)
    L.gemm(a0, bw, shift, out, a1=a1, out_mode=1, H=H, W=W, shift_mod=128)
    torch.cuda.synchronize()
    r = ref.view(n_img, H, W, 2, 2, 128).permute(0, 1, 3, 2, 4, 5).reshape(n_img, 2 * H, 2 * W, 128)
    return diag(name, out.float(), r)


def case_perf(name, mode, n=256, H=64, cin=128, cout=128, flags=1, iters=10):
    import torch
    L = _load()
    dev = "cuda"
    x = torch.randn(n, H, H, cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, 3, 3, cin, device=dev) / (3 * cin ** 0.5)).to(torch.bfloat16)
    scale = torch.ones(cout, device=dev)
    shift = torch.zeros(cout, device=dev)
    out = torch.empty(n, H, H, cout, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        L.conv3x3(x, w, scale, shift, out, mode=mode, flags=flags)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        L.conv3x3(x, w, scale, shift, out, mode=mode, flags=flags)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flop = 2.0 * n * H * H * cout * 9 * cin
    print(f"CASE {name}: PERF {ms:.3f} ms  {flop / ms / 1e9:.1f} TFLOP/s  (n={n} H={H} cin={cin} cout={cout})", flush=True)
    # cuDNN bf16 channels_last for comparison
    import torch.nn.functional as F
    xc = x.permute(0, 3, 1, 2)
    wc = w.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    for _ in range(3):
        F.conv2d(xc, wc, padding=1)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        F.conv2d(xc, wc, padding=1)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"CASE {name}/cudnn_bf16_nhwc: PERF {ms2:.3f} ms  {flop / ms2 / 1e9:.1f} TFLOP/s", flush=True)
    return True


def case_waits(name, n=2048, H=64, cin=128, cout=128, flags=1, mode=3):
    """Where the MMA issuer of conv3x3_sw_kernel waits (probe flag bit 28): cycles blocked on the TMEM buffer
    (epilogue too slow), on the pixel-halo ring and on the weight ring, per CTA, against the CTA's total."""
    import torch
    L = _load()
    dev = "cuda"
    x = torch.randn(n, H, H, cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, 3, 3, cin, device=dev) / (3 * cin ** 0.5)).to(torch.bfloat16)
    scale, shift = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    out = torch.empty(n, H, H, cout, device=dev, dtype=torch.bfloat16)
    dbg = torch.zeros(148, 8, device=dev)
    for _ in range(3):
        L.conv3x3(x, w, scale, shift, out, mode=mode, flags=flags | (1 << 28), gn_partial=dbg)
    torch.cuda.synchronize()
    d = dbg.cpu()
    tot = d[:, 3]
    print(f"CASE {name}: issuer cycles/CTA {tot.mean():.0f} (min {tot.min():.0f} max {tot.max():.0f}); blocked on "
          f"tmem {100 * (d[:, 0] / tot).mean():.1f}%  halo ring {100 * (d[:, 1] / tot).mean():.1f}%  weight ring "
          f"{100 * (d[:, 2] / tot).mean():.1f}%; epilogue warp: idle {100 * (d[:, 4] / tot).mean():.1f}%  busy "
          f"{100 * (d[:, 5] / tot).mean():.1f}%  (n={n} H={H} cin={cin} cout={cout} flags={flags})", flush=True)
    return True


def case_wgrad(name, n=256, H=64, cin=128, cout=128, use_ws=True):
    """3x3 weight-gradient GEMM (cdm_gemm_tn taps=9): correctness vs autograd on a small case + throughput."""
    import torch
    import torch.nn.functional as F
    L = _load()
    dev = "cuda"
    xs = torch.randn(2, H, H, cin, device=dev).to(torch.bfloat16)
    dzs = torch.randn(2, H, H, cout, device=dev).to(torch.bfloat16)
    dw = torch.zeros(cout, 9 * cin, device=dev)
    L.gemm_tn(dzs, xs, dw, n_img=2, H=H, W=H, a_c=cout, b_c=cin, M=cout, N=cin, ldc=9 * cin, taps=9, tap_stride=cin)
    w = torch.zeros(cout, cin, 3, 3, device=dev, requires_grad=True)
    F.conv2d(xs.float().permute(0, 3, 1, 2), w, padding=1).backward(dzs.float().permute(0, 3, 1, 2))
    err = rel_l2(dw.view(cout, 3, 3, cin).permute(0, 3, 1, 2), w.grad)
    ws = torch.empty(160 * 128 * 384, device=dev) if use_ws else None
    dw2 = torch.zeros(cout, 9 * cin, device=dev)
    L.gemm_tn(dzs, xs, dw2, n_img=2, H=H, W=H, a_c=cout, b_c=cin, M=cout, N=cin, ldc=9 * cin, taps=9, tap_stride=cin,
              workspace=ws)
    err = max(err, rel_l2(dw2.view(cout, 3, 3, cin).permute(0, 3, 1, 2), w.grad))
    x = torch.randn(n, H, H, cin, device=dev).to(torch.bfloat16)
    dz = torch.randn(n, H, H, cout, device=dev).to(torch.bfloat16)
    dw = torch.zeros(cout, 9 * cin, device=dev)
    for _ in range(3):
        L.gemm_tn(dz, x, dw, n_img=n, H=H, W=H, a_c=cout, b_c=cin, M=cout, N=cin, ldc=9 * cin, taps=9, tap_stride=cin,
                  workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        L.gemm_tn(dz, x, dw, n_img=n, H=H, W=H, a_c=cout, b_c=cin, M=cout, N=cin, ldc=9 * cin, taps=9, tap_stride=cin,
                  workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    flop = 2.0 * n * H * H * cout * 9 * cin
    pr = torch.zeros(148, 4, device=dev)
    L.gemm_tn(dz, x, dw, n_img=n, H=H, W=H, a_c=cout, b_c=cin, M=cout, N=cin, ldc=9 * cin, taps=9, tap_stride=cin, probe=pr)
    torch.cuda.synchronize()
    pr = pr.cpu()
    tot = pr[:, 2].clamp_min(1)
    print(f"   issuer: total {tot.mean():.0f} clk/CTA, blocked on TMA ring {100 * (pr[:, 0] / tot).mean():.1f}%, on accumulator "
          f"{100 * (pr[:, 1] / tot).mean():.1f}%; epilogue {100 * (pr[:, 3] / tot).mean():.1f}% of the issuer's span")
    print(f"CASE {name}: {'PASS' if err < 1e-4 else 'FAIL'} rel_l2={err:.2e}  PERF {ms:.3f} ms  {flop / ms / 1e9:.1f} TFLOP/s  "
          f"(n={n} H={H} cin={cin} cout={cout})", flush=True)
    return err < 1e-4


def case_probe_l2(name):
    import torch
    L = _load()
    dev = "cuda"
    for mb in (8, 64):
        n_rows = mb * 1024 * 1024 // 128
        buf = torch.zeros(n_rows, 64, device=dev, dtype=torch.bfloat16)
        iters = 4000
        L.probe_tma_l2(buf, n_rows, iters)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.probe_tma_l2(buf, n_rows, iters)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        sms = torch.cuda.get_device_properties(0).multi_processor_count
        tot = sms * (iters + 8) * 16384
        print(f"CASE {name}[{mb}MB]: PERF {ms:.3f} ms  {tot / ms / 1e6:.1f} GB/s TMA L2->smem "
              f"({tot / ms / 1e6 / sms:.1f} GB/s per SM)", flush=True)
    return True


def case_hbm_rates(name, gib=2):
    """Write-only, read-only and copy HBM rates of this box (torch fill_ / sum / copy_ over `gib` GiB, best of 10,
    CUDA events): conv_in writes 1 MiB per image and reads 16 KB, conv_out the opposite, so their rooflines are the
    one-directional rates, not the read+write copy figure of MEASURED_PEAKS.json."""
    import torch
    n = gib * (1 << 30) // 4
    a, b = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)

    def best(fn, nbytes):
        t = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t.append(e0.elapsed_time(e1))
        return nbytes / (min(t) * 1e-3) / 1e9

    w = best(lambda: a.fill_(1.0), 4 * n)
    r = best(lambda: a.sum(), 4 * n)
    c = best(lambda: b.copy_(a), 8 * n)
    print(f"CASE {name}: PERF write-only {w:.0f} GB/s, read-only {r:.0f} GB/s, copy (read+write) {c:.0f} GB/s "
          f"over {gib} GiB", flush=True)
    return True


def case_tmem_layout(name):
    """Register <- TMEM element mapping of tcgen05.ld.16x256b.x4 (probe library only)."""
    import ctypes as C
    import torch
    L = _load()
    out = torch.zeros(4, 2, 32, 16, device="cuda")
    rc = L.lib().cdm_probe_tmem_layout(C.c_void_p(out.data_ptr()), L.stream_ptr())
    torch.cuda.synchronize()
    assert rc == 0, L.lib().cdm_last_error()
    o = out.cpu()
    lane, col = (o // 1000).long(), (o % 1000).long()
    ok = True
    for w in range(4):
        for h in range(2):
            for t in range(32):
                for i in range(16):
                    j, k = i // 4, i % 4
                    exp_lane = w * 32 + h * 16 + t // 4 + (8 if k >= 2 else 0)
                    exp_col = 8 * j + 2 * (t % 4) + (k & 1)
                    if int(lane[w, h, t, i]) != exp_lane or int(col[w, h, t, i]) != exp_col:
                        if ok:
                            print(f"   first mismatch: warp {w} half {h} thread {t} reg {i}: got lane {int(lane[w, h, t, i])} "
                                  f"col {int(col[w, h, t, i])}, expected lane {exp_lane} col {exp_col}")
                        ok = False
    print(f"CASE {name}: {'PASS' if ok else 'FAIL'} (reg 4j+k of thread t <- lane t/4 + 8*(k>=2) [+16*half], column 8j + 2(t%4) + (k&1))")
    print("   thread 5, half 0, regs:", [(int(lane[0, 0, 5, i]), int(col[0, 0, 5, i])) for i in range(16)])
    return ok


CASES = {
    "tmem_layout": lambda: case_tmem_layout("tmem_layout"),
    "hbm_rates": lambda: case_hbm_rates("hbm_rates"),
    "gemm_basic": lambda: case_gemm("gemm_basic", 256, 64, 0, 128),
    "gemm_ragged": lambda: case_gemm("gemm_ragged", 300, 128, 64, 256),
    "gemm_shuffle": lambda: case_gemm("gemm_shuffle", 2 * 16 * 16, 256, 256, 512, out_mode=1, H=16, W=16),
    "conv_m0": lambda: case_conv("conv_m0", 0, 3, 32, 128, 0, 128),
    "conv_m1": lambda: case_conv("conv_m1", 1, 3, 32, 128, 0, 128),
    "conv_m2": lambda: case_conv("conv_m2", 2, 3, 32, 128, 0, 128),
    "conv_big_m0": lambda: case_conv("conv_big_m0", 0, 5, 64, 64, 64, 256),
    "conv_big_m1": lambda: case_conv("conv_big_m1", 1, 5, 64, 64, 64, 256),
    "conv_big_m2": lambda: case_conv("conv_big_m2", 2, 5, 64, 64, 64, 256),
    "conv_many_m0": lambda: case_conv("conv_many_m0", 0, 80, 64, 128, 0, 128),
    "conv_pool": lambda: case_conv("conv_pool", 0, 4, 32, 128, 0, 256, "pool"),
    "conv_film": lambda: case_conv("conv_film", 0, 4, 32, 128, 0, 128, "film"),
    "conv_shortcut": lambda: case_conv("conv_shortcut", 0, 4, 64, 128, 0, 128, "shortcut"),
    "conv_gn": lambda: case_conv("conv_gn", 0, 3, 64, 128, 128, 128, "gn"),
    "perf_m0": lambda: case_perf("perf_m0", 0),
    "perf_m1": lambda: case_perf("perf_m1", 1),
    "perf_m2": lambda: case_perf("perf_m2", 2),
    "perf_m0_c256": lambda: case_perf("perf_m0_c256", 0, n=256, H=32, cin=256, cout=256),
    "perf_m3": lambda: case_perf("perf_m3", 3, n=1024),
    "perf_m4": lambda: case_perf("perf_m4", 4, n=1024),
    "perf_m4_c256": lambda: case_perf("perf_m4_c256", 4, n=1024, H=32, cin=256, cout=256),
    "perf_m4_out0": lambda: case_perf("perf_m4_out0", 4, n=1024, H=64, cin=256, cout=128),
    "perf_m3_pool": lambda: case_perf("perf_m3_pool", 3, n=1024, flags=1 | 4),
    "perf_m4_pool": lambda: case_perf("perf_m4_pool", 4, n=1024, flags=1 | 4),
    "perf_m3_b2048": lambda: case_perf("perf_m3_b2048", 3, n=2048),
    "perf_m4_b2048": lambda: case_perf("perf_m4_b2048", 4, n=2048),
    "waits_m4": lambda: case_waits("waits_m4", mode=4),
    "perf_m4_long": lambda: case_perf("perf_m4_long", 4, n=1024, iters=600),
    "perf_m4_halfw": lambda: case_perf("perf_m4_halfw", 4, n=1024, flags=1 | (1 << 27)),
    "perf_m4_halfw_notma": lambda: case_perf("perf_m4_halfw_notma", 4, n=1024, flags=1 | (1 << 27) | (1 << 24)),
    "perf_m4_halfw_noepi": lambda: case_perf("perf_m4_halfw_noepi", 4, n=1024, flags=1 | (1 << 27) | (1 << 29)),
    "perf_m4_nostage": lambda: case_perf("perf_m4_nostage", 4, n=1024, flags=1 | (1 << 25)),
    "perf_m4_notma": lambda: case_perf("perf_m4_notma", 4, n=1024, flags=1 | (1 << 24)),
    "perf_m4_nostage_notma": lambda: case_perf("perf_m4_nostage_notma", 4, n=1024, flags=1 | (1 << 24) | (1 << 25)),
    "perf_m4_noepi": lambda: case_perf("perf_m4_noepi", 4, n=1024, flags=1 | (1 << 29)),
    "perf_m3_nostore": lambda: case_perf("perf_m3_nostore", 3, n=1024, flags=1 | (1 << 30)),
    "perf_m3_noepi": lambda: case_perf("perf_m3_noepi", 3, n=1024, flags=1 | (1 << 29)),
    "perf_m3_l2store": lambda: case_perf("perf_m3_l2store", 3, n=1024, flags=1 | (1 << 26)),
    "perf_m3_halfw": lambda: case_perf("perf_m3_halfw", 3, n=1024, flags=1 | (1 << 27)),
    "perf_m3_halfw_nostore": lambda: case_perf("perf_m3_halfw_nostore", 3, n=1024, flags=1 | (1 << 27) | (1 << 30)),
    "perf_m2_big": lambda: case_perf("perf_m2_big", 2, n=1024),
    "perf_m3_c256": lambda: case_perf("perf_m3_c256", 3, n=1024, H=32, cin=256, cout=256),
    "perf_m2_c256": lambda: case_perf("perf_m2_c256", 2, n=1024, H=32, cin=256, cout=256),
    "perf_m3_out0": lambda: case_perf("perf_m3_out0", 3, n=1024, H=64, cin=256, cout=128),
    "perf_m2_out0": lambda: case_perf("perf_m2_out0", 2, n=1024, H=64, cin=256, cout=128),
    "probe_l2": lambda: case_probe_l2("probe_l2"),
    "wgrad": lambda: case_wgrad("wgrad"),
    "wgrad_atomics": lambda: case_wgrad("wgrad_atomics", use_ws=False),
    "wgrad_c256": lambda: case_wgrad("wgrad_c256", H=32, cin=256, cout=256),
    "wgrad_b32": lambda: case_wgrad("wgrad_b32", n=32),
    "waits": lambda: case_waits("waits"),
    "waits_nostore": lambda: case_waits("waits_nostore", flags=1 | (1 << 30)),
    "waits_c256": lambda: case_waits("waits_c256", H=32, cin=256, cout=256),
    "waits_out0": lambda: case_waits("waits_out0", cin=256, cout=128),
    "waits_pool": lambda: case_waits("waits_pool", flags=1 | 4),
}


def main():
    args = sys.argv[1:]
    if len(args) == 2 and args[0] == "--case":
        ok = CASES[args[1]]()
        sys.exit(0 if ok else 1)
    names = args or list(CASES)
    t0 = time.time()
    for nme in names:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", nme], timeout=240,
                               capture_output=True, text=True)
            sys.stdout.write(r.stdout)
            if r.returncode != 0:
                tail = (r.stderr or "").strip().splitlines()[-6:]
                print(f"CASE {nme}: EXIT {r.returncode}\n   " + "\n   ".join(tail), flush=True)
        except subprocess.TimeoutExpired:
            print(f"CASE {nme}: TIMEOUT", flush=True)
    print(f"probe done in {time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
