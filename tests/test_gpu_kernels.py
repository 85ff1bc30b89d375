"""GPU parity of every kernel behind the C ABI against its torch.nn.functional counterpart
(fp32 reference on the same bf16-rounded inputs).  Tolerances: 4e-3 relative L2 for bf16 outputs
(one bf16 rounding of the result), bit-exact for the fp32 elementwise sampler/perturb kernels."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests._util import T, load, rel_l2

pytestmark = pytest.mark.gpu
BF16_TOL = 4e-3


@pytest.fixture(scope="module")
def L():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from camels_diffusion_model_b200 import _lib
    assert _lib.lib().cdm_device_ok() == 0, _lib.lib().cdm_last_error()
    return _lib


def _conv_ref(x, w, scale, shift):
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), padding=1)
    y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    return y.permute(0, 2, 3, 1).contiguous()


def _conv_inputs(n, H, cin, cout, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, H, H, cin, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, 3, 3, cin, device="cuda", generator=g) / (3 * cin ** 0.5)).to(torch.bfloat16)
    scale = torch.rand(cout, device="cuda", generator=g) + 0.5
    shift = torch.randn(cout, device="cuda", generator=g) * 0.1
    return x, w, scale, shift


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("n,H,c0,c1,cout", [(3, 32, 128, 0, 128), (2, 64, 64, 64, 256), (1, 16, 256, 0, 128)])
def test_conv3x3_modes(L, mode, n, H, c0, c1, cout):
    x, w, scale, shift = _conv_inputs(n, H, c0 + c1, cout)
    out = torch.full((n, H, H, cout), float("nan"), device="cuda").to(torch.bfloat16)
    s1 = x[..., c0:].contiguous() if c1 else None
    L.conv3x3(x[..., :c0].contiguous(), w, scale, shift, out, src1=s1, flags=L.EPI_RELU, mode=mode)
    assert rel_l2(out.float(), _conv_ref(x, w, scale, shift).clamp_min(0)) < BF16_TOL


@pytest.mark.parametrize("mode", [2, 3, 4])
def test_conv3x3_many_units_persistent(L, mode):
    """More work units than SMs x 2: exercises the persistent loop, ring wrap-around and TMEM double buffering."""
    x, w, scale, shift = _conv_inputs(40, 64, 128, 128, seed=3)
    out = torch.empty(40, 64, 64, 128, device="cuda", dtype=torch.bfloat16)
    L.conv3x3(x, w, scale, shift, out, mode=mode)
    assert rel_l2(out.float(), _conv_ref(x, w, scale, shift).clamp_min(0)) < BF16_TOL
    out2 = torch.empty_like(out)
    L.conv3x3(x, w, scale, shift, out2, mode=mode)
    assert torch.equal(out, out2), "conv3x3 must be deterministic"


@pytest.mark.parametrize("mode", [2, 3, 4])
def test_conv3x3_pool(L, mode):
    x, w, scale, shift = _conv_inputs(4, 32, 128, 256, seed=1)
    out = torch.empty(4, 16, 16, 256, device="cuda", dtype=torch.bfloat16)
    L.conv3x3(x, w, scale, shift, out, flags=L.EPI_RELU | L.EPI_POOL, mode=mode)
    ref = F.max_pool2d(_conv_ref(x, w, scale, shift).clamp_min(0).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert rel_l2(out.float(), ref) < BF16_TOL


@pytest.mark.parametrize("mode", [2, 3, 4])
@pytest.mark.parametrize("rows", [1, 4])
def test_conv3x3_film(L, rows, mode):
    n, cout = 4, 128
    x, w, scale, shift = _conv_inputs(n, 32, 128, cout, seed=2)
    fs = torch.randn(n, cout, device="cuda")
    fsh = torch.randn(3, rows, cout, device="cuda")
    step = torch.tensor([2], device="cuda", dtype=torch.int32)
    out = torch.empty(n, 32, 32, cout, device="cuda", dtype=torch.bfloat16)
    L.conv3x3(x, w, scale, shift, out, flags=L.EPI_RELU | L.EPI_FILM, film_scale=fs, film_shift=fsh,
              film_shift_rows=rows, step_ptr=step, mode=mode)
    sh = fsh[2].expand(n, cout) if rows == 1 else fsh[2]
    ref = _conv_ref(x, w, scale, shift).clamp_min(0) * fs.view(n, 1, 1, cout) + sh.reshape(n, 1, 1, cout)
    assert rel_l2(out.float(), ref) < BF16_TOL


@pytest.mark.parametrize("mode", [2, 3, 4])
@pytest.mark.parametrize("reps", [1, 2])
def test_conv3x3_shortcut_fanout(L, reps, mode):
    n, H, cout = 3, 64, 128
    x, w, scale, shift = _conv_inputs(n, H, 128, cout, seed=4)
    xs = torch.randn(n, H, H, device="cuda")
    tab = torch.rand(5, reps, 2, cout, device="cuda") * 2 - 1
    step = torch.tensor([3], device="cuda", dtype=torch.int32)
    out = torch.empty(reps * n, H, H, cout, device="cuda", dtype=torch.bfloat16)
    L.conv3x3(x, w, scale, shift, out, flags=L.EPI_RELU | L.EPI_SHORTCUT, sc_x=xs, sc_tab=tab, sc_reps=reps,
              step_ptr=step, mode=mode)
    base = _conv_ref(x, w, scale, shift).clamp_min(0)
    for r in range(reps):
        ref = base + xs.view(n, H, H, 1) * tab[3, r, 0].view(1, 1, 1, cout) + tab[3, r, 1].view(1, 1, 1, cout)
        assert rel_l2(out[r * n:(r + 1) * n].float(), ref) < BF16_TOL


@pytest.mark.parametrize("flags_kind", ["plain", "c256", "out0", "shortcut2", "small_batch"])
def test_conv3x3_tma_store_epilogue_is_bit_identical(L, flags_kind):
    """MODE 4 (staging + cp.async.bulk.tensor stores) writes exactly what MODE 3 (register-direct stores) writes."""
    n, H, c0, c1, cout, kw, reps = {"plain": (5, 64, 128, 0, 128, {}, 1), "c256": (5, 32, 256, 0, 256, {}, 1),
                                    "out0": (3, 64, 128, 128, 128, {}, 1), "small_batch": (1, 64, 128, 0, 128, {}, 1),
                                    "shortcut2": (3, 64, 128, 0, 128, None, 2)}[flags_kind]
    x, w, scale, shift = _conv_inputs(n, H, c0 + c1, cout, seed=11)
    flags = L.EPI_RELU
    if kw is None:
        flags |= L.EPI_SHORTCUT
        kw = dict(sc_x=torch.randn(n, H, H, device="cuda"), sc_tab=torch.rand(2, 2, 2, cout, device="cuda") * 2 - 1,
                  sc_reps=2, step_ptr=torch.tensor([1], device="cuda", dtype=torch.int32))
    outs = []
    for mode in (3, 4):
        out = torch.full((reps * n, H, H, cout), float("nan"), device="cuda").to(torch.bfloat16)
        L.conv3x3(x[..., :c0].contiguous(), w, scale, shift, out, src1=x[..., c0:].contiguous() if c1 else None,
                  flags=flags, mode=mode, **kw)
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    assert bool(torch.isfinite(outs[1].float()).all())


@pytest.mark.parametrize("mode", [2, 3, 4])
def test_conv3x3_gelu_and_res_scale(L, mode):
    """The activation / residual-scale parameters north_star lists (GELU, /1.414): off by default (the reference
    runs ReLU and has the scale commented out, diffusion_utilities.py:29,36,59), parity-tested here."""
    n, H, cout = 3, 64, 128
    x, w, scale, shift = _conv_inputs(n, H, 128, cout, seed=6)
    xs = torch.randn(n, H, H, device="cuda")
    tab = torch.rand(1, 1, 2, cout, device="cuda") * 2 - 1
    out = torch.empty(n, H, H, cout, device="cuda", dtype=torch.bfloat16)
    L.conv3x3(x, w, scale, shift, out, flags=L.EPI_GELU | L.EPI_SHORTCUT | L.EPI_RESSCALE, sc_x=xs, sc_tab=tab,
              sc_reps=1, res_scale=1 / 1.414, mode=mode)
    ref = F.gelu(_conv_ref(x, w, scale, shift))
    ref = (ref + xs.view(n, H, H, 1) * tab[0, 0, 0].view(1, 1, 1, cout) + tab[0, 0, 1].view(1, 1, 1, cout)) / 1.414
    assert rel_l2(out.float(), ref) < BF16_TOL


def test_conv3x3_rejects_unknown_flag_bits(L):
    """Bits outside CDM_EPI_ALL (e.g. the measurement probes of the -DCDM_PROBES build) are an argument error in the
    production library, never forwarded to the kernel."""
    x, w, scale, shift = _conv_inputs(1, 32, 128, 128)
    out = torch.empty(1, 32, 32, 128, device="cuda", dtype=torch.bfloat16)
    for bit in (8, 26, 27, 28, 29, 30):
        with pytest.raises(L.CdmError):
            L.conv3x3(x, w, scale, shift, out, flags=L.EPI_RELU | (1 << bit))


@pytest.mark.parametrize("mode", [2, 3, 4])
def test_conv3x3_gnstats(L, mode):
    n, H = 3, 64
    x, w, scale, shift = _conv_inputs(n, H, 256, 128, seed=5)
    part = torch.full((n, (H // 16) ** 2 * 8, 8, 2), float("nan"), device="cuda")  # every slot must be written
    out = torch.empty(n, H, H, 128, device="cuda", dtype=torch.bfloat16)
    L.conv3x3(x[..., :128].contiguous(), w, scale, shift, out, src1=x[..., 128:].contiguous(), flags=L.EPI_GNSTATS,
              gn_partial=part, mode=mode)
    ref = _conv_ref(x, w, scale, shift)
    assert rel_l2(out.float(), ref) < BF16_TOL
    mr = torch.empty(n, 8, 2, device="cuda")
    L.gn_finalize(part, 16.0 * H * H, mr)
    g = ref.view(n, H * H, 8, 16)
    mean = g.mean((1, 3))
    var = g.var((1, 3), unbiased=False)
    assert rel_l2(mr[..., 0], mean) < 1e-4
    assert rel_l2(mr[..., 1], torch.rsqrt(var + 1e-5)) < 1e-4


@pytest.mark.parametrize("M,k0,k1,N", [(256, 64, 0, 128), (300, 128, 64, 256), (5, 256, 0, 1024)])
def test_gemm_plain(L, M, k0, k1, N):
    g = torch.Generator(device="cuda").manual_seed(0)
    K = k0 + k1
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    bw = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    shift = torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda").to(torch.bfloat16)
    L.gemm(a[:, :k0].contiguous(), bw, shift, out, a1=a[:, k0:].contiguous() if k1 else None)
    assert rel_l2(out.float(), a.float() @ bw.float().t() + shift) < BF16_TOL


@pytest.mark.parametrize("M,k0,k1,N,shift_mod", [(2048, 256, 0, 65536, 256), (1100, 128, 128, 76800, 128),
                                                 (1024, 64, 0, 38912, 256)])
def test_gemm_resident_weights_many_groups(L, M, k0, k1, N, shift_mod):
    """up0 at sampling batch sizes: more two-tile weight groups than CTAs (256 / 300 / 152 groups = 2 / 3 / 2 per
    CTA), weights swapped between groups; ragged last row tile; bias indexed modulo shift_mod."""
    g = torch.Generator(device="cuda").manual_seed(0)
    K = k0 + k1
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    bw = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    shift = torch.randn(shift_mod, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda").to(torch.bfloat16)
    L.gemm(a[:, :k0].contiguous(), bw, shift, out, a1=a[:, k0:].contiguous() if k1 else None, shift_mod=shift_mod)
    ref = a.float() @ bw.float().t() + shift.repeat(N // shift_mod)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out.float(), ref) < BF16_TOL
    # the streaming kernel (taken below 1024 rows) accumulates in the same order: bit-identical rows
    out_small = torch.empty(512, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(a[:512, :k0].contiguous(), bw, shift, out_small, a1=a[:512, k0:].contiguous() if k1 else None,
           shift_mod=shift_mod)
    assert torch.equal(out[:512], out_small)


def test_gemm_pixel_shuffle_is_conv_transpose(L):
    """out_mode 1 == nn.ConvTranspose2d(cin, 128, 2, 2) on the channel concat of two NHWC sources."""
    g = torch.Generator(device="cuda").manual_seed(1)
    n, H, c0, c1 = 3, 16, 256, 256
    a = torch.randn(n, H, H, c0 + c1, device="cuda", generator=g).to(torch.bfloat16)
    wt = (torch.randn(c0 + c1, 128, 2, 2, device="cuda", generator=g) / 23.0).to(torch.bfloat16)  # IOHW
    bias = torch.randn(128, device="cuda", generator=g)
    bw = wt.permute(2, 3, 1, 0).reshape(4 * 128, c0 + c1).contiguous()
    out = torch.empty(n, 2 * H, 2 * H, 128, device="cuda", dtype=torch.bfloat16)
    L.gemm(a[..., :c0].reshape(-1, c0).contiguous(), bw, bias, out, a1=a[..., c0:].reshape(-1, c1).contiguous(),
           out_mode=1, H=H, W=H, shift_mod=128)
    ref = F.conv_transpose2d(a.float().permute(0, 3, 1, 2), wt.float(), bias, stride=2).permute(0, 2, 3, 1)
    assert rel_l2(out.float(), ref) < BF16_TOL


@pytest.mark.parametrize("n,H,c0,c1", [(8, 32, 128, 128), (40, 16, 256, 256), (9, 32, 256, 0)])
def test_gemm_resident_weights_path(L, n, H, c0, c1):
    """M >= 8192 rows and N = 512 take gemm_bres_kernel (weight tiles resident in shared memory)."""
    g = torch.Generator(device="cuda").manual_seed(11)
    a = torch.randn(n, H, H, c0 + c1, device="cuda", generator=g).to(torch.bfloat16)
    wt = (torch.randn(c0 + c1, 128, 2, 2, device="cuda", generator=g) / 20.0).to(torch.bfloat16)
    bias = torch.randn(128, device="cuda", generator=g)
    bw = wt.permute(2, 3, 1, 0).reshape(512, c0 + c1).contiguous()
    out = torch.full((n, 2 * H, 2 * H, 128), float("nan"), device="cuda").to(torch.bfloat16)
    L.gemm(a[..., :c0].reshape(-1, c0).contiguous(), bw, bias, out,
           a1=a[..., c0:].reshape(-1, c1).contiguous() if c1 else None, out_mode=1, H=H, W=H, shift_mod=128)
    ref = F.conv_transpose2d(a.float().permute(0, 3, 1, 2), wt.float(), bias, stride=2).permute(0, 2, 3, 1)
    assert rel_l2(out.float(), ref) < BF16_TOL
    # plain row-major output, ragged M, N = 256
    M = 128 * 64 + 37
    a2 = torch.randn(M, 512, device="cuda", generator=g).to(torch.bfloat16)
    b2 = (torch.randn(256, 512, device="cuda", generator=g) / 22.0).to(torch.bfloat16)
    sh = torch.randn(256, device="cuda", generator=g)
    o2 = torch.full((M, 256), float("nan"), device="cuda").to(torch.bfloat16)
    L.gemm(a2, b2, sh, o2)
    assert rel_l2(o2.float(), a2.float() @ b2.float().t() + sh) < BF16_TOL


@pytest.mark.parametrize("n,H,relu", [(5, 64, True), (1, 16, False), (300, 32, True)])
def test_conv_in(L, n, H, relu):
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(n, H, H, device="cuda", generator=g)
    w = torch.randn(128, 1, 3, 3, device="cuda", generator=g) / 3
    scale = torch.rand(128, device="cuda", generator=g) + 0.5
    shift = torch.randn(128, device="cuda", generator=g) * 0.1
    out = torch.full((n, H, H, 128), float("nan"), device="cuda").to(torch.bfloat16)
    L.conv_in(x, w.reshape(128, 9).t().contiguous(), scale, shift, out, relu=relu)
    ref = F.conv2d(x.unsqueeze(1), w, padding=1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    ref = (F.relu(ref) if relu else ref).permute(0, 2, 3, 1)
    assert rel_l2(out.float(), ref) < BF16_TOL
    # hi/lo-split operands keep fp32-level accuracy: the only rounding left is the bf16 output
    assert rel_l2(out.float(), ref.to(torch.bfloat16).float()) < 2e-4


@pytest.mark.parametrize("n", [3, 160])  # 160 images: the 32-row-tile instantiation; 3: the 8-row one
def test_conv_out_groupnorm_fused(L, n):
    g = torch.Generator(device="cuda").manual_seed(3)
    src = (torch.randn(n, 64, 64, 128, device="cuda", generator=g) * 2 + 0.3).to(torch.bfloat16)
    gamma = torch.rand(128, device="cuda", generator=g) + 0.5
    gamma[5] = -gamma[5]  # a negative affine scale must not be folded through the ReLU
    beta = torch.randn(128, device="cuda", generator=g) * 0.2
    w = torch.randn(1, 128, 3, 3, device="cuda", generator=g) / 30
    bias = torch.randn(1, device="cuda", generator=g)
    xs = src.float().permute(0, 3, 1, 2)
    gs = xs.reshape(n, 8, -1)
    mr = torch.stack([gs.mean(2), torch.rsqrt(gs.var(2, unbiased=False) + 1e-5)], -1).contiguous()
    out = torch.empty(n, 64, 64, device="cuda")
    wt = w[0].permute(1, 2, 0).reshape(9, 128).contiguous()
    L.conv_out(src, mr, gamma, beta, wt, bias, out)
    ref = F.conv2d(F.relu(F.group_norm(xs, 8, gamma, beta, 1e-5)), w, bias, padding=1)[:, 0]
    # the normalised activation and the weights enter the tensor core as bf16 (fp32 accumulate)
    assert rel_l2(out, ref) < BF16_TOL
    # tile height is a launch heuristic: results must not depend on it (batch-split invariance, bit for bit)
    out3 = torch.empty(2, 64, 64, device="cuda")
    L.conv_out(src[:2].contiguous(), mr[:2].contiguous(), gamma, beta, wt, bias, out3)
    assert torch.equal(out3, out[:2])


@pytest.mark.parametrize("din,emb,rows", [(1, 256, 1), (6, 128, 37), (3, 256, 1501)])
def test_embed_fc(L, din, emb, rows):
    g = torch.Generator(device="cuda").manual_seed(4)
    inp = torch.rand(rows, din, device="cuda", generator=g)
    w1, b1 = torch.randn(emb, din, device="cuda", generator=g), torch.randn(emb, device="cuda", generator=g)
    w2 = torch.randn(emb, emb, device="cuda", generator=g) / emb ** 0.5
    b2 = torch.randn(emb, device="cuda", generator=g)
    out = torch.empty(rows, emb, device="cuda")
    L.embed_fc(inp, w1, b1, w2, b2, out)
    assert rel_l2(out, F.linear(F.gelu(F.linear(inp, w1, b1)), w2, b2)) < 1e-5


# (256, 256) is up0's shape: the register-resident instantiation; the others take the generic three-pass one
@pytest.mark.parametrize("P,C", [(256, 256), (100, 256), (256, 128)])
def test_avgpool_gelu_and_gn_relu_film(L, P, C):
    g = torch.Generator(device="cuda").manual_seed(5)
    n = 5
    src = torch.randn(n, P, C, device="cuda", generator=g).to(torch.bfloat16)
    hid = torch.empty(n, C, device="cuda", dtype=torch.bfloat16)
    L.avgpool_gelu(src, hid)
    assert rel_l2(hid.float(), F.gelu(src.float().mean(1))) < BF16_TOL
    gamma = torch.rand(C, device="cuda", generator=g) + 0.5
    beta = torch.randn(C, device="cuda", generator=g) * 0.2
    fs = torch.randn(n, C, device="cuda", generator=g)
    fb = torch.randn(4, 1, C, device="cuda", generator=g)
    step = torch.tensor([1], device="cuda", dtype=torch.int32)
    out = torch.empty(n, P, C, device="cuda", dtype=torch.bfloat16)
    L.gn_relu_film(src, gamma, beta, out, film_scale=fs, film_shift=fb, film_rows=1, step_ptr=step)
    y = F.relu(F.group_norm(src.float().permute(0, 2, 1), 8, gamma, beta, 1e-5)).permute(0, 2, 1)
    assert rel_l2(out.float(), y * fs.view(n, 1, C) + fb[1]) < BF16_TOL


def test_elementwise_bit_exact_vs_reference_vectors(L):
    """ddpm_step / perturb against vectors produced by the reference's own closures (sampler.npz)."""
    import camels_diffusion_model_b200 as cdm
    g = load("sampler.npz")
    b_t, a_t, ab_t = (T(g["sched/" + k]).cuda() for k in ("b_t", "a_t", "ab_t"))
    x, n, z, e, t = (T(g["ew/" + k]) for k in ("x", "noise", "z", "eps", "t"))
    assert np.array_equal(cdm.perturb_input(x, t, n, ab_t).cpu().numpy(), g["ew/perturb_vec"])
    assert np.array_equal(cdm.perturb_input(x, 5, n, ab_t).cpu().numpy(), g["ew/perturb_scalar"])
    assert np.array_equal(cdm.denoise_add_noise(x, 7, e, z, b_t, a_t, ab_t).cpu().numpy(), g["ew/denoise_t7"])
    assert np.array_equal(cdm.denoise_add_noise(x, 1, e, 0, b_t, a_t, ab_t).cpu().numpy(), g["ew/denoise_t1"])


def test_ddpm_step_cfg_mix_and_snapshot(L):
    from camels_diffusion_model_b200.diffusion import _coef_table, make_schedule
    Tn = 50
    b_t, a_t, ab_t = make_schedule(Tn)
    coef = _coef_table(b_t, a_t, ab_t)
    g = torch.Generator(device="cuda").manual_seed(6)
    n = 4
    x = torch.randn(n, 1, 64, 64, device="cuda", generator=g)
    eps = torch.randn(2 * n, 1, 64, 64, device="cuda", generator=g)
    z = torch.randn(Tn, n * 4096, device="cuda", generator=g)
    step = torch.tensor([17], device="cuda", dtype=torch.int32)
    slot = torch.full((Tn + 1,), -1, dtype=torch.int32, device="cuda")
    slot[17] = 2
    snap = torch.zeros(3, n, 1, 64, 64, device="cuda")
    x_ref = x.clone()
    L.ddpm_step(x, eps, coef, Tn, reps=2, guide_w=1.5, step_ptr=step, z=z, z_iter_stride=n * 4096, snap=snap,
                snap_slot=slot)
    e = eps[n:] + 1.5 * (eps[:n] - eps[n:])
    zz = z[Tn - 17].view(n, 1, 64, 64)
    ref = (x_ref - e * ((1 - a_t[17]) / (1 - ab_t[17]).sqrt())) / a_t[17].sqrt() + b_t.sqrt()[17] * zz
    assert torch.equal(x, ref)
    assert torch.equal(snap[2], x) and float(snap[0].abs().max()) == 0.0
    L.step_advance(step, -1)
    assert int(step.item()) == 16


def test_philox_normal_statistics(L):
    from camels_diffusion_model_b200.diffusion import _coef_table, make_schedule
    Tn = 10
    coef = _coef_table(*make_schedule(Tn))
    coef[:, 0] = 0   # x <- x / sa + sb * z with x = 0  => z * sb
    n = 64
    x = torch.zeros(n, 1, 64, 64, device="cuda")
    eps = torch.zeros(n, 1, 64, 64, device="cuda")
    L.ddpm_step(x, eps, coef, Tn, step=5, seed=1234)
    zs = (x / coef[5, 2]).flatten().double()
    assert abs(zs.mean()) < 5e-3 and abs(zs.var() - 1) < 1e-2
    assert abs((zs ** 4).mean() - 3) < 5e-2
    x2 = torch.zeros_like(x)
    L.ddpm_step(x2, eps, coef, Tn, step=6, seed=1234)
    assert abs(torch.corrcoef(torch.stack([x.flatten(), x2.flatten()]))[0, 1]) < 5e-3  # steps decorrelated


def test_mse_accum(L):
    g = torch.Generator(device="cuda").manual_seed(7)
    n = 7
    p, t = torch.randn(n, 1, 64, 64, device="cuda", generator=g), torch.randn(n, 1, 64, 64, device="cuda", generator=g)
    w = torch.rand(11, device="cuda", generator=g)
    ti = torch.randint(1, 11, (n,), device="cuda", generator=g)
    mse = torch.zeros(n, device="cuda")
    acc = torch.ones(n, device="cuda")
    L.mse_accum(p, t, weight_tab=w, t_idx=ti, mse_out=mse, acc=acc)
    ref = ((p - t) ** 2).mean((1, 2, 3))
    assert rel_l2(mse, ref) < 1e-6
    assert rel_l2(acc, 1 + w[ti] * ref) < 1e-6


def test_bad_arguments_raise(L):
    import camels_diffusion_model_b200 as cdm
    x, w, scale, shift = _conv_inputs(1, 16, 96, 128)   # 96 channels: not a multiple of 64
    out = torch.empty(1, 16, 16, 128, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(cdm.CdmError):
        L.conv3x3(x, w, scale, shift, out)


def test_gemm_split_k_small_output_long_k(L):
    """Two output tiles and K = 65536 (the up0 data gradient): the K range is split over the idle SMs through a
    workspace; result == the unsplit kernel's up to fp32 summation order, and bit-reproducible."""
    g = torch.Generator(device="cuda").manual_seed(21)
    for M in (32, 200):
        a = (torch.randn(M, 65536, device="cuda", generator=g) / 16).to(torch.bfloat16)
        b = (torch.randn(256, 65536, device="cuda", generator=g) / 16).to(torch.bfloat16)
        sh = torch.randn(256, device="cuda", generator=g)
        ws = torch.empty(160 * 128 * 384, device="cuda")
        o_split = torch.full((M, 256), float("nan"), device="cuda").to(torch.bfloat16)
        L.gemm(a, b, sh, o_split, workspace=ws)
        o_plain = torch.full((M, 256), float("nan"), device="cuda").to(torch.bfloat16)
        L.gemm(a, b, sh, o_plain)
        ref = a.float() @ b.float().t() + sh
        assert rel_l2(o_split.float(), ref) < BF16_TOL and rel_l2(o_plain.float(), ref) < BF16_TOL
        o2 = torch.empty_like(o_split)
        L.gemm(a, b, sh, o2, workspace=ws)
        assert torch.equal(o_split, o2)
