"""ctypes binding of libcdm_b200.so (declared in include/cdm_b200.h).

There is no fallback: if the shared library is missing or the device is not an
sm_100 GPU every op raises.  PyTorch is only used for device memory and streams.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CDM_LIB selects another build of the same ABI (tools/gpu_probe.py loads the -DCDM_PROBES library this way)
LIB_PATH = os.environ.get("CDM_LIB") or os.path.join(_HERE, "libcdm_b200.so")

EPI_RELU, EPI_SHORTCUT, EPI_POOL, EPI_FILM, EPI_GNSTATS, EPI_BNSTATS, EPI_GELU, EPI_RESSCALE, EPI_BNBWD = (
    1, 2, 4, 8, 16, 32, 64, 128, 256)
CONV_MODE_COPIES, CONV_MODE_SHIFT24, CONV_MODE_SHIFT18, CONV_MODE_SWAPPED, CONV_MODE_SWAPPED_TMA = 0, 1, 2, 3, 4


class CdmError(RuntimeError):
    pass


class Conv3x3Args(C.Structure):
    _fields_ = [
        ("src0", C.c_void_p), ("src1", C.c_void_p), ("c0", C.c_int), ("c1", C.c_int),
        ("n_img", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("weight", C.c_void_p), ("cout", C.c_int),
        ("scale", C.c_void_p), ("shift", C.c_void_p), ("flags", C.c_int), ("out", C.c_void_p),
        ("sc_x", C.c_void_p), ("sc_reps", C.c_int), ("sc_tab", C.c_void_p),
        ("film_scale", C.c_void_p), ("film_shift", C.c_void_p), ("film_shift_rows", C.c_int),
        ("step_ptr", C.c_void_p), ("gn_partial", C.c_void_p), ("mode", C.c_int),
        ("bn_partial", C.c_void_p), ("bn_sums", C.c_void_p), ("xr", C.c_void_p), ("res_scale", C.c_float),
        ("bwd_z", C.c_void_p), ("bwd_scale", C.c_void_p), ("bwd_shift", C.c_void_p), ("bwd_mean", C.c_void_p),
        ("bwd_rstd", C.c_void_p),
    ]


class XrankArgs(C.Structure):
    """cdm_xrank: peer group of the fused reduce + cross-rank exchange (built by parallel.PeerExchange)."""
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("peer_slots", C.c_void_p), ("peer_flags", C.c_void_p),
                ("seq", C.c_void_p), ("ticket", C.c_void_p)]


def _xr_ptr(xr):
    return None if xr is None else C.cast(C.pointer(xr), C.c_void_p)


class GemmArgs(C.Structure):
    _fields_ = [
        ("a0", C.c_void_p), ("a1", C.c_void_p), ("k0", C.c_int), ("k1", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("bw", C.c_void_p), ("shift", C.c_void_p),
        ("shift_mod", C.c_int), ("out_mode", C.c_int), ("H", C.c_int), ("W", C.c_int), ("out", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_floats", C.c_longlong),
    ]


class ConvInArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("n_img", C.c_int), ("H", C.c_int), ("W", C.c_int), ("weight", C.c_void_p),
                ("cout", C.c_int), ("scale", C.c_void_p), ("shift", C.c_void_p), ("relu", C.c_int),
                ("out", C.c_void_p)]


class ConvOutArgs(C.Structure):
    _fields_ = [("src", C.c_void_p), ("n_img", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int),
                ("mean_rstd", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("weight", C.c_void_p),
                ("bias", C.c_void_p), ("out", C.c_void_p)]


class GnReluFilmArgs(C.Structure):
    _fields_ = [("src", C.c_void_p), ("n_img", C.c_int), ("P", C.c_int), ("C", C.c_int), ("groups", C.c_int),
                ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float), ("film_scale", C.c_void_p),
                ("film_shift", C.c_void_p), ("film_rows", C.c_int), ("step_ptr", C.c_void_p), ("out", C.c_void_p),
                ("mean_rstd_out", C.c_void_p)]


class DdpmStepArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("eps", C.c_void_p), ("n", C.c_int), ("hw", C.c_int), ("reps", C.c_int),
                ("guide_w", C.c_float), ("coef", C.c_void_p), ("step_ptr", C.c_void_p), ("step", C.c_int),
                ("timesteps", C.c_int), ("z", C.c_void_p), ("z_iter_stride", C.c_longlong),
                ("seed", C.c_ulonglong), ("snap", C.c_void_p), ("snap_slot", C.c_void_p),
                ("sample_offset", C.c_longlong)]


class PerturbArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("noise", C.c_void_p), ("out", C.c_void_p), ("n", C.c_int), ("hw", C.c_int),
                ("ca", C.c_void_p), ("cb", C.c_void_p), ("t_idx", C.c_void_p), ("t_shared", C.c_int),
                ("step_ptr", C.c_void_p), ("seed", C.c_ulonglong), ("stream_id", C.c_uint), ("noise_out", C.c_void_p),
                ("sample_offset", C.c_longlong)]


class MseAccumArgs(C.Structure):
    _fields_ = [("pred", C.c_void_p), ("target", C.c_void_p), ("n", C.c_int), ("hw", C.c_int),
                ("weight_tab", C.c_void_p), ("t_idx", C.c_void_p), ("t_shared", C.c_int),
                ("step_ptr", C.c_void_p), ("mse_out", C.c_void_p), ("acc", C.c_void_p),
                ("weight_tab2", C.c_void_p), ("acc2", C.c_void_p)]


class PlanDesc(C.Structure):
    _fields_ = [("n_cfeat", C.c_int), ("batch", C.c_int), ("reps", C.c_int), ("tensors", C.POINTER(C.c_void_p)),
                ("arena", C.c_void_p), ("arena_bytes", C.c_longlong), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_longlong), ("conv_mode", C.c_int)]


class ForwardArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("sc_tab", C.c_void_p), ("cemb1", C.c_void_p), ("temb1", C.c_void_p),
                ("cemb2", C.c_void_p), ("temb2", C.c_void_p), ("temb_rows", C.c_int), ("step_ptr", C.c_void_p),
                ("eps", C.c_void_p)]


class SampleStepArgs(C.Structure):
    _fields_ = [("fwd", ForwardArgs), ("x", C.c_void_p), ("step_ptr", C.c_void_p), ("guide_w", C.c_float),
                ("coef", C.c_void_p), ("timesteps", C.c_int), ("z", C.c_void_p), ("z_iter_stride", C.c_longlong),
                ("seed", C.c_ulonglong), ("sample_offset", C.c_longlong), ("snap", C.c_void_p),
                ("snap_slot", C.c_void_p)]


_lib = None


def lib():
    """Load (once) and return the ctypes handle. Raises CdmError if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CdmError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a). There is no CPU/PyTorch fallback for the hot path.")
    l = C.CDLL(LIB_PATH)
    l.cdm_last_error.restype = C.c_char_p
    l.cdm_plan_tensor_name.restype = C.c_char_p
    l.cdm_plan_launch_name.restype = C.c_char_p
    for fn in ("cdm_plan_tensor_numel", "cdm_plan_arena_bytes", "cdm_plan_workspace_bytes"):
        getattr(l, fn).restype = C.c_longlong
    l.cdm_plan_destroy.restype = None
    l.cdm_plan_destroy.argtypes = [C.c_void_p]
    for name in EXPORTS:
        getattr(l, name)  # raises AttributeError if the header and the library disagree
    _lib = l
    return l


# every symbol include/cdm_b200.h declares
EXPORTS = [
    "cdm_version", "cdm_last_error", "cdm_device_ok", "cdm_num_sms",
    "cdm_conv3x3", "cdm_gemm", "cdm_probe_tma_l2",
    "cdm_conv_in", "cdm_conv_out", "cdm_embed_fc", "cdm_avgpool_gelu", "cdm_gn_relu_film", "cdm_gn_finalize",
    "cdm_ddpm_step", "cdm_step_advance", "cdm_perturb", "cdm_mse_accum",
    "cdm_gemm_tn", "cdm_chan_reduce", "cdm_bn_finalize", "cdm_bn_apply", "cdm_bn_bwd_apply", "cdm_maxpool2_fwd",
    "cdm_maxpool2_bwd", "cdm_add_bf16", "cdm_space_to_depth", "cdm_film_bwd", "cdm_gn_bwd", "cdm_rows_sum",
    "cdm_avgpool_gelu_train", "cdm_avgpool_gelu_bwd", "cdm_outer_wgrad", "cdm_embed_bwd", "cdm_mse_grad",
    "cdm_adam_step", "cdm_power_spectrum", "cdm_pixel_histogram",
    "cdm_minmax", "cdm_preprocess_maps", "cdm_normalize_params", "cdm_xrank_sum", "cdm_xrank_set_timeout", "cdm_pack_bf16", "cdm_pack_transpose_bf16",
    "cdm_plan_n_tensors", "cdm_plan_tensor_name", "cdm_plan_tensor_numel", "cdm_plan_arena_bytes",
    "cdm_plan_workspace_bytes", "cdm_plan_create", "cdm_plan_refresh", "cdm_plan_destroy", "cdm_plan_buffer",
    "cdm_plan_embed", "cdm_forward_eval", "cdm_sample_step", "cdm_plan_n_launches", "cdm_plan_launch_name",
    "cdm_plan_profile",
]


def check(rc, what):
    if rc != 0:
        msg = lib().cdm_last_error().decode()
        raise CdmError(f"{what} failed (status {rc}): {msg}")


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device-resident contiguous tensors only"
    return t.data_ptr()


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def on_device(fn):
    """Decorator for methods of objects with a `.dev` / `.device` attribute (or a ContextUnet): run the body with that
    device current, so that stream_ptr(), the kernels' attribute caches and every allocation refer to the device the
    tensors live on even when the caller's current device is another GPU."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        dev = getattr(self, "dev", None) or getattr(self, "device", None)
        if dev is None and hasattr(self, "_check_supported"):
            dev = self._check_supported()
        if dev is None or torch.device(dev).type != "cuda":
            return fn(self, *a, **k)
        with torch.cuda.device(dev):
            return fn(self, *a, **k)
    return wrapper


_SMS = {}


def num_sms():
    """Multiprocessor count of the current device (sizes the per-CTA partial-sum workspaces)."""
    d = torch.cuda.current_device()
    if d not in _SMS:
        n = lib().cdm_num_sms()
        if n <= 0:
            raise CdmError("cdm_num_sms failed: " + lib().cdm_last_error().decode())
        _SMS[d] = n
    return _SMS[d]


def conv3x3(src0, weight, scale, shift, out, *, src1=None, flags=EPI_RELU, sc_x=None, sc_tab=None, sc_reps=1,
            film_scale=None, film_shift=None, film_shift_rows=1, step_ptr=None, gn_partial=None,
            mode=CONV_MODE_SWAPPED, bn_partial=None, bn_sums=None, xr=None, res_scale=1.0, bwd=None):
    """src*: bf16 [n,H,W,c]; weight bf16 [cout,3,3,cin]; out bf16 NHWC. See cdm_conv3x3 in cdm_b200.h."""
    n, H, W, c0 = src0.shape
    a = Conv3x3Args()
    a.src0, a.c0 = ptr(src0), c0
    a.src1, a.c1 = (ptr(src1), src1.shape[3]) if src1 is not None else (None, 0)
    a.n_img, a.H, a.W = n, H, W
    a.weight, a.cout = ptr(weight), weight.shape[0]
    assert weight.shape[3] == a.c0 + a.c1
    a.scale, a.shift, a.flags, a.out = ptr(scale), ptr(shift), flags, ptr(out)
    a.sc_x, a.sc_reps, a.sc_tab = ptr(sc_x), sc_reps, ptr(sc_tab)
    a.film_scale, a.film_shift, a.film_shift_rows = ptr(film_scale), ptr(film_shift), film_shift_rows
    a.step_ptr, a.gn_partial, a.mode = ptr(step_ptr), ptr(gn_partial), mode
    a.bn_partial, a.bn_sums, a.xr, a.res_scale = ptr(bn_partial), ptr(bn_sums), _xr_ptr(xr), res_scale
    if bwd is not None:  # EPI_BNBWD: (z, scale, shift, mean, rstd) of the layer whose dy this launch produces
        a.bwd_z, a.bwd_scale, a.bwd_shift, a.bwd_mean, a.bwd_rstd = (ptr(t) for t in bwd)
    check(lib().cdm_conv3x3(C.byref(a), stream_ptr()), "cdm_conv3x3")
    return out


def gemm(a0, bw, shift, out, *, a1=None, shift_mod=None, out_mode=0, H=0, W=0, workspace=None):
    """a*: bf16 [M,k]; bw bf16 [N,K]; see cdm_gemm in cdm_b200.h."""
    g = GemmArgs()
    g.a0, g.k0 = ptr(a0), a0.shape[1]
    g.a1, g.k1 = (ptr(a1), a1.shape[1]) if a1 is not None else (None, 0)
    g.M, g.N = a0.shape[0], bw.shape[0]
    assert bw.shape[1] == g.k0 + g.k1
    g.bw, g.shift = ptr(bw), ptr(shift)
    g.shift_mod = shift_mod if shift_mod is not None else shift.numel()
    g.out_mode, g.H, g.W, g.out = out_mode, H, W, ptr(out)
    g.workspace, g.workspace_floats = ptr(workspace), 0 if workspace is None else workspace.numel()
    check(lib().cdm_gemm(C.byref(g), stream_ptr()), "cdm_gemm")
    return out


def probe_tma_l2(buf, n_rows, iters):
    check(lib().cdm_probe_tma_l2(C.c_void_p(ptr(buf)), n_rows, iters, stream_ptr()), "cdm_probe_tma_l2")


def conv_in(x, weight, scale, shift, out, relu=True):
    """x fp32 [n,H,W]; weight fp32 [9,cout]; out bf16 [n,H,W,cout]."""
    a = ConvInArgs()
    a.x, (a.n_img, a.H, a.W) = ptr(x), x.shape
    a.weight, a.cout, a.scale, a.shift, a.relu, a.out = ptr(weight), weight.shape[1], ptr(scale), ptr(shift), int(relu), ptr(out)
    check(lib().cdm_conv_in(C.byref(a), stream_ptr()), "cdm_conv_in")
    return out


def conv_out(src, mean_rstd, gamma, beta, weight, bias, out):
    """src bf16 [n,H,W,128]; weight fp32 [9,128]; out fp32 [n,H,W]."""
    a = ConvOutArgs()
    a.src, (a.n_img, a.H, a.W, a.C) = ptr(src), src.shape
    a.mean_rstd, a.gamma, a.beta, a.weight, a.bias, a.out = (ptr(mean_rstd), ptr(gamma), ptr(beta), ptr(weight),
                                                             ptr(bias), ptr(out))
    check(lib().cdm_conv_out(C.byref(a), stream_ptr()), "cdm_conv_out")
    return out


def embed_fc(inp, w1, b1, w2, b2, out):
    """inp fp32 [rows,din]; w1 [emb,din]; w2 [emb,emb]; out fp32 [rows,emb]."""
    rows, din = inp.shape
    check(lib().cdm_embed_fc(C.c_void_p(ptr(inp)), rows, din, C.c_void_p(ptr(w1)), C.c_void_p(ptr(b1)),
                             C.c_void_p(ptr(w2)), C.c_void_p(ptr(b2)), w2.shape[0], C.c_void_p(ptr(out)),
                             stream_ptr()), "cdm_embed_fc")
    return out


def avgpool_gelu(src, out):
    """src bf16 [n,P,C] -> out bf16 [n,C]."""
    n, P, Cc = src.shape
    check(lib().cdm_avgpool_gelu(C.c_void_p(ptr(src)), n, P, Cc, C.c_void_p(ptr(out)), stream_ptr()),
          "cdm_avgpool_gelu")
    return out


def gn_relu_film(src, gamma, beta, out, *, groups=8, eps=1e-5, film_scale=None, film_shift=None, film_rows=1,
                 step_ptr=None, mean_rstd_out=None):
    """src/out bf16 [n,P,C]."""
    a = GnReluFilmArgs()
    a.src, (a.n_img, a.P, a.C) = ptr(src), src.shape
    a.groups, a.gamma, a.beta, a.eps = groups, ptr(gamma), ptr(beta), eps
    a.film_scale, a.film_shift, a.film_rows, a.step_ptr, a.out = (ptr(film_scale), ptr(film_shift), film_rows,
                                                                  ptr(step_ptr), ptr(out))
    a.mean_rstd_out = ptr(mean_rstd_out)
    check(lib().cdm_gn_relu_film(C.byref(a), stream_ptr()), "cdm_gn_relu_film")
    return out


def gn_finalize(partial, count, mean_rstd, eps=1e-5):
    """partial fp32 [n,slots,8,2] -> mean_rstd fp32 [n,8,2]."""
    n, slots = partial.shape[0], partial.shape[1]
    check(lib().cdm_gn_finalize(C.c_void_p(ptr(partial)), n, slots, C.c_float(count), C.c_float(eps),
                                C.c_void_p(ptr(mean_rstd)), stream_ptr()), "cdm_gn_finalize")
    return mean_rstd


def ddpm_step(x, eps, coef, timesteps, *, reps=1, guide_w=0.0, step=0, step_ptr=None, z=None, z_iter_stride=0,
              seed=0, snap=None, snap_slot=None, sample_offset=0):
    """x fp32 [n,...] in place; eps fp32 [reps*n,...]; coef fp32 [T+1,4]."""
    a = DdpmStepArgs()
    n = x.shape[0]
    a.x, a.eps, a.n, a.hw, a.reps, a.guide_w = ptr(x), ptr(eps), n, x.numel() // n, reps, guide_w
    a.coef, a.step_ptr, a.step, a.timesteps = ptr(coef), ptr(step_ptr), step, timesteps
    a.z, a.z_iter_stride, a.seed, a.snap, a.snap_slot = ptr(z), z_iter_stride, seed, ptr(snap), ptr(snap_slot)
    a.sample_offset = sample_offset
    check(lib().cdm_ddpm_step(C.byref(a), stream_ptr()), "cdm_ddpm_step")
    return x


def step_advance(step_ptr, delta):
    check(lib().cdm_step_advance(C.c_void_p(ptr(step_ptr)), delta, stream_ptr()), "cdm_step_advance")


def perturb(x, out, ca, cb, *, noise=None, t_idx=None, t_shared=0, step_ptr=None, seed=0, stream_id=0,
            noise_out=None, sample_offset=0):
    a = PerturbArgs()
    n = x.shape[0]
    a.x, a.noise, a.out, a.n, a.hw = ptr(x), ptr(noise), ptr(out), n, x.numel() // n
    a.ca, a.cb, a.t_idx, a.t_shared, a.step_ptr = ptr(ca), ptr(cb), ptr(t_idx), t_shared, ptr(step_ptr)
    a.seed, a.stream_id, a.noise_out, a.sample_offset = seed, stream_id, ptr(noise_out), sample_offset
    check(lib().cdm_perturb(C.byref(a), stream_ptr()), "cdm_perturb")
    return out


def mse_accum(pred, target, *, weight_tab=None, t_idx=None, t_shared=0, step_ptr=None, mse_out=None, acc=None,
              weight_tab2=None, acc2=None):
    a = MseAccumArgs()
    n = pred.shape[0]
    a.pred, a.target, a.n, a.hw = ptr(pred), ptr(target), n, pred.numel() // n
    a.weight_tab, a.t_idx, a.t_shared, a.step_ptr = ptr(weight_tab), ptr(t_idx), t_shared, ptr(step_ptr)
    a.mse_out, a.acc, a.weight_tab2, a.acc2 = ptr(mse_out), ptr(acc), ptr(weight_tab2), ptr(acc2)
    check(lib().cdm_mse_accum(C.byref(a), stream_ptr()), "cdm_mse_accum")


# ----------------------------------------------------------------------------- composite plan API
def plan_tensor_names():
    l = lib()
    return [l.cdm_plan_tensor_name(i).decode() for i in range(l.cdm_plan_n_tensors())]


class Plan:
    """cdm_plan handle: packed weights + activation workspace + every layer's prepared launch for one
    (batch, reps) shape.  The arena / workspace are torch byte tensors owned here (the library allocates nothing)."""

    def __init__(self, tensors, n_cfeat, batch, reps, device, conv_mode=0):
        l = lib()
        self.names = plan_tensor_names()
        assert len(tensors) == len(self.names)
        for nme, t in zip(self.names, tensors):
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                raise CdmError(f"plan tensor {nme}: need a contiguous fp32 device tensor")
        self.tensors = list(tensors)  # keep-alive: the plan reads the small fp32 vectors in place
        self._ptrs = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        self.batch, self.reps, self.n = batch, reps, batch * reps
        self.arena = torch.empty(int(l.cdm_plan_arena_bytes(n_cfeat)), device=device, dtype=torch.uint8)
        self.ws = torch.empty(int(l.cdm_plan_workspace_bytes(batch, reps)), device=device, dtype=torch.uint8)
        d = PlanDesc(n_cfeat, batch, reps, self._ptrs, self.arena.data_ptr(), self.arena.numel(), self.ws.data_ptr(),
                     self.ws.numel(), conv_mode)
        h = C.c_void_p()
        check(l.cdm_plan_create(C.byref(d), stream_ptr(), C.byref(h)), "cdm_plan_create")
        self.handle = h

    def refresh(self):
        check(lib().cdm_plan_refresh(self.handle, stream_ptr()), "cdm_plan_refresh")

    def buffer(self, name, dtype, shape):
        """Typed view of a named workspace buffer."""
        p, nb = C.c_void_p(), C.c_longlong()
        check(lib().cdm_plan_buffer(self.handle, name.encode(), C.byref(p), C.byref(nb)), "cdm_plan_buffer")
        off = p.value - self.ws.data_ptr()
        return self.ws[off:off + nb.value].view(dtype).view(shape)

    def embed(self, which, inp, out):
        check(lib().cdm_plan_embed(self.handle, which, C.c_void_p(ptr(inp)), inp.shape[0], C.c_void_p(ptr(out)),
                                   stream_ptr()), "cdm_plan_embed")
        return out

    def _fwd_args(self, x, sc_tab, cemb1, temb1, cemb2, temb2, temb_rows, step_ptr, eps):
        return ForwardArgs(ptr(x), ptr(sc_tab), ptr(cemb1), ptr(temb1), ptr(cemb2), ptr(temb2), temb_rows,
                           ptr(step_ptr), ptr(eps))

    def forward_eval(self, x, sc_tab, cemb1, temb1, cemb2, temb2, temb_rows, step_ptr=None, eps=None):
        f = self._fwd_args(x, sc_tab, cemb1, temb1, cemb2, temb2, temb_rows, step_ptr, eps)
        check(lib().cdm_forward_eval(self.handle, C.byref(f), stream_ptr()), "cdm_forward_eval")

    def profile(self, x, sc_tab, cemb1, temb1, cemb2, temb2, temb_rows, step_ptr=None):
        """[(launch name, ms)] of one forward, CUDA events on the launch stream (cdm_plan_profile)."""
        l = lib()
        n = l.cdm_plan_n_launches()
        ms = (C.c_float * n)()
        f = self._fwd_args(x, sc_tab, cemb1, temb1, cemb2, temb2, temb_rows, step_ptr, None)
        check(l.cdm_plan_profile(self.handle, C.byref(f), ms, stream_ptr()), "cdm_plan_profile")
        return [(l.cdm_plan_launch_name(i).decode(), float(ms[i])) for i in range(n)]

    def sample_step(self, x, sc_tab, cemb1, temb1, cemb2, temb2, step_ptr, coef, timesteps, *, guide_w=0.0, z=None,
                    z_iter_stride=0, seed=0, sample_offset=0, snap=None, snap_slot=None):
        f = self._fwd_args(None, sc_tab, cemb1, temb1, cemb2, temb2, 1, None, None)
        a = SampleStepArgs(f, ptr(x), ptr(step_ptr), guide_w, ptr(coef), timesteps, ptr(z), z_iter_stride, seed,
                           sample_offset, ptr(snap), ptr(snap_slot))
        check(lib().cdm_sample_step(self.handle, C.byref(a), stream_ptr()), "cdm_sample_step")

    def close(self):
        if getattr(self, "handle", None) is not None and _lib is not None:
            _lib.cdm_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


# ----------------------------------------------------------------------------- training path
VP, I, F, LL = C.c_void_p, C.c_int, C.c_float, C.c_longlong


class GemmTnArgs(C.Structure):
    _fields_ = [("a", VP), ("a_c", I), ("b", VP), ("b_c", I), ("n_img", I), ("H", I), ("W", I), ("taps", I),
                ("m_off", I), ("M", I), ("n_off", I), ("N", I), ("c", VP), ("ldc", I), ("tap_stride", I),
                ("k_split", I), ("workspace", VP), ("workspace_floats", LL), ("probe", VP)]


class ChanReduceArgs(C.Structure):
    _fields_ = [("a", VP), ("lda", I), ("z", VP), ("ldz", I), ("scale", VP), ("shift", VP), ("mean", VP),
                ("rstd", VP), ("relu", I), ("mode", I), ("P", LL), ("C", I), ("workspace", VP),
                ("workspace_blocks", I), ("out", VP), ("xr", VP)]


class BnApplyArgs(C.Structure):
    _fields_ = [("z", VP), ("P", LL), ("C", I), ("relu", I), ("scale", VP), ("shift", VP), ("y", VP), ("sc_x", VP),
                ("sc_w", VP), ("sc_b", VP), ("film_scale", VP), ("film_shift", VP), ("film_rows", I),
                ("px_per_img", I), ("yf", VP), ("sums", VP), ("gamma", VP), ("beta", VP), ("count", F), ("eps", F),
                ("momentum", F), ("running_mean", VP), ("running_var", VP), ("scale_out", VP), ("shift_out", VP),
                ("mean_out", VP), ("rstd_out", VP)]


class BnBwdArgs(C.Structure):
    _fields_ = [("dy", VP), ("lddy", I), ("z", VP), ("P", LL), ("C", I), ("relu", I), ("scale", VP), ("shift", VP),
                ("mean", VP), ("rstd", VP), ("sums", VP), ("count", F), ("dz", VP)]


class GnBwdArgs(C.Structure):
    _fields_ = [("x", VP), ("dyf", VP), ("lddyf", I), ("n_img", I), ("P", I), ("C", I), ("groups", I),
                ("mean_rstd", VP), ("gamma", VP), ("beta", VP), ("film_scale", VP), ("dx", VP), ("dgamma_nc", VP),
                ("dbeta_nc", VP), ("dfs", VP), ("dfb", VP)]


class OuterWgradArgs(C.Structure):
    _fields_ = [("s", VP), ("v", VP), ("n_img", I), ("H", I), ("W", I), ("C", I), ("flip", I), ("mean_rstd", VP),
                ("gamma", VP), ("beta", VP), ("workspace", VP), ("workspace_blocks", I), ("out", VP)]


class EmbedBwdArgs(C.Structure):
    _fields_ = [("inp", VP), ("rows", I), ("din", I), ("emb", I), ("w1", VP), ("b1", VP), ("w2", VP), ("dout", VP),
                ("pre", VP), ("h", VP), ("dpre", VP), ("dw1", VP), ("db1", VP), ("dw2", VP), ("db2", VP)]


def rawptr(t):
    """data_ptr of a (possibly channel-sliced) device tensor view."""
    return None if t is None else t.data_ptr()


def gemm_tn(a, b, c, *, n_img, H, W, a_c, b_c, M, N, ldc, m_off=0, n_off=0, taps=1, tap_stride=0, k_split=0,
            workspace=None, probe=None):
    g = GemmTnArgs(rawptr(a), a_c, rawptr(b), b_c, n_img, H, W, taps, m_off, M, n_off, N, rawptr(c), ldc, tap_stride,
                   k_split, rawptr(workspace), 0 if workspace is None else workspace.numel(), rawptr(probe))
    check(lib().cdm_gemm_tn(C.byref(g), stream_ptr()), "cdm_gemm_tn")


def chan_reduce(a, lda, P, Cn, out, ws, *, mode=0, z=None, ldz=0, scale=None, shift=None, mean=None, rstd=None,
                relu=1, xr=None):
    g = ChanReduceArgs(rawptr(a), lda, rawptr(z), ldz, rawptr(scale), rawptr(shift), rawptr(mean), rawptr(rstd),
                       relu, mode, P, Cn, rawptr(ws), ws.numel() // (2 * Cn), rawptr(out), _xr_ptr(xr))
    check(lib().cdm_chan_reduce(C.byref(g), stream_ptr()), "cdm_chan_reduce")


def bn_finalize(sums, Cn, count, gamma, beta, eps, momentum, rm, rv, scale, shift, mean, rstd):
    check(lib().cdm_bn_finalize(VP(rawptr(sums)), Cn, F(count), VP(rawptr(gamma)), VP(rawptr(beta)), F(eps),
                                F(momentum), VP(rawptr(rm)), VP(rawptr(rv)), VP(rawptr(scale)), VP(rawptr(shift)),
                                VP(rawptr(mean)), VP(rawptr(rstd)), stream_ptr()), "cdm_bn_finalize")


def bn_apply(z, P, Cn, scale, shift, y, *, relu=1, sc_x=None, sc_w=None, sc_b=None, film_scale=None, film_shift=None,
             film_rows=1, px_per_img=1, yf=None, finalize=None):
    """finalize = (sums, gamma, beta, count, eps, momentum, running_mean, running_var, mean_out, rstd_out): derive
    scale / shift from the batch sums inside the launch (they are written to `scale` / `shift`)."""
    g = BnApplyArgs(rawptr(z), P, Cn, relu, rawptr(scale), rawptr(shift), rawptr(y), rawptr(sc_x), rawptr(sc_w),
                    rawptr(sc_b), rawptr(film_scale), rawptr(film_shift), film_rows, px_per_img, rawptr(yf))
    if finalize is not None:
        sums, gamma, beta, count, eps, mom, rm, rv, mean_out, rstd_out = finalize
        g.sums, g.gamma, g.beta, g.count, g.eps, g.momentum = rawptr(sums), rawptr(gamma), rawptr(beta), count, eps, mom
        g.running_mean, g.running_var = rawptr(rm), rawptr(rv)
        g.scale_out, g.shift_out, g.mean_out, g.rstd_out = rawptr(scale), rawptr(shift), rawptr(mean_out), rawptr(rstd_out)
    check(lib().cdm_bn_apply(C.byref(g), stream_ptr()), "cdm_bn_apply")


def bn_bwd_apply(dy, lddy, z, P, Cn, scale, shift, mean, rstd, sums, count, dz, relu=1):
    g = BnBwdArgs(rawptr(dy), lddy, rawptr(z), P, Cn, relu, rawptr(scale), rawptr(shift), rawptr(mean), rawptr(rstd),
                  rawptr(sums), count, rawptr(dz))
    check(lib().cdm_bn_bwd_apply(C.byref(g), stream_ptr()), "cdm_bn_bwd_apply")


def maxpool2_fwd(y, out):
    n, H, W, Cn = y.shape
    check(lib().cdm_maxpool2_fwd(VP(rawptr(y)), n, H, W, Cn, VP(rawptr(out)), stream_ptr()), "cdm_maxpool2_fwd")


def maxpool2_bwd(dpool, lddp, y, dy):
    n, H, W, Cn = y.shape
    check(lib().cdm_maxpool2_bwd(VP(rawptr(dpool)), lddp, VP(rawptr(y)), n, H, W, Cn, VP(rawptr(dy)), stream_ptr()),
          "cdm_maxpool2_bwd")


def add_bf16(a, lda, b, ldb, P, Cn):
    check(lib().cdm_add_bf16(VP(rawptr(a)), lda, VP(rawptr(b)), ldb, LL(P), Cn, stream_ptr()), "cdm_add_bf16")


def space_to_depth(dv, out):
    n, H2, W2, Cn = dv.shape
    check(lib().cdm_space_to_depth(VP(rawptr(dv)), n, H2 // 2, W2 // 2, Cn, VP(rawptr(out)), stream_ptr()),
          "cdm_space_to_depth")


def film_bwd(dyf, lddyf, y, n_img, px, Cn, fs, dy, dfs, dfb):
    check(lib().cdm_film_bwd(VP(rawptr(dyf)), lddyf, VP(rawptr(y)), n_img, px, Cn, VP(rawptr(fs)), VP(rawptr(dy)),
                             VP(rawptr(dfs)), VP(rawptr(dfb)), stream_ptr()), "cdm_film_bwd")


def gn_bwd(x, dyf, lddyf, n_img, P, Cn, groups, mean_rstd, gamma, beta, dx, dgamma_nc, dbeta_nc, *, film_scale=None,
           dfs=None, dfb=None):
    g = GnBwdArgs(rawptr(x), rawptr(dyf), lddyf, n_img, P, Cn, groups, rawptr(mean_rstd), rawptr(gamma), rawptr(beta),
                  rawptr(film_scale), rawptr(dx), rawptr(dgamma_nc), rawptr(dbeta_nc), rawptr(dfs), rawptr(dfb))
    check(lib().cdm_gn_bwd(C.byref(g), stream_ptr()), "cdm_gn_bwd")


def rows_sum(inp, rows, Cn, out):
    check(lib().cdm_rows_sum(VP(rawptr(inp)), rows, Cn, VP(rawptr(out)), stream_ptr()), "cdm_rows_sum")


def avgpool_gelu_train(src, pre, out):
    n, P, Cn = src.shape
    check(lib().cdm_avgpool_gelu_train(VP(rawptr(src)), n, P, Cn, VP(rawptr(pre)), VP(rawptr(out)), stream_ptr()),
          "cdm_avgpool_gelu_train")


def avgpool_gelu_bwd(pre, dh, n, P, Cn, dx):
    check(lib().cdm_avgpool_gelu_bwd(VP(rawptr(pre)), VP(rawptr(dh)), n, P, Cn, VP(rawptr(dx)), stream_ptr()),
          "cdm_avgpool_gelu_bwd")


def outer_wgrad(s, v, n_img, H, W, Cn, out, ws, *, flip=0, mean_rstd=None, gamma=None, beta=None):
    g = OuterWgradArgs(rawptr(s), rawptr(v), n_img, H, W, Cn, flip, rawptr(mean_rstd), rawptr(gamma), rawptr(beta),
                       rawptr(ws), ws.numel() // (9 * Cn), rawptr(out))
    check(lib().cdm_outer_wgrad(C.byref(g), stream_ptr()), "cdm_outer_wgrad")


def embed_bwd(inp, w1, b1, w2, dout, pre, h, dpre, dw1, db1, dw2, db2):
    rows, din = inp.shape
    g = EmbedBwdArgs(rawptr(inp), rows, din, w2.shape[0], rawptr(w1), rawptr(b1), rawptr(w2), rawptr(dout),
                     rawptr(pre), rawptr(h), rawptr(dpre), rawptr(dw1), rawptr(db1), rawptr(dw2), rawptr(db2))
    check(lib().cdm_embed_bwd(C.byref(g), stream_ptr()), "cdm_embed_bwd")


def mse_grad(pred, target, inv_count, dpred, partial, loss_sum):
    check(lib().cdm_mse_grad(VP(rawptr(pred)), VP(rawptr(target)), LL(pred.numel()), F(inv_count), VP(rawptr(dpred)),
                             VP(rawptr(partial)), partial.numel(), VP(rawptr(loss_sum)), stream_ptr()), "cdm_mse_grad")


def adam_step(table, n_tensors, max_numel, lr, beta1, beta2, eps, step, lr_dev=None, step_dev=None):
    check(lib().cdm_adam_step(VP(rawptr(table)), n_tensors, LL(max_numel), F(lr), F(beta1), F(beta2), F(eps), step,
                              VP(rawptr(lr_dev)), VP(rawptr(step_dev)), stream_ptr()), "cdm_adam_step")


# --------------------------------------------------------------------------- map statistics
def power_spectrum(maps, bin_start, bin_items, scale, pk):
    """maps fp32 [n,N,N]; bin_start int32 [n_bins+1]; bin_items int32 [N*N]; pk fp64 [n,n_bins]."""
    n, N, _ = maps.shape
    check(lib().cdm_power_spectrum(C.c_void_p(ptr(maps)), n, N, C.c_void_p(ptr(bin_start)), C.c_void_p(ptr(bin_items)),
                                   pk.shape[1], C.c_double(scale), C.c_void_p(ptr(pk)), stream_ptr()),
          "cdm_power_spectrum")
    return pk


def pixel_histogram(maps, edges, counts):
    """maps fp32 [n,P]; edges fp64 [n_bins+1]; counts int32 [n,n_bins]."""
    n, P = maps.shape
    check(lib().cdm_pixel_histogram(C.c_void_p(ptr(maps)), n, P, C.c_void_p(ptr(edges)), counts.shape[1],
                                    C.c_void_p(ptr(counts)), stream_ptr()), "cdm_pixel_histogram")
    return counts


# --------------------------------------------------------------------------- data preparation
def minmax(x, workspace, out):
    """x fp32 (any shape, contiguous); workspace fp32 zero-filled once; out fp32 [2] = (min, max)."""
    check(lib().cdm_minmax(C.c_void_p(ptr(x)), C.c_longlong(x.numel()), C.c_void_p(ptr(workspace)), workspace.numel(),
                           C.c_void_p(ptr(out)), stream_ptr()), "cdm_minmax")
    return out


def preprocess_maps(maps, raw_minmax, out):
    """maps fp32 [n,Hi,Wi]; raw_minmax fp32 [2] (device); out fp32 [n,Ho,Wo]."""
    n, Hi, Wi = maps.shape
    check(lib().cdm_preprocess_maps(C.c_void_p(ptr(maps)), n, Hi, Wi, C.c_void_p(ptr(raw_minmax)), out.shape[1],
                                    out.shape[2], C.c_void_p(ptr(out)), stream_ptr()), "cdm_preprocess_maps")
    return out


def normalize_params(x, repeat, out, col_min, col_max):
    """x fp32 [rows,cols]; out fp32 [rows*repeat,out_cols]; col_min/col_max fp32 [cols]."""
    rows, cols = x.shape
    check(lib().cdm_normalize_params(C.c_void_p(ptr(x)), rows, cols, repeat, out.shape[1], C.c_void_p(ptr(out)),
                                     C.c_void_p(ptr(col_min)), C.c_void_p(ptr(col_max)), stream_ptr()),
          "cdm_normalize_params")
    return out


def xrank_sum(partial, out, xr=None):
    """partial fp32 [n_blocks, n]; out fp32 [n] = sum over blocks (fixed order) and over the ranks of `xr`."""
    n_blocks, n = partial.shape
    check(lib().cdm_xrank_sum(C.c_void_p(ptr(partial)), n_blocks, n, C.c_void_p(ptr(out)), _xr_ptr(xr), stream_ptr()),
          "cdm_xrank_sum")
    return out


def xrank_set_timeout(seconds):
    check(lib().cdm_xrank_set_timeout(C.c_double(seconds)), "cdm_xrank_set_timeout")


def pack_transpose_bf16(src, dst, batches, R, Cc, sb, sr, db, dc):
    """dst[b][c][r] (bf16) = src[b][r][c] (fp32), strides in elements (cdm_pack_transpose_bf16)."""
    LLc = C.c_longlong
    check(lib().cdm_pack_transpose_bf16(C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), batches, R, Cc, LLc(sb),
                                        LLc(sr), LLc(db), LLc(dc), stream_ptr()), "cdm_pack_transpose_bf16")


def pack_bf16(table, n_rows, total_vec):
    """table int64 [n_rows, 12] on the device (see cdm_pack_bf16)."""
    check(lib().cdm_pack_bf16(C.c_void_p(ptr(table)), n_rows, C.c_longlong(total_vec), stream_ptr()), "cdm_pack_bf16")
