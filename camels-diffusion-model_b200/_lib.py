"""ctypes binding of libcdm_b200.so (declared in include/cdm_b200.h).

There is no fallback: if the shared library is missing or the device is not an
sm_100 GPU every op raises.  PyTorch is only used for device memory and streams.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcdm_b200.so")

EPI_RELU, EPI_SHORTCUT, EPI_POOL, EPI_FILM, EPI_GNSTATS = 1, 2, 4, 8, 16
CONV_MODE_COPIES, CONV_MODE_SHIFT24, CONV_MODE_SHIFT18 = 0, 1, 2


class CdmError(RuntimeError):
    pass


class Conv3x3Args(C.Structure):
    _fields_ = [
        ("src0", C.c_void_p), ("src1", C.c_void_p), ("c0", C.c_int), ("c1", C.c_int),
        ("n_img", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("weight", C.c_void_p), ("cout", C.c_int),
        ("scale", C.c_void_p), ("shift", C.c_void_p), ("flags", C.c_int), ("out", C.c_void_p),
        ("sc_x", C.c_void_p), ("sc_nx", C.c_int), ("sc_tab", C.c_void_p),
        ("film_scale", C.c_void_p), ("film_shift", C.c_void_p), ("film_shift_rows", C.c_int),
        ("step_ptr", C.c_void_p), ("gn_partial", C.c_void_p), ("mode", C.c_int),
    ]


class GemmArgs(C.Structure):
    _fields_ = [
        ("a0", C.c_void_p), ("a1", C.c_void_p), ("k0", C.c_int), ("k1", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("bw", C.c_void_p), ("shift", C.c_void_p),
        ("shift_mod", C.c_int), ("out_mode", C.c_int), ("H", C.c_int), ("W", C.c_int), ("out", C.c_void_p),
    ]


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
    for name in EXPORTS:
        getattr(l, name)  # raises AttributeError if the header and the library disagree
    _lib = l
    return l


# every symbol include/cdm_b200.h declares
EXPORTS = [
    "cdm_version", "cdm_last_error", "cdm_device_ok",
    "cdm_conv3x3", "cdm_gemm", "cdm_probe_tma_l2",
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


def conv3x3(src0, weight, scale, shift, out, *, src1=None, flags=EPI_RELU, sc_x=None, sc_tab=None,
            film_scale=None, film_shift=None, film_shift_rows=1, step_ptr=None, gn_partial=None,
            mode=CONV_MODE_COPIES):
    """src*: bf16 [n,H,W,c]; weight bf16 [cout,3,3,cin]; out bf16 NHWC. See cdm_conv3x3 in cdm_b200.h."""
    n, H, W, c0 = src0.shape
    a = Conv3x3Args()
    a.src0, a.c0 = ptr(src0), c0
    a.src1, a.c1 = (ptr(src1), src1.shape[3]) if src1 is not None else (None, 0)
    a.n_img, a.H, a.W = n, H, W
    a.weight, a.cout = ptr(weight), weight.shape[0]
    assert weight.shape[3] == a.c0 + a.c1
    a.scale, a.shift, a.flags, a.out = ptr(scale), ptr(shift), flags, ptr(out)
    a.sc_x, a.sc_nx, a.sc_tab = ptr(sc_x), (sc_x.shape[0] if sc_x is not None else 0), ptr(sc_tab)
    a.film_scale, a.film_shift, a.film_shift_rows = ptr(film_scale), ptr(film_shift), film_shift_rows
    a.step_ptr, a.gn_partial, a.mode = ptr(step_ptr), ptr(gn_partial), mode
    check(lib().cdm_conv3x3(C.byref(a), stream_ptr()), "cdm_conv3x3")
    return out


def gemm(a0, bw, shift, out, *, a1=None, shift_mod=None, out_mode=0, H=0, W=0):
    """a*: bf16 [M,k]; bw bf16 [N,K]; see cdm_gemm in cdm_b200.h."""
    g = GemmArgs()
    g.a0, g.k0 = ptr(a0), a0.shape[1]
    g.a1, g.k1 = (ptr(a1), a1.shape[1]) if a1 is not None else (None, 0)
    g.M, g.N = a0.shape[0], bw.shape[0]
    assert bw.shape[1] == g.k0 + g.k1
    g.bw, g.shift = ptr(bw), ptr(shift)
    g.shift_mod = shift_mod if shift_mod is not None else shift.numel()
    g.out_mode, g.H, g.W, g.out = out_mode, H, W, ptr(out)
    check(lib().cdm_gemm(C.byref(g), stream_ptr()), "cdm_gemm")
    return out


def probe_tma_l2(buf, n_rows, iters):
    check(lib().cdm_probe_tma_l2(C.c_void_p(ptr(buf)), n_rows, iters, stream_ptr()), "cdm_probe_tma_l2")
