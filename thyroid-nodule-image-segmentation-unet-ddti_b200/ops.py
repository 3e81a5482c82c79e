"""Tensor-level wrappers over the C ABI (include/b2s.h). PyTorch supplies device memory and streams only; every
computation below is a libb2s kernel. Activations are NHWC bf16 `Act` views (a channel slice of a buffer)."""
import ctypes
import functools

import torch

from . import _lib
from ._lib import B2S_FLAG_RELU, B2S_FLAG_STATS, check

BF16 = torch.bfloat16

# When set to a list, every wrapper below appends (name, kind, work, start_event, end_event, nbytes): kind "tensor"
# -> work = algorithmic FLOPs, kind "hbm" -> work = algorithmic bytes; nbytes = algorithmic HBM bytes of
# a "tensor" launch whose arithmetic intensity is below the ridge (the high-resolution transposed convs), else None.
# Used by bench.py's roofline pass.
PROFILE = None


def _timed(name, kind, work, fn, nbytes=None):
    if PROFILE is None:
        return fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    r = fn()
    e.record()
    PROFILE.append((name, kind, float(work), s, e, nbytes))
    return r


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def on_device_of_input(fn):
    """Decorator for entry points (module forward, autograd Function forward / backward): runs the body with the CUDA
    device of the first tensor argument current. libb2s launches on the CURRENT device's current stream and never
    calls cudaSetDevice, so a model on cuda:1 called while cuda:0 is current would otherwise launch on the wrong GPU."""
    @functools.wraps(fn)
    def wrapper(self, x, *args, **kwargs):
        if isinstance(x, torch.Tensor) and x.is_cuda and x.device.index != torch.cuda.current_device():
            with torch.cuda.device(x.device):
                return fn(self, x, *args, **kwargs)
        return fn(self, x, *args, **kwargs)
    return wrapper


def _p(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.B2SError("libb2s kernels need CUDA tensors; there is no CPU fallback")


class Act:
    """Channel slice [c0, c0+C) of an NHWC bf16 buffer [N,H,W,Ctot]."""

    __slots__ = ("buf", "c0", "C")

    def __init__(self, buf, c0=0, C=None):
        assert buf.dtype == BF16 and buf.dim() == 4 and buf.is_contiguous()
        self.buf = buf
        self.c0 = c0
        self.C = buf.shape[3] - c0 if C is None else C
        assert 0 <= c0 and c0 + self.C <= buf.shape[3] and c0 % 8 == 0

    @staticmethod
    def empty(N, H, W, C, device):
        return Act(torch.empty((N, H, W, C), dtype=BF16, device=device))

    @property
    def N(self):
        return self.buf.shape[0]

    @property
    def H(self):
        return self.buf.shape[1]

    @property
    def W(self):
        return self.buf.shape[2]

    @property
    def cstride(self):
        return self.buf.shape[3]

    @property
    def ptr(self):
        return ctypes.c_void_p(self.buf.data_ptr() + 2 * self.c0)

    def slice(self, c0, C):
        return Act(self.buf, self.c0 + c0, C)

    def view(self):
        return self.buf[..., self.c0:self.c0 + self.C]

    def to_nchw_float(self):
        return self.view().permute(0, 3, 1, 2).float().contiguous()

    @staticmethod
    def from_nchw(x):
        return Act(x.permute(0, 2, 3, 1).contiguous().to(BF16))


# ---------------------------------------------------------------------------------------------------------
# weights
# ---------------------------------------------------------------------------------------------------------
def pack_conv_weight(w, want_dgrad=True):
    """w [Cout,Cin,k,k] fp32 -> (w_fwd [k*k,Cout,Cin] bf16, w_dgrad [k*k,Cin,Cout] bf16 | None)."""
    _need_cuda(w)
    Cout, Cin, k, _ = w.shape
    w = w.detach().contiguous().float()
    wf = torch.empty((k * k, Cout, Cin), dtype=BF16, device=w.device)
    wd = torch.empty((k * k, Cin, Cout), dtype=BF16, device=w.device) if want_dgrad else None
    _timed("pack_conv_weight", "hbm", w.numel() * (4.0 + 2.0 * (2 if want_dgrad else 1)), lambda: check(
        _lib.lib().b2s_pack_conv_weight(_p(w), _p(wf), _p(wd), Cout, Cin, k, _stream()), "pack_conv_weight"))
    return wf, wd


def pack_convt_weight(w):
    """w [Cin,Cout,2,2] fp32 -> (w_fwd [4*Cout,Cin], w_dgrad [4*Cin,Cout]) bf16."""
    _need_cuda(w)
    Cin, Cout = w.shape[0], w.shape[1]
    w = w.detach().contiguous().float()
    wf = torch.empty((4 * Cout, Cin), dtype=BF16, device=w.device)
    wd = torch.empty((4 * Cin, Cout), dtype=BF16, device=w.device)
    _timed("pack_convt_weight", "hbm", w.numel() * 8.0, lambda: check(
        _lib.lib().b2s_pack_convt_weight(_p(w), _p(wf), _p(wd), Cin, Cout, _stream()), "pack_convt_weight"))
    return wf, wd


class PackPlan:
    """Pre-allocated bf16 operands of a set of conv / transposed-conv weights and the argument arrays of the single
    b2s_pack_weights_all launch that refreshes all of them."""

    def __init__(self, weights, want_dgrad, outputs=None):
        """weights: list of (name, fp32 tensor, is_convt); 4-D [Cout,Cin,k,k] or transposed-conv [Cin,Cout,2,2].
        outputs: optional name -> (w_fwd, w_dgrad) of ANOTHER plan to write into (a sub-plan refreshing part of it)."""
        self.items = []
        n = len(weights)
        self._w = (ctypes.c_void_p * n)()
        self._wf = (ctypes.c_void_p * n)()
        self._wd = (ctypes.c_void_p * n)()
        self._d0, self._d1 = (ctypes.c_int * n)(), (ctypes.c_int * n)()
        self._taps, self._kind = (ctypes.c_int * n)(), (ctypes.c_int * n)()
        self.packed, self.nbytes, self.n = {}, 0.0, n
        for i, (name, w, is_convt) in enumerate(weights):
            _need_cuda(w)
            assert w.dtype == torch.float32 and w.is_contiguous()
            given = outputs[name] if outputs is not None else None
            if is_convt:
                Cin, Cout = w.shape[0], w.shape[1]
                wf = given[0] if given else torch.empty((4 * Cout, Cin), dtype=BF16, device=w.device)
                wd = given[1] if given else torch.empty((4 * Cin, Cout), dtype=BF16, device=w.device)
                d0, d1, taps, kind = Cin, Cout, 4, 1
            else:
                Cout, Cin, k, _ = w.shape
                wf = given[0] if given else torch.empty((k * k, Cout, Cin), dtype=BF16, device=w.device)
                wd = (given[1] if given else torch.empty((k * k, Cin, Cout), dtype=BF16, device=w.device)) \
                    if want_dgrad else None
                d0, d1, taps, kind = Cout, Cin, k * k, 0
            self.packed[name] = (wf, wd)
            self._w[i], self._wf[i] = w.data_ptr(), wf.data_ptr()
            self._wd[i] = wd.data_ptr() if wd is not None else None
            self._d0[i], self._d1[i], self._taps[i], self._kind[i] = d0, d1, taps, kind
            self.nbytes += w.numel() * (4.0 + 2.0 * (2 if wd is not None else 1))
            self.items.append(w)          # keeps the sources alive (their addresses are baked into the arrays)

    def run(self):
        _timed("pack_weights_all", "hbm", self.nbytes, lambda: check(
            _lib.lib().b2s_pack_weights_all(self.n, self._w, self._wf, self._wd, self._d0, self._d1, self._taps,
                                            self._kind, _stream()), "b2s_pack_weights_all"))


# ---------------------------------------------------------------------------------------------------------
# tensor-core convs
# ---------------------------------------------------------------------------------------------------------
def conv_fwd(x, w_packed, bias, y, ksize=3, relu=False, stats=None, tile_n=0):
    """y = conv(x, w) (+bias)(+ReLU); stats: fp32 [conv_stats_rows(...), 2, Cout] partial buffer or None."""
    flags = (B2S_FLAG_RELU if relu else 0) | (B2S_FLAG_STATS if stats is not None else 0)
    Cout = y.C
    flops = 2.0 * x.N * x.H * x.W * x.C * Cout * ksize * ksize
    _timed(f"conv{ksize}x{ksize}[{x.C}->{Cout}@{x.H}x{x.W}]", "tensor", flops, lambda: check(
        _lib.lib().b2s_conv_fwd(x.ptr, x.cstride, _p(w_packed), _p(bias), y.ptr, y.cstride, _p(stats), x.N, x.H,
                                x.W, x.C, Cout, ksize, flags, tile_n, _stream()), "b2s_conv_fwd"))


def conv_fwd_affine(x, w_packed, bias, post_scale, post_shift, y, ksize=3, relu=True, tile_n=0):
    """inference: y = act(conv(x) + bias) * post_scale + post_shift (eval-mode BatchNorm in the epilogue)"""
    flags = B2S_FLAG_RELU if relu else 0
    Cout = y.C
    flops = 2.0 * x.N * x.H * x.W * x.C * Cout * ksize * ksize
    _timed(f"conv{ksize}x{ksize}+bn[{x.C}->{Cout}@{x.H}x{x.W}]", "tensor", flops, lambda: check(
        _lib.lib().b2s_conv_fwd_affine(x.ptr, x.cstride, _p(w_packed), _p(bias), _p(post_scale), _p(post_shift), y.ptr,
                                       y.cstride, x.N, x.H, x.W, x.C, Cout, ksize, flags, tile_n, _stream()),
        "b2s_conv_fwd_affine"))


def maxpool2x2(x, pooled):
    _timed("maxpool2x2", "hbm", x.N * x.H * x.W * x.C * 2.5, lambda: check(
        _lib.lib().b2s_maxpool2x2(x.ptr, x.cstride, pooled.ptr, x.N, x.H, x.W, x.C, _stream()), "b2s_maxpool2x2"))


def conv_stats_rows(N, H, W, Cout, tile_n=0):
    """rows of the [rows, 2, Cout] partial-statistics buffer conv_fwd(..., stats=...) fills for this shape"""
    r = _lib.lib().b2s_conv_stats_rows(N, H, W, Cout, tile_n)
    if r <= 0:
        raise _lib.B2SError("b2s_conv_stats_rows: unsupported shape")
    return r


def convt_fwd(x, w_packed, bias, y, tile_n=0):
    flops = 8.0 * x.N * x.H * x.W * x.C * y.C
    nbytes = 2.0 * x.N * x.H * x.W * (x.C + 4 * y.C) + 8.0 * x.C * y.C      # read x, write the 2x up-sampled y, weights
    _timed(f"convT_fwd[{x.C}->{y.C}@{x.H}x{x.W}]", "tensor", flops, lambda: check(
        _lib.lib().b2s_convt2x2_fwd(x.ptr, x.cstride, _p(w_packed), _p(bias), y.ptr, y.cstride, x.N, x.H, x.W, x.C,
                                    y.C, tile_n, _stream()), "b2s_convt2x2_fwd"), nbytes)


def convt_dgrad(dy, w_packed_d, dx, tile_n=0):
    flops = 8.0 * dx.N * dx.H * dx.W * dx.C * dy.C
    nbytes = 2.0 * dx.N * dx.H * dx.W * (dx.C + 4 * dy.C) + 8.0 * dx.C * dy.C
    _timed(f"convT_dgrad[{dx.C}<-{dy.C}@{dx.H}x{dx.W}]", "tensor", flops, lambda: check(
        _lib.lib().b2s_convt2x2_dgrad(dy.ptr, dy.cstride, _p(w_packed_d), dx.ptr, dx.cstride, dx.N, dx.H, dx.W,
                                      dx.C, dy.C, tile_n, _stream()), "b2s_convt2x2_dgrad"), nbytes)


def conv_dgrad_bnred(dz, w_packed_d, dy, r, partial, ksize=3, tile_n=0):
    """dy = input gradient of a conv (conv of dz with the rotated weights) AND partial [rows][2][C] = {sum dy, sum dy*r}
    for the BatchNorm backward that consumes dy (r: its saved input). Returns the partial's row count, or 0 when the
    shape takes the one-tile kernel (nothing launched: use conv_fwd + the two-pass bn_bwd)."""
    flops = 2.0 * dz.N * dz.H * dz.W * dz.C * dy.C * ksize * ksize
    rc = _timed(f"conv{ksize}x{ksize}[{dz.C}->{dy.C}@{dz.H}x{dz.W}]", "tensor", flops, lambda: _lib.lib().b2s_conv_dgrad_bnred(
        dz.ptr, dz.cstride, _p(w_packed_d), dy.ptr, dy.cstride, r.ptr, r.cstride, _p(partial), dz.N, dz.H, dz.W, dz.C,
        dy.C, ksize, tile_n, _stream()))
    if rc == 1:
        if PROFILE is not None:
            PROFILE.pop()
        return 0
    check(rc, "b2s_conv_dgrad_bnred")
    return conv_stats_rows(dz.N, dz.H, dz.W, dy.C, tile_n)


def convt_dgrad_bnred(dy, w_packed_d, dx, r, partial, tile_n=0):
    """transposed-conv input gradient with the same fused reduction over its output dx; returns rows or 0"""
    flops = 8.0 * dx.N * dx.H * dx.W * dx.C * dy.C
    nbytes = 2.0 * dx.N * dx.H * dx.W * (dx.C + 4 * dy.C) + 8.0 * dx.C * dy.C
    rc = _timed(f"convT_dgrad[{dx.C}<-{dy.C}@{dx.H}x{dx.W}]", "tensor", flops, lambda: _lib.lib().b2s_convt2x2_dgrad_bnred(
        dy.ptr, dy.cstride, _p(w_packed_d), dx.ptr, dx.cstride, r.ptr, r.cstride, _p(partial), dx.N, dx.H, dx.W, dx.C,
        dy.C, tile_n, _stream()), nbytes)
    if rc == 1:
        if PROFILE is not None:
            PROFILE.pop()
        return 0
    check(rc, "b2s_convt2x2_dgrad_bnred")
    rows = _lib.lib().b2s_convt2x2_dgrad_rows(dx.N, dx.H, dx.W, dx.C, tile_n)
    if rows <= 0:
        raise _lib.B2SError("b2s_convt2x2_dgrad_rows: unsupported shape")
    return rows


def wgrad_workspace(N, H, W, Cin, Cout, taps, tile_n=0, splits=0):
    s = ctypes.c_int(0)
    nbytes = _lib.lib().b2s_conv_wgrad_workspace(N, H, W, Cin, Cout, 3 if taps == 9 else taps, tile_n, splits,
                                                 ctypes.byref(s))
    if nbytes < 0:
        check(-1, "b2s_conv_wgrad_workspace")
    return nbytes, s.value


def conv3x3_wgrad(x, dz, ws, dw, tile_n=0, splits=0):
    """dw [Cout,Cin,3,3] fp32 = sum_pix dz (x) x ; ws: fp32 workspace tensor."""
    nbytes, s = wgrad_workspace(x.N, x.H, x.W, x.C, dz.C, 9, tile_n, splits)
    assert ws.numel() * 4 >= nbytes, "wgrad workspace too small"
    L = _lib.lib()
    flops = 18.0 * x.N * x.H * x.W * x.C * dz.C
    _timed(f"wgrad3x3[{x.C}->{dz.C}@{x.H}x{x.W}]", "tensor", flops, lambda: check(
        L.b2s_conv3x3_wgrad(x.ptr, x.cstride, dz.ptr, dz.cstride, _p(ws), x.N, x.H, x.W, x.C, dz.C, tile_n, splits,
                            _stream()), "b2s_conv3x3_wgrad"))
    _timed(f"wgrad_reduce[{x.C}->{dz.C},s={s}]", "hbm", 4.0 * 9 * x.C * dz.C * (s + 1), lambda: check(
        L.b2s_wgrad_reduce(_p(ws), s, 9, x.C, dz.C, _p(dw), 0, _stream()), "b2s_wgrad_reduce"))


def convt_wgrad(x, dy, ws, dw, tile_n=0, splits=0):
    """dw [Cin,Cout,2,2] fp32; x [N,Hi,Wi,Cin], dy [N,2Hi,2Wi,Cout]."""
    nbytes, s = wgrad_workspace(x.N, x.H, x.W, x.C, dy.C, 4, tile_n, splits)
    assert ws.numel() * 4 >= nbytes, "wgrad workspace too small"
    L = _lib.lib()
    flops = 8.0 * x.N * x.H * x.W * x.C * dy.C
    nbytes = 2.0 * x.N * x.H * x.W * (x.C + 4 * dy.C) + 16.0 * x.C * dy.C * s
    _timed(f"convT_wgrad[{x.C}->{dy.C}@{x.H}x{x.W}]", "tensor", flops, lambda: check(
        L.b2s_convt2x2_wgrad(x.ptr, x.cstride, dy.ptr, dy.cstride, _p(ws), x.N, x.H, x.W, x.C, dy.C, tile_n, splits,
                             _stream()), "b2s_convt2x2_wgrad"), nbytes)
    _timed(f"wgradT_reduce[{x.C}->{dy.C},s={s}]", "hbm", 4.0 * 4 * x.C * dy.C * (s + 1), lambda: check(
        L.b2s_wgrad_reduce(_p(ws), s, 4, x.C, dy.C, _p(dw), 1, _stream()), "b2s_wgrad_reduce"))


# ---------------------------------------------------------------------------------------------------------
# bandwidth kernels
# ---------------------------------------------------------------------------------------------------------
def c1_rows(N, H, W):
    return _lib.lib().b2s_c1_rows(N, H, W)


def conv3x3_c1_fwd(x, w, bias, r, relu=True, stats=None):
    """x [N,1,H,W] or [N,H,W] fp32 -> r Act [N,H,W,Cout]."""
    assert x.dtype == torch.float32 and x.is_contiguous()
    flags = (B2S_FLAG_RELU if relu else 0) | (B2S_FLAG_STATS if stats is not None else 0)
    assert r.c0 == 0 and r.C == r.cstride
    npix = r.N * r.H * r.W
    _timed("conv3x3_c1_fwd", "hbm", npix * (4.0 + 2.0 * r.C), lambda: check(
        _lib.lib().b2s_conv3x3_c1_fwd(_p(x), _p(w), _p(bias), r.ptr, _p(stats), r.N, r.H, r.W, r.C, flags,
                                      _stream()), "b2s_conv3x3_c1_fwd"))


def conv3x3_c1_fwd_affine(x, w, bias, post_scale, post_shift, y, relu=True):
    """inference: y = act(conv(x) + bias) * post_scale + post_shift"""
    assert x.dtype == torch.float32 and x.is_contiguous() and y.c0 == 0 and y.C == y.cstride
    npix = y.N * y.H * y.W
    _timed("conv3x3_c1_fwd+bn", "hbm", npix * (4.0 + 2.0 * y.C), lambda: check(
        _lib.lib().b2s_conv3x3_c1_fwd_affine(_p(x), _p(w), _p(bias), _p(post_scale), _p(post_shift), y.ptr, y.N, y.H, y.W,
                                             y.C, B2S_FLAG_RELU if relu else 0, _stream()), "b2s_conv3x3_c1_fwd_affine"))


def conv3x3_c1_wgrad(x, dz, partial, scratch, dw):
    assert dz.c0 == 0 and dz.C == dz.cstride
    L = _lib.lib()
    rows = L.b2s_c1_rows(dz.N, dz.H, dz.W)
    npix = dz.N * dz.H * dz.W
    _timed("conv3x3_c1_wgrad", "hbm", npix * (4.0 + 2.0 * dz.C), lambda: check(
        L.b2s_conv3x3_c1_wgrad(_p(x), dz.ptr, _p(partial), dz.N, dz.H, dz.W, dz.C, _stream()), "c1_wgrad"))
    check(L.b2s_reduce_rows(_p(partial), rows, dz.C * 9, _p(scratch), _p(dw), _stream()), "reduce_rows")


def reduce_rows(partial, rows, K, scratch, out):
    check(_lib.lib().b2s_reduce_rows(_p(partial), rows, K, _p(scratch), _p(out), _stream()), "b2s_reduce_rows")


def ew_rows():
    return _lib.lib().b2s_ew_rows()


def bn_finalize(partial, rows, C, count, gamma, beta, running_mean, running_var, nbt, momentum, eps, scale, shift,
                mean, invstd, scratch):
    check(_lib.lib().b2s_bn_finalize(_p(partial), rows, C, float(count), _p(gamma), _p(beta), _p(running_mean),
                                     _p(running_var), _p(nbt), momentum, eps, _p(scale), _p(shift), _p(mean),
                                     _p(invstd), _p(scratch), _stream()), "b2s_bn_finalize")


def bn_eval_affine(gamma, beta, rm, rv, eps, scale, shift):
    check(_lib.lib().b2s_bn_eval_affine(_p(gamma), _p(beta), _p(rm), _p(rv), eps, _p(scale), _p(shift),
                                        scale.numel(), _stream()), "b2s_bn_eval_affine")


def bn_apply(r, scale, shift, y, pooled=None):
    nbytes = r.N * r.H * r.W * r.C * 2.0 * (2.25 if pooled is not None else 2.0)
    _timed("bn_apply_pool" if pooled is not None else "bn_apply", "hbm", nbytes, lambda: check(
        _lib.lib().b2s_bn_apply(r.ptr, r.cstride, _p(scale), _p(shift), y.ptr, y.cstride,
                                pooled.ptr if pooled is not None else None, r.N, r.H, r.W, r.C, _stream()),
        "b2s_bn_apply"))


def bn_bwd(dy, dpool, r, scale, shift, mean, invstd, gamma, count, dz, partial, scratch, coef, dgamma, dbeta, dbias,
           pre=None):
    """Full BatchNorm(+ReLU, + optional max-pool routing) backward: writes dz, dgamma, dbeta, dbias.
    pre = (partial, rows): {sum dy, sum dy*r} partials already emitted by the launch that produced dy
    (conv_dgrad_bnred / convt_dgrad_bnred); the reduce pass over dy and r is then skipped."""
    L = _lib.lib()
    rows = L.b2s_ew_rows()
    C = r.C
    dp = dpool.ptr if dpool is not None else None
    nel = r.N * r.H * r.W * C * 2.0
    extra = 0.25 if dpool is not None else 0.0
    if pre is not None:
        assert dpool is None
        check(L.b2s_bn_bwd_finalize_raw(_p(pre[0]), pre[1], C, float(count), _p(gamma), _p(mean), _p(invstd), _p(dgamma),
                                        _p(dbeta), _p(coef), _p(scratch), _stream()), "b2s_bn_bwd_finalize_raw")
    else:
        _timed("bn_bwd_reduce", "hbm", nel * (2.0 + extra), lambda: check(
            L.b2s_bn_bwd_reduce(dy.ptr, dy.cstride, dp, r.ptr, r.cstride, _p(scale), _p(shift), _p(mean), _p(invstd),
                                _p(partial), r.N, r.H, r.W, C, _stream()), "b2s_bn_bwd_reduce"))
        check(L.b2s_bn_bwd_finalize(_p(partial), rows, C, float(count), _p(gamma), _p(invstd), _p(dgamma), _p(dbeta),
                                    _p(coef), _p(scratch), _stream()), "b2s_bn_bwd_finalize")
    _timed("bn_bwd_apply", "hbm", nel * (3.0 + extra), lambda: check(
        L.b2s_bn_bwd_apply(dy.ptr, dy.cstride, dp, r.ptr, r.cstride, _p(scale), _p(shift), _p(mean), _p(invstd),
                           _p(coef), dz.ptr, dz.cstride, _p(partial), r.N, r.H, r.W, C, _stream()),
        "b2s_bn_bwd_apply"))
    check(L.b2s_reduce_rows(_p(partial), rows, C, _p(scratch), _p(dbias), _stream()), "b2s_reduce_rows")


def head_fwd(r, scale, shift, w, b, logits, mask=None):
    """logits [N,O,H,W] fp32 = conv1x1(BN(r)); mask uint8 optional."""
    O = logits.shape[1]
    npix = r.N * r.H * r.W
    _timed("head_fwd", "hbm", npix * (2.0 * r.C + 4.0 * O + (O if mask is not None else 0)), lambda: check(
        _lib.lib().b2s_head_fwd(r.ptr, r.cstride, _p(scale), _p(shift), _p(w), _p(b), _p(logits), _p(mask), r.N,
                                r.H * r.W, r.C, O, _stream()), "b2s_head_fwd"))


def head_bwd(dlogits, r, scale, shift, w, dy, partial, scratch, dw_db):
    """dy (Act) and dw_db fp32 [O*C+O] (weight grad then bias grad)."""
    L = _lib.lib()
    O = dlogits.shape[1]
    npix = r.N * r.H * r.W
    _timed("head_bwd", "hbm", npix * (4.0 * r.C + 4.0 * O), lambda: check(
        L.b2s_head_bwd(_p(dlogits), r.ptr, r.cstride, _p(scale), _p(shift), _p(w), dy.ptr, dy.cstride, _p(partial),
                       r.N, r.H * r.W, r.C, O, _stream()), "b2s_head_bwd"))
    check(L.b2s_reduce_rows(_p(partial), L.b2s_ew_rows(), O * r.C + O, _p(scratch), _p(dw_db), _stream()),
          "b2s_reduce_rows")


def loss_chunks(per_sample):
    return _lib.lib().b2s_loss_chunks(per_sample)


def seg_loss_fwd(logits, targets, partial, sums, out, dice_smooth=1.0, w_bce=1.0, w_dice=1.0, w_ft=0.0, ft_alpha=0.4,
                 ft_beta=0.6, ft_gamma=2.0, ft_smooth=1e-6):
    B = logits.shape[0]
    per = logits.numel() // B
    _timed("seg_loss_fwd", "hbm", 8.0 * B * per, lambda: check(
        _lib.lib().b2s_seg_loss_fwd(_p(logits), _p(targets), B, per, _p(partial), _p(sums), _p(out), dice_smooth,
                                    w_bce, w_dice, w_ft, ft_alpha, ft_beta, ft_gamma, ft_smooth, _stream()),
        "b2s_seg_loss_fwd"))


def seg_loss_bwd(logits, targets, sums, ft_tot, grad_out, dlogits, dice_smooth=1.0, w_bce=1.0, w_dice=1.0, w_ft=0.0,
                 ft_alpha=0.4, ft_beta=0.6, ft_gamma=2.0, ft_smooth=1e-6):
    B = logits.shape[0]
    per = logits.numel() // B
    _timed("seg_loss_bwd", "hbm", 12.0 * B * per, lambda: check(
        _lib.lib().b2s_seg_loss_bwd(_p(logits), _p(targets), _p(sums), _p(ft_tot), B, per, B * per, B,
                                    _p(grad_out), _p(dlogits), dice_smooth, w_bce, w_dice, w_ft, ft_alpha, ft_beta,
                                    ft_gamma, ft_smooth, _stream()), "b2s_seg_loss_bwd"))


def seg_metrics(logits, targets, counters):
    """adds the confusion counts of sigmoid(logits) > 0.5 against targets to counters (int64 [7], device)"""
    lg, tg = logits.detach().float().contiguous(), targets.detach().float().contiguous()
    n = lg.numel()
    partial = torch.empty(_lib.lib().b2s_metrics_blocks(n) * 6, dtype=torch.int32, device=lg.device)
    _timed("seg_metrics", "hbm", 8.0 * n, lambda: check(
        _lib.lib().b2s_seg_metrics(_p(lg), _p(tg), n, _p(partial), _p(counters), _stream()), "b2s_seg_metrics"))


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    _timed("adamw", "hbm", 28.0 * p.numel(), lambda: check(
        _lib.lib().b2s_adamw_step(_p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, weight_decay, step,
                                  grad_scale, _stream()), "b2s_adamw_step"))


def adamw_step_dev(p, g, m, v, hyper):
    """hyper: device fp32 [8] = lr, beta1, beta2, eps, wd, 1-beta1^t, sqrt(1-beta2^t), grad_scale"""
    _timed("adamw", "hbm", 28.0 * p.numel(), lambda: check(
        _lib.lib().b2s_adamw_step_dev(_p(p), _p(g), _p(m), _p(v), p.numel(), _p(hyper), _stream()), "b2s_adamw_step_dev"))


def copy_channels(src, dst):
    npix = src.N * src.H * src.W
    _timed("copy_channels", "hbm", npix * src.C * 4.0, lambda: check(
        _lib.lib().b2s_copy_channels(src.ptr, src.cstride, dst.ptr, dst.cstride, npix, src.C, _stream()),
        "b2s_copy_channels"))


# ---------------------------------------------------------------------------------------------------------
# V-Net variant (models/vnet.py): stride-2 conv, 1x1 wgrad, BN->ReLU->Dropout(+residual), SE block, channel sums
# ---------------------------------------------------------------------------------------------------------
def conv3x3_s2_fwd(x, w_packed, bias, y, tile_n=0):
    """y [N,H/2,W/2,Cout] = conv3x3(x, stride 2, pad 1) + bias"""
    flops = 18.0 * y.N * y.H * y.W * x.C * y.C
    _timed(f"conv3x3s2[{x.C}->{y.C}@{x.H}x{x.W}]", "tensor", flops, lambda: check(
        _lib.lib().b2s_conv3x3_s2_fwd(x.ptr, x.cstride, _p(w_packed), _p(bias), y.ptr, y.cstride, x.N, x.H, x.W, x.C,
                                      y.C, tile_n, _stream()), "b2s_conv3x3_s2_fwd"))


def conv3x3_s2_dgrad(dz, w_packed_d, dx, tile_n=0):
    """dx [N,H,W,Cin] from dz [N,H/2,W/2,Cout]; returns False when the image is too small (use zero insertion)"""
    flops = 18.0 * dz.N * dz.H * dz.W * dx.C * dz.C
    rc = _timed(f"dgrad3x3s2[{dx.C}<-{dz.C}@{dx.H}x{dx.W}]", "tensor", flops, lambda: _lib.lib().b2s_conv3x3_s2_dgrad(
        dz.ptr, dz.cstride, _p(w_packed_d), dx.ptr, dx.cstride, dx.N, dx.H, dx.W, dx.C, dz.C, tile_n, _stream()))
    if rc == 1:
        return False
    check(rc, "b2s_conv3x3_s2_dgrad")
    return True


def conv3x3_s2_wgrad(x, dz, dw, tile_n=0, splits=0):
    """dw [Cout,Cin,3,3] fp32 of the stride-2 conv"""
    LEGACY = 1 << 12
    nbytes, s = wgrad_workspace(dz.N, dz.H, dz.W, x.C, dz.C, 9, tile_n | LEGACY, splits)
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=dw.device)
    L = _lib.lib()
    flops = 18.0 * dz.N * dz.H * dz.W * x.C * dz.C
    _timed(f"wgrad3x3s2[{x.C}->{dz.C}@{x.H}x{x.W}]", "tensor", flops, lambda: check(
        L.b2s_conv3x3_s2_wgrad(x.ptr, x.cstride, dz.ptr, dz.cstride, _p(ws), x.N, x.H, x.W, x.C, dz.C, tile_n, splits,
                               _stream()), "b2s_conv3x3_s2_wgrad"))
    check(L.b2s_wgrad_reduce(_p(ws), s, 9, x.C, dz.C, _p(dw), 0, _stream()), "b2s_wgrad_reduce")


def upsample_zero2x(src, dst):
    _timed("upsample_zero2x", "hbm", 2.0 * src.C * src.N * src.H * src.W * 5.0, lambda: check(
        _lib.lib().b2s_upsample_zero2x(src.ptr, src.cstride, dst.ptr, dst.cstride, src.N, src.H, src.W, src.C,
                                       _stream()), "b2s_upsample_zero2x"))


def conv1x1_wgrad(x, dz, dw, tile_n=0, splits=0):
    """dw [Cout,Cin,1,1] fp32"""
    nbytes, s = wgrad_workspace(x.N, x.H, x.W, x.C, dz.C, 1, tile_n, splits)
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=dw.device)
    L = _lib.lib()
    _timed(f"wgrad1x1[{x.C}->{dz.C}@{x.H}x{x.W}]", "tensor", 2.0 * x.N * x.H * x.W * x.C * dz.C, lambda: check(
        L.b2s_conv1x1_wgrad(x.ptr, x.cstride, dz.ptr, dz.cstride, _p(ws), x.N, x.H, x.W, x.C, dz.C, tile_n, splits,
                            _stream()), "b2s_conv1x1_wgrad"))
    check(L.b2s_wgrad_reduce(_p(ws), s, 1, x.C, dz.C, _p(dw), 0, _stream()), "b2s_wgrad_reduce")


def bn_act_apply(z, scale, shift, res, out, relu=True, dropout_p=0.0, seed=0, step_counter=None):
    nb = z.N * z.H * z.W * z.C * 2.0 * (3.0 if res is not None else 2.0)
    _timed("bn_act_apply", "hbm", nb, lambda: check(
        _lib.lib().b2s_bn_act_apply(z.ptr, z.cstride, _p(scale), _p(shift), res.ptr if res is not None else None,
                                    res.cstride if res is not None else 0, out.ptr, out.cstride,
                                    z.N * z.H * z.W, z.C, int(relu), float(dropout_p), int(seed) & 0xFFFFFFFF,
                                    _p(step_counter), _stream()), "b2s_bn_act_apply"))


def bn_act_bwd(da, z, scale, shift, mean, invstd, gamma, count, dz, dgamma, dbeta, dbias, relu=True, dropout_p=0.0,
               seed=0, step_counter=None):
    """BatchNorm(train) -> ReLU -> Dropout backward: writes dz (Act), dgamma, dbeta, dbias (fp32 [C])."""
    L = _lib.lib()
    rows, C = L.b2s_ew_rows(), z.C
    dev = z.buf.device
    partial = torch.empty(rows * 2 * C, dtype=torch.float32, device=dev)
    scratch = torch.empty(128 * 2 * C, dtype=torch.float32, device=dev)
    coef = torch.empty(3 * C, dtype=torch.float32, device=dev)
    npix = z.N * z.H * z.W
    seed = int(seed) & 0xFFFFFFFF
    _timed("bn_act_bwd_reduce", "hbm", npix * C * 4.0, lambda: check(
        L.b2s_bn_act_bwd_reduce(da.ptr, da.cstride, z.ptr, z.cstride, _p(scale), _p(shift), _p(mean), _p(invstd),
                                _p(partial), npix, C, int(relu), float(dropout_p), seed, _p(step_counter), _stream()),
        "b2s_bn_act_bwd_reduce"))
    check(L.b2s_bn_bwd_finalize(_p(partial), rows, C, float(count), _p(gamma), _p(invstd), _p(dgamma), _p(dbeta),
                                _p(coef), _p(scratch), _stream()), "b2s_bn_bwd_finalize")
    _timed("bn_act_bwd_apply", "hbm", npix * C * 6.0, lambda: check(
        L.b2s_bn_act_bwd_apply(da.ptr, da.cstride, z.ptr, z.cstride, _p(scale), _p(shift), _p(mean), _p(invstd),
                               _p(coef), dz.ptr, dz.cstride, _p(partial), npix, C, int(relu), float(dropout_p), seed,
                               _p(step_counter), _stream()), "b2s_bn_act_bwd_apply"))
    check(L.b2s_reduce_rows(_p(partial), rows, C, _p(scratch), _p(dbias), _stream()), "b2s_reduce_rows")


def relu_bwd(dy, y, dx):
    _timed("relu_bwd", "hbm", y.N * y.H * y.W * y.C * 6.0, lambda: check(
        _lib.lib().b2s_relu_bwd(dy.ptr, dy.cstride, y.ptr, y.cstride, dx.ptr, dx.cstride, y.N * y.H * y.W, y.C, _stream()),
        "b2s_relu_bwd"))


def maxpool2x2_bwd(x, dpool, dx):
    _timed("maxpool2x2_bwd", "hbm", x.N * x.H * x.W * x.C * 4.5, lambda: check(
        _lib.lib().b2s_maxpool2x2_bwd(x.ptr, x.cstride, dpool.ptr, dpool.cstride, dx.ptr, dx.cstride, x.N, x.H, x.W, x.C,
                                      _stream()), "b2s_maxpool2x2_bwd"))


def channel_sums(x, out):
    """out [C] fp32 = sum over pixels of x (Act)"""
    L = _lib.lib()
    rows, C = L.b2s_ew_rows(), x.C
    partial = torch.empty(rows * C, dtype=torch.float32, device=out.device)
    scratch = torch.empty(128 * C, dtype=torch.float32, device=out.device)
    _timed("channel_sums", "hbm", x.N * x.H * x.W * C * 2.0, lambda: check(
        L.b2s_channel_sums(x.ptr, x.cstride, _p(partial), x.N * x.H * x.W, C, _stream()), "b2s_channel_sums"))
    check(L.b2s_reduce_rows(_p(partial), rows, C, _p(scratch), _p(out), _stream()), "b2s_reduce_rows")


def se_forward(x, w1, b1, w2, b2, y):
    """SEBlock forward: returns (mean, hidden, gate) fp32 [N,C], [N,C/r], [N,C]; y = x * gate."""
    L = _lib.lib()
    N, HW, C, Cr = x.N, x.H * x.W, x.C, w1.shape[0]
    dev = x.buf.device
    chunks = L.b2s_se_chunks(HW)
    partial = torch.empty(N * chunks * C, dtype=torch.float32, device=dev)
    mean = torch.empty((N, C), dtype=torch.float32, device=dev)
    hidden = torch.empty((N, Cr), dtype=torch.float32, device=dev)
    gate = torch.empty((N, C), dtype=torch.float32, device=dev)
    _timed("se_pool", "hbm", N * HW * C * 2.0, lambda: check(
        L.b2s_se_pool(x.ptr, x.cstride, None, 0, _p(partial), N, HW, C, _stream()), "b2s_se_pool"))
    check(L.b2s_se_fc_fwd(_p(partial), chunks, HW, _p(w1), _p(b1), _p(w2), _p(b2), _p(mean), _p(hidden), _p(gate), N, C,
                          Cr, _stream()), "b2s_se_fc_fwd")
    _timed("se_scale", "hbm", N * HW * C * 4.0, lambda: check(
        L.b2s_se_scale(x.ptr, x.cstride, _p(gate), None, 0.0, y.ptr, y.cstride, N, HW, C, _stream()), "b2s_se_scale"))
    return mean, hidden, gate


def se_backward(dy, x, mean, hidden, gate, w1, w2, dx):
    """returns (dw1 [Cr,C], db1, dw2 [C,Cr], db2); dx = dy*gate + dmean/HW"""
    L = _lib.lib()
    N, HW, C, Cr = x.N, x.H * x.W, x.C, w1.shape[0]
    dev = x.buf.device
    f32 = dict(dtype=torch.float32, device=dev)
    chunks = L.b2s_se_chunks(HW)
    partial = torch.empty(N * chunks * C, **f32)
    ds, dh, dmean = torch.empty((N, C), **f32), torch.empty((N, Cr), **f32), torch.empty((N, C), **f32)
    dw1, db1 = torch.empty((Cr, C), **f32), torch.empty(Cr, **f32)
    dw2, db2 = torch.empty((C, Cr), **f32), torch.empty(C, **f32)
    _timed("se_pool_dot", "hbm", N * HW * C * 4.0, lambda: check(
        L.b2s_se_pool(dy.ptr, dy.cstride, x.ptr, x.cstride, _p(partial), N, HW, C, _stream()), "b2s_se_pool"))
    check(L.b2s_se_fc_bwd(_p(partial), chunks, _p(gate), _p(hidden), _p(mean), _p(w1), _p(w2), _p(ds), _p(dh),
                          _p(dmean), _p(dw1), _p(db1), _p(dw2), _p(db2), N, C, Cr, _stream()), "b2s_se_fc_bwd")
    _timed("se_scale_bwd", "hbm", N * HW * C * 4.0, lambda: check(
        L.b2s_se_scale(dy.ptr, dy.cstride, _p(gate), _p(dmean), 1.0 / HW, dx.ptr, dx.cstride, N, HW, C, _stream()),
        "b2s_se_scale"))
    return dw1, db1, dw2, db2


# ---------------------------------------------------------------------------------------------------------
# models/mod.py AttentionGate, bilinear re-size branch, multi-channel input images
# ---------------------------------------------------------------------------------------------------------
def _map_ptrs(maps):
    arr = (ctypes.c_void_p * 4)()
    for i, m in enumerate(maps):
        assert m.dtype == torch.float32 and m.is_contiguous() and m.numel() == maps[0].numel()
        arr[i] = m.data_ptr()
    return arr


def psi_forward(maps, gamma, beta, rm, rv, nbt, training):
    """psi = sigmoid(BatchNorm2d(1)(sum(maps))) (models/mod.py:223-227,233). maps: 1..4 fp32 tensors [N,1,H,W].
    Returns (psi, mean, invstd); mean / invstd are None in eval mode."""
    L = _lib.lib()
    _need_cuda(*maps)
    n = maps[0].numel()
    dev = maps[0].device
    f32 = dict(dtype=torch.float32, device=dev)
    arr = _map_ptrs(maps)
    scale, shift = torch.empty(1, **f32), torch.empty(1, **f32)
    mean = invstd = None
    if training:
        rows = L.b2s_psi_rows(n)
        partial = torch.empty(rows * 2, **f32)
        scratch = torch.empty(128 * 2, **f32)
        mean, invstd = torch.empty(1, **f32), torch.empty(1, **f32)
        _timed("psi_stats", "hbm", 4.0 * n * len(maps), lambda: check(
            L.b2s_psi_stats(arr, len(maps), n, _p(partial), _stream()), "b2s_psi_stats"))
        bn_finalize(partial, rows, 1, float(n), gamma, beta, rm, rv, nbt, 0.1, 1e-5, scale, shift, mean, invstd, scratch)
    else:
        bn_eval_affine(gamma, beta, rm, rv, 1e-5, scale, shift)
    psi = torch.empty_like(maps[0])
    _timed("psi_fwd", "hbm", 4.0 * n * (len(maps) + 1), lambda: check(
        L.b2s_psi_fwd(arr, len(maps), _p(scale), _p(shift), _p(psi), n, _stream()), "b2s_psi_fwd"))
    return psi, mean, invstd


def psi_backward(maps, psi, dpsi, mean, invstd, gamma):
    """returns (dv [same shape as a map], dgamma [1], dbeta [1])"""
    L = _lib.lib()
    n = psi.numel()
    f32 = dict(dtype=torch.float32, device=psi.device)
    arr = _map_ptrs(maps)
    rows = L.b2s_psi_rows(n)
    partial, scratch = torch.empty(rows * 2, **f32), torch.empty(128 * 2, **f32)
    coef, dgamma, dbeta = torch.empty(3, **f32), torch.empty(1, **f32), torch.empty(1, **f32)
    dpsi = dpsi.contiguous().float()
    _timed("psi_bwd_reduce", "hbm", 4.0 * n * (len(maps) + 2), lambda: check(
        L.b2s_psi_bwd_reduce(arr, len(maps), _p(psi), _p(dpsi), _p(mean), _p(invstd), n, _p(partial), _stream()),
        "b2s_psi_bwd_reduce"))
    check(L.b2s_bn_bwd_finalize(_p(partial), rows, 1, float(n), _p(gamma), _p(invstd), _p(dgamma), _p(dbeta), _p(coef),
                                _p(scratch), _stream()), "b2s_bn_bwd_finalize")
    dv = torch.empty_like(psi)
    _timed("psi_bwd_apply", "hbm", 4.0 * n * (len(maps) + 3), lambda: check(
        L.b2s_psi_bwd_apply(arr, len(maps), _p(psi), _p(dpsi), _p(mean), _p(invstd), _p(coef), _p(dv), n, _stream()),
        "b2s_psi_bwd_apply"))
    return dv, dgamma, dbeta


def pixel_scale_fwd(x, psi, out):
    npix = x.N * x.H * x.W
    assert psi.dtype == torch.float32 and psi.is_contiguous() and psi.numel() == npix
    _timed("pixel_scale", "hbm", npix * (4.0 * x.C + 4.0), lambda: check(
        _lib.lib().b2s_pixel_scale_fwd(x.ptr, x.cstride, _p(psi), out.ptr, out.cstride, npix, x.C, _stream()),
        "b2s_pixel_scale_fwd"))


def pixel_scale_bwd(x, psi, dy, dx, dpsi):
    npix = x.N * x.H * x.W
    _timed("pixel_scale_bwd", "hbm", npix * (6.0 * x.C + 8.0), lambda: check(
        _lib.lib().b2s_pixel_scale_bwd(x.ptr, x.cstride, _p(psi), dy.ptr, dy.cstride, dx.ptr, dx.cstride, _p(dpsi), npix,
                                       x.C, _stream()), "b2s_pixel_scale_bwd"))


def bilinear_fwd(x, y):
    _timed("bilinear", "hbm", 2.0 * x.C * x.N * (x.H * x.W + y.H * y.W), lambda: check(
        _lib.lib().b2s_bilinear_fwd(x.ptr, x.cstride, y.ptr, y.cstride, x.N, x.H, x.W, y.H, y.W, x.C, _stream()),
        "b2s_bilinear_fwd"))


def bilinear_bwd(dy, dx):
    _timed("bilinear_bwd", "hbm", 2.0 * dx.C * dx.N * (dx.H * dx.W + dy.H * dy.W), lambda: check(
        _lib.lib().b2s_bilinear_bwd(dy.ptr, dy.cstride, dx.ptr, dx.cstride, dx.N, dx.H, dx.W, dy.H, dy.W, dx.C, _stream()),
        "b2s_bilinear_bwd"))


def image_to_nhwc(x, cpad=64, out=None):
    """x [N,C,H,W] fp32 -> Act [N,H,W,cpad] bf16 (channels >= C zero)"""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 4
    x = x.contiguous()
    N, C, H, W = x.shape
    y = out if out is not None else Act.empty(N, H, W, cpad, x.device)
    assert y.c0 == 0 and y.C == y.cstride == cpad
    _timed("image_to_nhwc", "hbm", N * H * W * (4.0 * C + 2.0 * cpad), lambda: check(
        _lib.lib().b2s_image_to_nhwc(_p(x), y.ptr, N, C, H * W, cpad, _stream()), "b2s_image_to_nhwc"))
    return y
