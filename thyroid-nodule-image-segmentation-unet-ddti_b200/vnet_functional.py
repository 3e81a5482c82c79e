"""Autograd nodes of the V-Net variant (reference models/vnet.py) over libb2s kernels.

Activations flow between the nodes as NHWC bf16 tensors ``[N,H,W,C]`` (contiguous, or a channel slice of a
contiguous concat buffer). Every node's forward AND backward is a sequence of libb2s kernel launches (ops.py);
torch.autograd only walks the graph (fan-out gradient sums at the residual / skip branches are its one arithmetic
contribution). The UNet path (engine.py) does not use autograd internally; this composition trades some fusion
for covering the V-Net's larger op set (stride-2 conv, 1x1 projections, SE blocks, Conv->BN->ReLU->Dropout order).
"""
import torch

from . import ops
from .ops import Act, BF16

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def as_act(t):
    """NHWC bf16 tensor (contiguous or a channel slice of a contiguous buffer) -> ops.Act"""
    if t.dtype != BF16 or t.dim() != 4:
        raise ValueError(f"expected a 4-D bf16 NHWC tensor, got {t.dtype} {tuple(t.shape)}")
    if not t.is_cuda:
        raise ops._lib.B2SError("libb2s kernels need CUDA tensors; there is no CPU fallback")
    N, H, W, C = t.shape
    if t.is_contiguous() and t.storage_offset() % 8 == 0:
        return Act(t)
    s = t.stride()
    Ct = s[2]
    if s[3] != 1 or s[1] != W * Ct or s[0] != H * W * Ct or Ct < C:
        return Act(t.contiguous())
    c0 = t.storage_offset() % Ct
    base = torch.as_strided(t, (N, H, W, Ct), (H * W * Ct, W * Ct, Ct, 1), t.storage_offset() - c0)
    return Act(base, c0, C)


# Packed bf16 operands registered by the owning module (the nets pack every conv weight in a few launches per step):
# data_ptr -> (weakref to the parameter, version, w_fwd, w_dgrad). An entry is honoured only while its parameter is
# alive (a freed parameter's address can be reused by an unrelated tensor) and unchanged; otherwise the node packs
# the weight on the spot.
PACKED = {}


def register_packed(param, wf, wd):
    import weakref
    PACKED[param.data_ptr()] = (weakref.ref(param), param._version, wf, wd)


def _lookup(w, need_dgrad):
    e = PACKED.get(w.data_ptr())
    if e is None:
        return None
    owner = e[0]()
    if owner is None or owner.data_ptr() != w.data_ptr() or owner.shape != w.shape or e[1] != w._version:
        return None
    if need_dgrad and e[3] is None:
        return None
    return e[2], e[3]


def packed_conv(w, want_dgrad):
    hit = _lookup(w, want_dgrad)
    return hit if hit is not None else ops.pack_conv_weight(w, want_dgrad=want_dgrad)


def packed_convt(w):
    hit = _lookup(w, True)
    return hit if hit is not None else ops.pack_convt_weight(w)


def new_act(N, H, W, C, device):
    return torch.empty((N, H, W, C), dtype=BF16, device=device)


def _bn_affine(training, stats, rows, count, gamma, beta, rm, rv, nbt, C, device):
    f32 = dict(dtype=torch.float32, device=device)
    scale, shift, mean, invstd = (torch.empty(C, **f32) for _ in range(4))
    if training:
        scratch = torch.empty(128 * 2 * C, **f32)
        ops.bn_finalize(stats, rows, C, count, gamma, beta, rm, rv, nbt, BN_MOMENTUM, BN_EPS, scale, shift, mean, invstd,
                        scratch)
    else:
        ops.bn_eval_affine(gamma, beta, rm, rv, BN_EPS, scale, shift)
        mean, invstd = None, None
    return scale, shift, mean, invstd


_PAIR = 1 << 10          # conv entry points' tile_n flag: force the tile-pair kernel (its partial-statistics row count
                         # is the one b2s_conv_stats_rows reports for the same flag, also for 1x1 convs)


class ConvBnAct(torch.autograd.Function):
    """out = dropout(relu(BN(conv(x) + b))) (+ res)      models/vnet.py:51-59 (one loop trip, + the residual add)

    The conv is 3x3 (pad 1) or 1x1, read off w (AttentionGate's W_g / W_x are 1x1 conv + BN, models/mod.py:214-221).
    x: NHWC bf16 activation, or the fp32 image [N,1,H,W] for the first conv of a branch (Cin = 1 kernel)."""

    @staticmethod
    def forward(ctx, x, res, w, b, gamma, beta, rm, rv, nbt, training, p_drop, seed, relu_mode=1):
        """relu_mode 1: ReLU before the residual add (V-Net ConvBlock); 2: after it (ResidualBlock, models/mod.py:84);
        b may be None (bias=False convs of models/mod.py)."""
        dev = w.device
        bias = b.detach() if b is not None else None
        Cout = w.shape[0]
        first = x.dtype == torch.float32
        if first:
            N, _, H, W = x.shape
            x = x.contiguous()
        else:
            xa = as_act(x)
            N, H, W = xa.N, xa.H, xa.W
        z = new_act(N, H, W, Cout, dev)
        za = Act(z)
        if first:
            rows = ops.c1_rows(N, H, W)
            stats = torch.empty(rows * 2 * Cout, dtype=torch.float32, device=dev) if training else None
            ops.conv3x3_c1_fwd(x, w.detach(), bias, za, relu=False, stats=stats)
        else:
            k = w.shape[2]
            tile_n = 0 if k == 3 else _PAIR
            wf, _ = packed_conv(w, False)
            rows = ops.conv_stats_rows(N, H, W, Cout, tile_n)
            stats = torch.empty(rows * 2 * Cout, dtype=torch.float32, device=dev) if training else None
            ops.conv_fwd(xa, wf, bias, za, ksize=k, relu=False, stats=stats, tile_n=tile_n)
        scale, shift, mean, invstd = _bn_affine(training, stats, rows, float(N * H * W), gamma.detach(), beta.detach(),
                                                rm, rv, nbt, Cout, dev)
        out = new_act(N, H, W, Cout, dev)
        # the dropout mask is keyed on (seed, this layer's num_batches_tracked as read on the device): a replayed CUDA
        # graph therefore draws a fresh mask every step, and backward re-derives the mask of its own step
        use_drop = training and p_drop > 0
        ops.bn_act_apply(za, scale, shift, as_act(res) if res is not None else None, Act(out), relu=relu_mode,
                         dropout_p=p_drop if training else 0.0, seed=seed, step_counter=nbt if use_drop else None)
        ctx.nbt = nbt if use_drop else None
        ctx.save_for_backward(x, z, w, gamma, scale, shift, mean if mean is not None else scale,
                              invstd if invstd is not None else scale, out if relu_mode == 2 else scale)
        ctx.cfg = (first, training, p_drop, seed, res is not None, relu_mode, b is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, z, w, gamma, scale, shift, mean, invstd, out_saved = ctx.saved_tensors
        first, training, p_drop, seed, has_res, relu_mode, has_bias = ctx.cfg
        if not training:
            raise RuntimeError("b200seg V-Net: backward through eval-mode BatchNorm is not implemented")
        dev = w.device
        Cout, Cin = w.shape[0], w.shape[1]
        za = Act(z)
        N, H, W = za.N, za.H, za.W
        da = as_act(dout)
        if relu_mode == 2:     # ReLU after the add: mask the incoming gradient with the saved output once, for both branches
            dmasked = new_act(N, H, W, Cout, dev)
            ops.relu_bwd(da, Act(out_saved), Act(dmasked))
            dout, da = dmasked, Act(dmasked)
        f32 = dict(dtype=torch.float32, device=dev)
        dz = new_act(N, H, W, Cout, dev)
        dgamma, dbeta, dbias = torch.empty(Cout, **f32), torch.empty(Cout, **f32), torch.empty(Cout, **f32)
        ops.bn_act_bwd(da, za, scale, shift, mean, invstd, gamma.detach(), float(N * H * W), Act(dz), dgamma, dbeta,
                       dbias, relu=1 if relu_mode == 1 else 0, dropout_p=p_drop, seed=seed, step_counter=ctx.nbt)
        k = w.shape[2]
        dw = torch.empty((Cout, Cin, k, k), **f32)
        dx = None
        if first:
            rows = ops.c1_rows(N, H, W)
            partial = torch.empty(rows * Cout * 9, **f32)
            scratch = torch.empty(128 * Cout * 9, **f32)
            ops.conv3x3_c1_wgrad(x, Act(dz), partial, scratch, dw)
        else:
            xa = as_act(x)
            if k == 3:
                nbytes, _ = ops.wgrad_workspace(N, H, W, Cin, Cout, 9)
                ws = torch.empty(nbytes // 4, **f32)
                ops.conv3x3_wgrad(xa, Act(dz), ws, dw)
            else:
                ops.conv1x1_wgrad(xa, Act(dz), dw)
            if ctx.needs_input_grad[0]:
                _, wd = packed_conv(w, True)
                dx = new_act(N, H, W, Cin, dev)
                ops.conv_fwd(Act(dz), wd, None, Act(dx), ksize=k)
        dres = dout if has_res else None
        return dx, dres, dw, (dbias if has_bias else None), dgamma, dbeta, None, None, None, None, None, None, None


class Conv1x1(torch.autograd.Function):
    """nn.Conv2d(Cin, Cout, 1) on NHWC bf16 (residual projections, models/vnet.py:46,58). For the fp32 image
    (Cin = 1) the 1x1 conv is the centre tap of the Cin = 1 3x3 kernel."""

    @staticmethod
    def forward(ctx, x, w, b):
        dev = w.device
        Cout = w.shape[0]
        first = x.dtype == torch.float32
        if first:
            N, _, H, W = x.shape
            x = x.contiguous()
            w3 = torch.zeros((Cout, 1, 3, 3), dtype=torch.float32, device=dev)
            w3[:, :, 1, 1] = w.detach()[:, :, 0, 0]
            out = new_act(N, H, W, Cout, dev)
            ops.conv3x3_c1_fwd(x, w3, b.detach() if b is not None else None, Act(out), relu=False, stats=None)
        else:
            xa = as_act(x)
            N, H, W = xa.N, xa.H, xa.W
            wf, _ = packed_conv(w, False)
            out = new_act(N, H, W, Cout, dev)
            ops.conv_fwd(xa, wf, b.detach() if b is not None else None, Act(out), ksize=1)
        ctx.save_for_backward(x, w)
        ctx.first = first
        ctx.has_bias = b is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        dev = w.device
        Cout, Cin = w.shape[0], w.shape[1]
        dy = as_act(dout)
        N, H, W = dy.N, dy.H, dy.W
        f32 = dict(dtype=torch.float32, device=dev)
        db = None
        if ctx.has_bias:
            db = torch.empty(Cout, **f32)
            ops.channel_sums(dy, db)
        dx = None
        if ctx.first:
            dyc = Act(dout.contiguous()) if dy.c0 or dy.cstride != dy.C else dy
            rows = ops.c1_rows(N, H, W)
            partial = torch.empty(rows * Cout * 9, **f32)
            scratch = torch.empty(128 * Cout * 9, **f32)
            dw3 = torch.empty((Cout, 1, 3, 3), **f32)
            ops.conv3x3_c1_wgrad(x, dyc, partial, scratch, dw3)
            dw = dw3[:, :, 1:2, 1:2].contiguous()
        else:
            xa = as_act(x)
            dw = torch.empty((Cout, Cin, 1, 1), **f32)
            ops.conv1x1_wgrad(xa, dy, dw)
            if ctx.needs_input_grad[0]:
                _, wd = packed_conv(w, True)
                dx = new_act(N, H, W, Cin, dev)
                ops.conv_fwd(dy, wd, None, Act(dx), ksize=1)
        return dx, dw, db


class ConvS2(torch.autograd.Function):
    """nn.Conv2d(C, 2C, 3, stride=2, padding=1) (models/vnet.py:97). Forward and weight gradient read x through TMA
    boxes over one pixel-parity class per tap; the input gradient is four parity-class convolutions over dz."""

    @staticmethod
    def forward(ctx, x, w, b):
        xa = as_act(x)
        Cout = w.shape[0]
        wf, _ = packed_conv(w, False)
        out = new_act(xa.N, xa.H // 2, xa.W // 2, Cout, w.device)
        ops.conv3x3_s2_fwd(xa, wf, b.detach(), Act(out))
        ctx.save_for_backward(x, w)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        dev = w.device
        xa = as_act(x)
        dy = as_act(dout)
        Cout, Cin = w.shape[0], w.shape[1]
        f32 = dict(dtype=torch.float32, device=dev)
        db = torch.empty(Cout, **f32)
        ops.channel_sums(dy, db)
        dw = torch.empty((Cout, Cin, 3, 3), **f32)
        ops.conv3x3_s2_wgrad(xa, dy, dw)                 # x read as parity-class boxes: no zero insertion
        dx = None
        if ctx.needs_input_grad[0]:
            _, wd = packed_conv(w, True)
            dx = new_act(xa.N, xa.H, xa.W, Cin, dev)
            if not ops.conv3x3_s2_dgrad(dy, wd, Act(dx)):   # tiny image: stride-1 dgrad of the zero-inserted gradient
                up = new_act(xa.N, xa.H, xa.W, Cout, dev)
                ops.upsample_zero2x(dy, Act(up))
                ops.conv_fwd(Act(up), wd, None, Act(dx), ksize=3)
        return dx, dw, db


class ConvT2x2(torch.autograd.Function):
    """nn.ConvTranspose2d(Cin, Cout, 2, stride=2) (models/vnet.py:101-104)"""

    @staticmethod
    def forward(ctx, x, w, b):
        xa = as_act(x)
        Cout = w.shape[1]
        wf, _ = packed_convt(w)
        out = new_act(xa.N, 2 * xa.H, 2 * xa.W, Cout, w.device)
        ops.convt_fwd(xa, wf, b.detach(), Act(out))
        ctx.save_for_backward(x, w)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        dev = w.device
        xa = as_act(x)
        dy = as_act(dout)
        Cin, Cout = w.shape[0], w.shape[1]
        f32 = dict(dtype=torch.float32, device=dev)
        db = torch.empty(Cout, **f32)
        ops.channel_sums(dy, db)
        dw = torch.empty((Cin, Cout, 2, 2), **f32)
        nbytes, _ = ops.wgrad_workspace(xa.N, xa.H, xa.W, Cin, Cout, 4)
        ws = torch.empty(nbytes // 4, **f32)
        ops.convt_wgrad(xa, dy, ws, dw)
        dx = None
        if ctx.needs_input_grad[0]:
            _, wd = packed_convt(w)
            dx = new_act(xa.N, xa.H, xa.W, Cin, dev)
            ops.convt_dgrad(dy, wd, Act(dx))
        return dx, dw, db


class SE(torch.autograd.Function):
    """SEBlock (models/vnet.py:18-26): x * sigmoid(W2 relu(W1 avgpool(x) + b1) + b2)"""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        xa = as_act(x)
        C, Cr = w1.shape[1], w1.shape[0]
        out = new_act(xa.N, xa.H, xa.W, C, w1.device)
        w1m, w2m = w1.detach().reshape(Cr, C).contiguous(), w2.detach().reshape(C, Cr).contiguous()
        mean, hidden, gate = ops.se_forward(xa, w1m, b1.detach(), w2m, b2.detach(), Act(out))
        ctx.save_for_backward(x, mean, hidden, gate, w1m, w2m)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, mean, hidden, gate, w1m, w2m = ctx.saved_tensors
        xa = as_act(x)
        dy = as_act(dout)
        dx = new_act(xa.N, xa.H, xa.W, xa.C, x.device)
        dw1, db1, dw2, db2 = ops.se_backward(dy, xa, mean, hidden, gate, w1m, w2m, Act(dx))
        Cr, C = dw1.shape
        return dx, dw1.view(Cr, C, 1, 1), db1, dw2.view(C, Cr, 1, 1), db2


class MaxPool2x2(torch.autograd.Function):
    """nn.MaxPool2d(2, 2) (models/mod.py:27): backward routes the gradient to the first maximum of each window"""

    @staticmethod
    def forward(ctx, x):
        xa = as_act(x)
        out = new_act(xa.N, xa.H // 2, xa.W // 2, xa.C, x.device)
        ops.maxpool2x2(xa, Act(out))
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        xa = as_act(x)
        dx = new_act(xa.N, xa.H, xa.W, xa.C, x.device)
        if xa.H % 2 or xa.W % 2:       # floor pooling: the last row / column is outside every window (zero gradient)
            dx.zero_()
        ops.maxpool2x2_bwd(xa, as_act(dout), Act(dx))
        return dx


class Cat(torch.autograd.Function):
    """torch.cat along channels (models/vnet.py:127-150) with the strided channel-slice copy kernel.

    Cat.apply(dense, *xs): dense[i] says how input i receives its gradient. False: as a channel-slice view of the
    incoming gradient (no copy; every libb2s node reads slices in place). True: copied into a dense tensor with the
    slice-copy kernel -- for inputs with further consumers (the skip tensors), because autograd summing a strided view
    into a dense gradient goes through torch's non-vectorised elementwise kernel (310 us per add at 512^2, 4.7 % of the
    V-Net step) instead of the 16-byte one. dense=None: all True."""

    @staticmethod
    def forward(ctx, dense, *xs):
        acts = [as_act(t) for t in xs]
        a0 = acts[0]
        Ct = sum(a.C for a in acts)
        out = new_act(a0.N, a0.H, a0.W, Ct, xs[0].device)
        o = 0
        for a in acts:
            ops.copy_channels(a, Act(out, o, a.C))
            o += a.C
        ctx.sizes = [a.C for a in acts]
        ctx.dense = tuple(dense) if dense is not None else (True,) * len(xs)
        return out

    @staticmethod
    def backward(ctx, dout):
        da = as_act(dout)
        grads, o = [None], 0
        for i, c in enumerate(ctx.sizes):
            if not ctx.needs_input_grad[i + 1]:
                grads.append(None)
            elif ctx.dense[i]:
                g = new_act(da.N, da.H, da.W, c, dout.device)
                ops.copy_channels(da.slice(o, c), Act(g))
                grads.append(g)
            else:
                grads.append(dout[..., o:o + c])
            o += c
        return tuple(grads)


class Head(torch.autograd.Function):
    """final nn.Conv2d(C, num_classes, 1) (models/vnet.py:115): NHWC bf16 -> fp32 logits [N,O,H,W]"""

    @staticmethod
    def forward(ctx, x, w, b):
        xa = as_act(x)
        O = w.shape[0]
        logits = torch.empty((xa.N, O, xa.H, xa.W), dtype=torch.float32, device=w.device)
        ops.head_fwd(xa, None, None, w.detach().reshape(O, -1).contiguous(), b.detach() if b is not None else None,
                     logits, None)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        x, w = ctx.saved_tensors
        xa = as_act(x)
        O, C = w.shape[0], w.shape[1]
        dev = w.device
        f32 = dict(dtype=torch.float32, device=dev)
        dx = new_act(xa.N, xa.H, xa.W, C, dev)
        partial = torch.empty(ops.ew_rows() * (O * C + O), **f32)
        scratch = torch.empty(128 * (O * C + O), **f32)
        dwdb = torch.empty(O * C + O, **f32)
        ops.head_bwd(dlogits.contiguous().float(), xa, None, None, w.detach().reshape(O, C).contiguous(), Act(dx),
                     partial, scratch, dwdb)
        return dx, dwdb[:O * C].view(O, C, 1, 1).clone(), (dwdb[O * C:].clone() if ctx.has_bias else None)


class PsiGate(torch.autograd.Function):
    """psi = sigmoid(BatchNorm2d(1)(sum(maps)))      AttentionGate.psi after its 1x1 conv (models/mod.py:223-227,233)

    maps: the fp32 [N,1,H,W] outputs of the F_int -> 1 conv, one per channel slice of at most 256 channels (Head over a
    slice); the kernels sum them on the fly."""

    @staticmethod
    def forward(ctx, gamma, beta, rm, rv, nbt, training, *maps):
        maps = [m.contiguous() for m in maps]
        psi, mean, invstd = ops.psi_forward(maps, gamma.detach(), beta.detach(), rm, rv, nbt, training)
        ctx.training = training
        if training:
            ctx.save_for_backward(gamma, psi, mean, invstd, *maps)
        return psi

    @staticmethod
    def backward(ctx, dpsi):
        if not ctx.training:
            raise RuntimeError("b200seg AttentionGate: backward through eval-mode BatchNorm is not implemented")
        gamma, psi, mean, invstd, *maps = ctx.saved_tensors
        dv, dgamma, dbeta = ops.psi_backward(maps, psi, dpsi, mean, invstd, gamma.detach())
        return (dgamma, dbeta, None, None, None, None) + tuple(dv for _ in maps)


class PixelScale(torch.autograd.Function):
    """x * psi with psi [N,1,H,W] broadcast over the channels of the NHWC bf16 x (models/mod.py:234)"""

    @staticmethod
    def forward(ctx, x, psi):
        xa = as_act(x)
        out = new_act(xa.N, xa.H, xa.W, xa.C, x.device)
        psi = psi.contiguous()
        ops.pixel_scale_fwd(xa, psi, Act(out))
        ctx.save_for_backward(x, psi)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, psi = ctx.saved_tensors
        xa, dy = as_act(x), as_act(dout)
        dx = new_act(xa.N, xa.H, xa.W, xa.C, x.device)
        dpsi = torch.empty_like(psi)
        ops.pixel_scale_bwd(xa, psi, dy, Act(dx), dpsi)
        return dx, dpsi


class Bilinear(torch.autograd.Function):
    """F.interpolate(x, size=(Ho, Wo), mode='bilinear', align_corners=False) on NHWC bf16 (models/mod.py:61-62)"""

    @staticmethod
    def forward(ctx, x, Ho, Wo):
        xa = as_act(x)
        out = new_act(xa.N, Ho, Wo, xa.C, x.device)
        ops.bilinear_fwd(xa, Act(out))
        ctx.in_shape = (xa.N, xa.H, xa.W, xa.C)
        return out

    @staticmethod
    def backward(ctx, dout):
        N, H, W, C = ctx.in_shape
        dx = new_act(N, H, W, C, dout.device)
        ops.bilinear_bwd(as_act(dout), Act(dx))
        return dx, None, None


class ImageToAct(torch.autograd.Function):
    """fp32 image [N,C,H,W] with 1 < C <= 64 -> NHWC bf16 [N,H,W,64], channels >= C zero (no input gradient: the
    drop-in nets do not differentiate with respect to the image)"""

    @staticmethod
    def forward(ctx, x):
        return ops.image_to_nhwc(x.float(), 64).buf

    @staticmethod
    def backward(ctx, dout):
        return None
