"""Drop-in replacements for two nets of the reference's ``models/mod.py``: ``UNet`` (Conv -> BN -> ReLU blocks,
``bias=False``, parametrised ``depth`` / ``base_filters``; reference models/mod.py:9-66) and ``ResUNet`` with its
``ResidualBlock`` (relu(BN(conv(relu(BN(conv x)))) + 1x1 skip); reference models/mod.py:71-131) — the model
``main.py:120-122`` actually runs. Same constructor signatures, attribute names and construction order, hence the same
``state_dict`` and seeded initialisation. The forwards chain the libb2s autograd nodes of ``vnet_functional.py`` over
NHWC bf16 activations; CUDA only, no CPU fallback. Skip concat order is [skip, upsampled] (models/mod.py:63,128).

``AttentionUNet`` with its ``AttentionGate`` (models/mod.py:211-295) runs on the same nodes plus the one-channel
BatchNorm + sigmoid and per-pixel scale kernels of csrc/attn_ops.cu. The bilinear re-size branch
(``F.interpolate(..., mode='bilinear')`` when H or W is not a multiple of 2^depth, models/mod.py:61-62,126-127,289-290)
is ``b2s_bilinear_fwd/bwd``; ``in_channels > 1`` (up to 64) takes the tensor-core first conv on a zero-padded NHWC image.

The other nets of models/mod.py (ASPPUNet, TransUNet, VNet2D, ...) are out of scope.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import vnet_functional as VF


def _require(x, depth):
    if not x.is_cuda:
        raise RuntimeError("b200seg models run on CUDA (sm_100a) only; there is no CPU fallback")
    if x.requires_grad:
        raise RuntimeError("b200seg models do not compute the gradient with respect to the input image")
    if (x.shape[2] >> depth) < 1 or (x.shape[3] >> depth) < 1:
        raise RuntimeError(f"input {tuple(x.shape[2:])} is too small for {depth} 2x2 poolings")


def _check_channels(in_channels, base_filters):
    if not 1 <= in_channels <= 64 or base_filters % 64:
        raise NotImplementedError("the B200 path implements 1 <= in_channels <= 64 and base_filters a multiple of 64")


def _image(x, in_channels):
    """the network input as the first conv takes it: the fp32 image itself (Cin = 1 kernel) or NHWC bf16 zero-padded
    to 64 channels (tensor-core kernel with a zero-padded weight)"""
    x = x.float().contiguous()
    return x if in_channels == 1 else VF.ImageToAct.apply(x)


def _pad_cin(w, in_channels):
    """first-conv weight [Cout,Cin,k,k] -> [Cout,64,k,k] (zeros), differentiable: the gradient of the real channels
    flows back through the slice"""
    return w if in_channels == 1 else F.pad(w, (0, 0, 0, 0, 0, 64 - w.shape[1]))


def _conv_bn(x, res, conv, bn, training, relu_mode, weight=None):
    return VF.ConvBnAct.apply(x, res, conv.weight if weight is None else weight, conv.bias, bn.weight, bn.bias,
                              bn.running_mean, bn.running_var, bn.num_batches_tracked, training, 0.0, 0, relu_mode)


def _match_size(x, skip):
    """models/mod.py:61-62: the up-sampled tensor is re-sized to the skip's spatial size when they differ"""
    if x.shape[1:3] != skip.shape[1:3]:
        x = VF.Bilinear.apply(x, skip.shape[1], skip.shape[2])
    return x


class _PackMixin:
    def invalidate_packed(self):
        """Forces the next forward to re-pack the bf16 GEMM operands (needed only after a weight update that bypasses
        torch's version counters, e.g. `p.data.copy_()` or an external kernel)."""
        self._pack_key = None

    def _pack_all_weights(self):
        """bf16 operands of every tensor-core conv weight, refreshed in a few launches when a parameter changed"""
        from .. import ops
        ws = [(n, p) for n, p in self.named_parameters()
              if p.dim() == 4 and p.shape[0] >= 64 and p.shape[1] >= 64 and p.shape[0] % 64 == 0 and p.shape[1] % 64 == 0]
        # self.training is part of the key: it decides whether the dgrad operands are packed (an eval forward between
        # two train forwards must not leave the train step without them)
        key = (self.training,) + tuple((id(p), p.data_ptr(), p._version) for _, p in ws)
        if not ws:
            return        # nn.DataParallel replica (named_parameters() is empty there): the nodes pack on the spot
        if getattr(self, "_pack_key", None) == key:
            return
        layout = tuple((p.data_ptr(), self.training) for _, p in ws)
        if getattr(self, "_pack_layout", None) != layout:
            items = [(n, p.detach(), n.startswith("upconvs")) for n, p in ws]
            self._pack_plans = [ops.PackPlan(items[i:i + 40], want_dgrad=self.training) for i in range(0, len(items), 40)]
            self._pack_layout = layout
        for old in getattr(self, "_pack_registered", ()):
            VF.PACKED.pop(old, None)
        reg = []
        for plan in self._pack_plans:
            plan.run()
        for n, p in ws:
            for plan in self._pack_plans:
                if n in plan.packed:
                    VF.register_packed(p, *plan.packed[n])
                    reg.append(p.data_ptr())
        self._pack_registered = tuple(reg)
        self._pack_key = key


class UNet(nn.Module, _PackMixin):
    """reference models/mod.py:9-66"""

    def __init__(self, in_channels: int = 1, out_channels: int = 1, base_filters: int = 64, depth: int = 5, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.base_filters, self.depth = base_filters, depth
        self.encoders, self.pools = nn.ModuleList(), nn.ModuleList()
        prev_ch = in_channels
        channels = [base_filters * (2 ** i) for i in range(depth)]
        for ch in channels:
            self.encoders.append(self._block(prev_ch, ch))
            self.pools.append(nn.MaxPool2d(2, 2))
            prev_ch = ch
        self.bottleneck = self._block(prev_ch, prev_ch * 2)
        self.upconvs, self.decoders = nn.ModuleList(), nn.ModuleList()
        prev_ch = channels[-1] * 2
        for ch in channels[::-1]:
            self.upconvs.append(nn.ConvTranspose2d(prev_ch, ch, kernel_size=2, stride=2))
            self.decoders.append(self._block(prev_ch, ch))
            prev_ch = ch
        self.final_conv = nn.Conv2d(base_filters, out_channels, kernel_size=1)
        _check_channels(in_channels, base_filters)

    def _block(self, in_ch, out_ch):
        return nn.Sequential(nn.Conv2d(in_ch, out_ch, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(out_ch),
                             nn.ReLU(inplace=True),
                             nn.Conv2d(out_ch, out_ch, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(out_ch),
                             nn.ReLU(inplace=True))

    def _run_block(self, blk, x, first=False):
        w0 = _pad_cin(blk[0].weight, self.in_channels) if first else None
        x = _conv_bn(x, None, blk[0], blk[1], self.training, 1, weight=w0)
        return _conv_bn(x, None, blk[3], blk[4], self.training, 1)

    @VF.ops.on_device_of_input
    def forward(self, x):
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise RuntimeError(f"expected input [N,{self.in_channels},H,W], got {tuple(x.shape)}")
        _require(x, self.depth)
        self._pack_all_weights()
        x = _image(x, self.in_channels)
        skips = []
        for i, enc in enumerate(self.encoders):
            x = self._run_block(enc, x, first=(i == 0))
            skips.append(x)
            x = VF.MaxPool2x2.apply(x)
        x = self._run_block(self.bottleneck, x)
        for up, dec, skip in zip(self.upconvs, self.decoders, reversed(skips)):
            x = _match_size(VF.ConvT2x2.apply(x, up.weight, up.bias), skip)
            x = VF.Cat.apply((True, False), skip, x)    # the skip also feeds the max-pool; the up-conv output only this
            x = self._run_block(dec, x)
        return VF.Head.apply(x, self.final_conv.weight, self.final_conv.bias)


class ResidualBlock(nn.Module):
    """reference models/mod.py:71-84"""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_ch, out_ch, 3, padding=1, bias=False), nn.BatchNorm2d(out_ch),
                                  nn.ReLU(inplace=True),
                                  nn.Conv2d(out_ch, out_ch, 3, padding=1, bias=False), nn.BatchNorm2d(out_ch))
        self.skip = nn.Conv2d(in_ch, out_ch, 1, bias=False)
        self.relu = nn.ReLU(inplace=True)

    def forward_nhwc(self, x, in_channels=1):
        """x: NHWC bf16, or the fp32 image [N,1,H,W] when in_ch == 1; in_channels > 1: x is the image zero-padded to
        64 NHWC channels and both convs reading it use zero-padded weights"""
        pad = 1 < self.skip.weight.shape[1] < 64
        s = VF.Conv1x1.apply(x, _pad_cin(self.skip.weight, in_channels) if pad else self.skip.weight, None)
        a = _conv_bn(x, None, self.conv[0], self.conv[1], self.training, 1,
                     weight=_pad_cin(self.conv[0].weight, in_channels) if pad else None)
        return _conv_bn(a, s, self.conv[3], self.conv[4], self.training, 2)      # relu(BN(conv(a)) + skip(x))

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("b200seg models run on CUDA (sm_100a) only; there is no CPU fallback")
        xin = x.float().contiguous() if x.shape[1] == 1 else x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        return self.forward_nhwc(xin).permute(0, 3, 1, 2).float().contiguous()


class ResUNet(nn.Module, _PackMixin):
    """reference models/mod.py:86-131"""

    def __init__(self, in_channels: int = 1, out_channels: int = 1, base_filters: int = 64, depth: int = 5, **kwargs):
        super().__init__()
        self.base_filters, self.depth = base_filters, depth
        self.in_channels = in_channels
        self.encoders, self.pools = nn.ModuleList(), nn.ModuleList()
        prev_ch = in_channels
        channels = [base_filters * (2 ** i) for i in range(depth)]
        for ch in channels:
            self.encoders.append(ResidualBlock(prev_ch, ch))
            self.pools.append(nn.MaxPool2d(2, 2))
            prev_ch = ch
        self.bottleneck = ResidualBlock(prev_ch, prev_ch * 2)
        self.upconvs, self.decoders = nn.ModuleList(), nn.ModuleList()
        prev_ch = channels[-1] * 2
        for ch in channels[::-1]:
            self.upconvs.append(nn.ConvTranspose2d(prev_ch, ch, 2, 2))
            self.decoders.append(ResidualBlock(prev_ch, ch))
            prev_ch = ch
        self.final_conv = nn.Conv2d(base_filters, out_channels, 1)
        _check_channels(in_channels, base_filters)

    @VF.ops.on_device_of_input
    def forward(self, x):
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise RuntimeError(f"expected input [N,{self.in_channels},H,W], got {tuple(x.shape)}")
        _require(x, self.depth)
        self._pack_all_weights()
        x = _image(x, self.in_channels)
        skips = []
        for i, enc in enumerate(self.encoders):
            x = enc.forward_nhwc(x, self.in_channels if i == 0 else 1)
            skips.append(x)
            x = VF.MaxPool2x2.apply(x)
        x = self.bottleneck.forward_nhwc(x)
        for up, dec, skip in zip(self.upconvs, self.decoders, reversed(skips)):
            x = _match_size(VF.ConvT2x2.apply(x, up.weight, up.bias), skip)
            x = VF.Cat.apply((True, False), skip, x)    # the skip also feeds the max-pool; the up-conv output only this
            x = dec.forward_nhwc(x)
        return VF.Head.apply(x, self.final_conv.weight, self.final_conv.bias)


class AttentionGate(nn.Module):
    """reference models/mod.py:211-234: psi = sigmoid(BN(conv1x1(relu(BN(W_g g) + BN(W_x x))))), returns x * psi.

    F_int < 64 (the gate of the 64-channel level has F_int = 32) runs the two 1x1 convs with their output channels
    zero-padded to 64: the padded channels carry exact zeros through BatchNorm (gamma 1, beta 0 on a constant-zero
    channel), the ReLU and the zero-padded psi conv, and their gradients are dropped by the slices autograd applies to
    the padded parameters."""

    def __init__(self, F_g, F_l, F_int):
        super().__init__()
        self.W_g = nn.Sequential(nn.Conv2d(F_g, F_int, 1, bias=True), nn.BatchNorm2d(F_int))
        self.W_x = nn.Sequential(nn.Conv2d(F_l, F_int, 1, bias=True), nn.BatchNorm2d(F_int))
        self.psi = nn.Sequential(nn.Conv2d(F_int, 1, 1, bias=True), nn.BatchNorm2d(1), nn.Sigmoid())
        self.relu = nn.ReLU(inplace=True)
        self.F_int = F_int
        if F_int % 64 and (F_int > 64 or 64 % F_int):
            raise NotImplementedError("the B200 path implements F_int = 32 or a multiple of 64")

    def _branch(self, x, res, seq, relu_mode):
        conv, bn = seq[0], seq[1]
        F_int = self.F_int
        if F_int % 64 == 0:
            return VF.ConvBnAct.apply(x, res, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                      bn.num_batches_tracked, self.training, 0.0, 0, relu_mode)
        pad = 64 - F_int
        w = F.pad(conv.weight, (0, 0, 0, 0, 0, 0, 0, pad))
        b = F.pad(conv.bias, (0, pad))
        gamma = F.pad(bn.weight, (0, pad), value=1.0)
        beta = F.pad(bn.bias, (0, pad))
        with torch.no_grad():       # padded running statistics live for this call only; the real ones are copied back
            rm = F.pad(bn.running_mean, (0, pad))
            rv = F.pad(bn.running_var, (0, pad), value=1.0)
        out = VF.ConvBnAct.apply(x, res, w, b, gamma, beta, rm, rv, bn.num_batches_tracked, self.training, 0.0, 0,
                                 relu_mode)
        if self.training:
            with torch.no_grad():
                bn.running_mean.copy_(rm[:F_int])
                bn.running_var.copy_(rv[:F_int])
        return out

    def forward_nhwc(self, g, x):
        g1 = self._branch(g, None, self.W_g, 0)                  # BN(W_g g)
        s = self._branch(x, g1, self.W_x, 2)                     # relu(BN(W_x x) + g1)
        conv, bn = self.psi[0], self.psi[1]
        w = conv.weight if self.F_int % 64 == 0 else F.pad(conv.weight, (0, 0, 0, 0, 0, 64 - self.F_int))
        C = w.shape[1]
        maps = []
        for c0 in range(0, C, 256):                              # the Cout = 1 conv kernel takes at most 256 channels
            c1 = min(C, c0 + 256)
            maps.append(VF.Head.apply(s if C <= 256 else s[..., c0:c1], w if C <= 256 else w[:, c0:c1],
                                      conv.bias if c0 == 0 else None))
        psi = VF.PsiGate.apply(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                               self.training, *maps)
        return VF.PixelScale.apply(x, psi)

    def forward(self, g, x):
        if not x.is_cuda:
            raise RuntimeError("b200seg models run on CUDA (sm_100a) only; there is no CPU fallback")
        to = lambda t: t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        return self.forward_nhwc(to(g), to(x)).permute(0, 3, 1, 2).float().contiguous()


class AttentionUNet(nn.Module, _PackMixin):
    """reference models/mod.py:236-295"""

    def __init__(self, in_channels: int = 1, out_channels: int = 1, base_filters: int = 64, depth: int = 5, **kwargs):
        super().__init__()
        self.base_filters, self.depth = base_filters, depth
        self.in_channels = in_channels
        self.encoders, self.pools = nn.ModuleList(), nn.ModuleList()
        prev_ch = in_channels
        channels = [base_filters * (2 ** i) for i in range(depth)]
        for ch in channels:
            self.encoders.append(self._block(prev_ch, ch))
            self.pools.append(nn.MaxPool2d(2, 2))
            prev_ch = ch
        self.bottleneck = self._block(prev_ch, prev_ch * 2)
        self.upconvs, self.attn_gates, self.decoders = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        prev_ch = channels[-1] * 2
        for ch in channels[::-1]:
            self.upconvs.append(nn.ConvTranspose2d(prev_ch, ch, 2, 2))
            self.attn_gates.append(AttentionGate(F_g=ch, F_l=ch, F_int=ch // 2))
            self.decoders.append(self._block(prev_ch, ch))
            prev_ch = ch
        self.final_conv = nn.Conv2d(base_filters, out_channels, 1)
        _check_channels(in_channels, base_filters)

    _block = UNet._block

    def _run_block(self, blk, x, first=False):
        w0 = _pad_cin(blk[0].weight, self.in_channels) if first else None
        x = _conv_bn(x, None, blk[0], blk[1], self.training, 1, weight=w0)
        return _conv_bn(x, None, blk[3], blk[4], self.training, 1)

    @VF.ops.on_device_of_input
    def forward(self, x):
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise RuntimeError(f"expected input [N,{self.in_channels},H,W], got {tuple(x.shape)}")
        _require(x, self.depth)
        self._pack_all_weights()
        x = _image(x, self.in_channels)
        skips = []
        for i, enc in enumerate(self.encoders):
            x = self._run_block(enc, x, first=(i == 0))
            skips.append(x)
            x = VF.MaxPool2x2.apply(x)
        x = self._run_block(self.bottleneck, x)
        for up, gate, dec, skip in zip(self.upconvs, self.attn_gates, self.decoders, reversed(skips)):
            x = _match_size(VF.ConvT2x2.apply(x, up.weight, up.bias), skip)
            skip_att = gate.forward_nhwc(x, skip)
            x = VF.Cat.apply((False, True), skip_att, x)     # x also feeds the gate: its gradient arrives dense
            x = self._run_block(dec, x)
        return VF.Head.apply(x, self.final_conv.weight, self.final_conv.bias)
