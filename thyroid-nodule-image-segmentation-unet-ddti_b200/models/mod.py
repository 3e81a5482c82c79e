"""Drop-in replacements for two nets of the reference's ``models/mod.py``: ``UNet`` (Conv -> BN -> ReLU blocks,
``bias=False``, parametrised ``depth`` / ``base_filters``; reference models/mod.py:9-66) and ``ResUNet`` with its
``ResidualBlock`` (relu(BN(conv(relu(BN(conv x)))) + 1x1 skip); reference models/mod.py:71-131) — the model
``main.py:120-122`` actually runs. Same constructor signatures, attribute names and construction order, hence the same
``state_dict`` and seeded initialisation. The forwards chain the libb2s autograd nodes of ``vnet_functional.py`` over
NHWC bf16 activations; CUDA only, no CPU fallback. Skip concat order is [skip, upsampled] (models/mod.py:63,128).

The other nets of models/mod.py (ASPPUNet, AttentionUNet, TransUNet, VNet2D, ...) are out of scope.
"""
import torch
import torch.nn as nn

from .. import vnet_functional as VF


def _require(x, depth):
    if not x.is_cuda:
        raise RuntimeError("b200seg models run on CUDA (sm_100a) only; there is no CPU fallback")
    m = 1 << depth
    if x.shape[2] % m or x.shape[3] % m:
        # the reference falls back to F.interpolate when the up-sampled size differs from the skip (models/mod.py:61-62);
        # with sizes divisible by 2^depth the branch is never taken, which is the case implemented here
        raise RuntimeError(f"H and W must be multiples of {m} (2^depth); the bilinear re-size branch is not implemented")


def _conv_bn(x, res, conv, bn, training, relu_mode):
    return VF.ConvBnAct.apply(x, res, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                              bn.num_batches_tracked, training, 0.0, 0, relu_mode)


class _PackMixin:
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
        if in_channels != 1 or base_filters % 64:
            raise NotImplementedError("the B200 path implements in_channels=1 and base_filters a multiple of 64")

    def _block(self, in_ch, out_ch):
        return nn.Sequential(nn.Conv2d(in_ch, out_ch, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(out_ch),
                             nn.ReLU(inplace=True),
                             nn.Conv2d(out_ch, out_ch, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(out_ch),
                             nn.ReLU(inplace=True))

    def _run_block(self, blk, x):
        x = _conv_bn(x, None, blk[0], blk[1], self.training, 1)
        return _conv_bn(x, None, blk[3], blk[4], self.training, 1)

    def forward(self, x):
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise RuntimeError(f"expected input [N,{self.in_channels},H,W], got {tuple(x.shape)}")
        _require(x, self.depth)
        self._pack_all_weights()
        x = x.float().contiguous()
        skips = []
        for enc in self.encoders:
            x = self._run_block(enc, x)
            skips.append(x)
            x = VF.MaxPool2x2.apply(x)
        x = self._run_block(self.bottleneck, x)
        for up, dec, skip in zip(self.upconvs, self.decoders, reversed(skips)):
            x = VF.ConvT2x2.apply(x, up.weight, up.bias)
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

    def forward_nhwc(self, x):
        """x: NHWC bf16, or the fp32 image [N,1,H,W] when in_ch == 1"""
        s = VF.Conv1x1.apply(x, self.skip.weight, None)
        a = _conv_bn(x, None, self.conv[0], self.conv[1], self.training, 1)
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
        if in_channels != 1 or base_filters % 64:
            raise NotImplementedError("the B200 path implements in_channels=1 and base_filters a multiple of 64")

    def forward(self, x):
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise RuntimeError(f"expected input [N,{self.in_channels},H,W], got {tuple(x.shape)}")
        _require(x, self.depth)
        self._pack_all_weights()
        x = x.float().contiguous()
        skips = []
        for enc in self.encoders:
            x = enc.forward_nhwc(x)
            skips.append(x)
            x = VF.MaxPool2x2.apply(x)
        x = self.bottleneck.forward_nhwc(x)
        for up, dec, skip in zip(self.upconvs, self.decoders, reversed(skips)):
            x = VF.ConvT2x2.apply(x, up.weight, up.bias)
            x = VF.Cat.apply((True, False), skip, x)    # the skip also feeds the max-pool; the up-conv output only this
            x = dec.forward_nhwc(x)
        return VF.Head.apply(x, self.final_conv.weight, self.final_conv.bias)
