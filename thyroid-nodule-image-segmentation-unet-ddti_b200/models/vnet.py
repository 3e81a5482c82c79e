"""Drop-in replacement for the reference's ``models/vnet.py``: ``SEBlock``, ``ConvBlock``, ``ImprovedVNet``.

Same constructor signatures, attribute names and construction order as the reference (models/vnet.py:5-115), hence
the same ``state_dict`` keys/shapes (455 entries, 160 435 681 parameters at the defaults) and bit-identical default
initialisation under a given ``torch.manual_seed``. The forward passes (models/vnet.py:18-26, 48-60, 117-155) do not
run the torch layers: they chain the libb2s autograd nodes of ``vnet_functional.py`` over NHWC bf16 activations.
CUDA tensors only; there is no CPU fallback.

Dropout (p = 0.05 by default, models/vnet.py:38,55,67) uses the library's own counter-based mask, which cannot
reproduce torch's Philox stream: parity tests construct the net with ``dropout_rate=0.0`` or compare in ``eval()``
(SURVEY.md App. B.6).
"""
import itertools

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import vnet_functional as VF
from ..ops import BF16

_seed_counter = itertools.count(1)


def _next_seed():
    return (torch.initial_seed() * 2654435761 + next(_seed_counter) * 40503) & 0xFFFFFFFF


def _to_nhwc(x):
    """API boundary: NCHW float -> NHWC bf16 (layout/cast only)"""
    return x.permute(0, 2, 3, 1).contiguous().to(BF16)


def _to_nchw(y):
    return y.permute(0, 3, 1, 2).float().contiguous()


def _require_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("b200seg V-Net runs on CUDA (sm_100a) only; there is no CPU fallback")


class SEBlock(nn.Module):
    """Squeeze-and-excitation (reference models/vnet.py:5-26)."""

    def __init__(self, channels, reduction=4):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(channels, channels // reduction, kernel_size=1)
        self.fc2 = nn.Conv2d(channels // reduction, channels, kernel_size=1)
        self.relu = nn.ReLU(inplace=True)
        self.sigmoid = nn.Sigmoid()

    def forward_nhwc(self, x):
        return VF.SE.apply(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)

    def forward(self, x):
        _require_cuda(x)
        return _to_nchw(self.forward_nhwc(_to_nhwc(x)))


class ConvBlock(nn.Module):
    """num_convs x (Conv3x3 -> BatchNorm -> ReLU -> Dropout) + residual (1x1 projection when the channel counts
    differ) (reference models/vnet.py:28-60)."""

    def __init__(self, in_channels, out_channels, num_convs, dropout_rate):
        super().__init__()
        self.convs = nn.ModuleList()
        self.bns = nn.ModuleList()
        self.relu = nn.ReLU(inplace=True)
        self.drop = nn.Dropout(dropout_rate)
        for i in range(num_convs):
            conv_in = in_channels if i == 0 else out_channels
            self.convs.append(nn.Conv2d(conv_in, out_channels, kernel_size=3, stride=1, padding=1))
            self.bns.append(nn.BatchNorm2d(out_channels))
        self.res_proj = nn.Conv2d(in_channels, out_channels, kernel_size=1) if in_channels != out_channels else None

    def _layer_seed(self, i):
        """fixed per (process seed, block, conv): the per-step variation comes from the device-side step counter"""
        if not hasattr(self, "_seed_base"):
            self._seed_base = _next_seed()
        return (self._seed_base + 7919 * i) & 0xFFFFFFFF

    def forward_nhwc(self, x):
        """x: NHWC bf16, or the fp32 image [N,1,H,W] when in_channels == 1. For 1 < in_channels < 64 the caller passes
        the image zero-padded to 64 NHWC channels (VF.ImageToAct) and the weights reading it are zero-padded here (the
        gradient of the real input channels flows back through autograd's slice)."""
        cin = self.convs[0].weight.shape[1]
        pad = (lambda w: F.pad(w, (0, 0, 0, 0, 0, 64 - cin))) if 1 < cin < 64 else (lambda w: w)
        if self.res_proj is not None:
            residual = VF.Conv1x1.apply(x, pad(self.res_proj.weight), self.res_proj.bias)
        else:
            residual = x
        p = float(self.drop.p)
        n = len(self.convs)
        for i, (conv, bn) in enumerate(zip(self.convs, self.bns)):
            res = residual if i == n - 1 else None       # the residual add is fused into the last stage's apply pass
            w = pad(conv.weight) if i == 0 else conv.weight
            x = VF.ConvBnAct.apply(x, res, w, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                   bn.num_batches_tracked, self.training, p, self._layer_seed(i) if (self.training and p > 0) else 0, 1)
        return x

    def forward(self, x):
        _require_cuda(x)
        if x.shape[1] == 1:
            return _to_nchw(self.forward_nhwc(x.float().contiguous()))
        return _to_nchw(self.forward_nhwc(_to_nhwc(x)))


class ImprovedVNet(nn.Module):
    """Three-branch encoder with SE blocks and one shared decoder (reference models/vnet.py:62-155)."""

    def __init__(self, in_channels=1, num_classes=1, base_num_filters=64, dropout_rate=0.05, se_reduction=4):
        super().__init__()
        self.num_branches = 3
        self.in_channels = in_channels
        filters = [base_num_filters * (2 ** i) for i in range(5)]
        self.enc_blocks = nn.ModuleList([nn.ModuleList() for _ in range(self.num_branches)])
        self.enc_ses = nn.ModuleList([nn.ModuleList() for _ in range(self.num_branches)])
        self.down_convs = nn.ModuleList([nn.ModuleList() for _ in range(self.num_branches)])
        enc_conv_counts = [2, 2, 3, 3, 3]
        for b in range(self.num_branches):
            for i in range(5):
                in_ch = in_channels if i == 0 else filters[i]
                out_ch = filters[i]
                self.enc_blocks[b].append(ConvBlock(in_ch, out_ch, enc_conv_counts[i], dropout_rate))
                self.enc_ses[b].append(SEBlock(out_ch, reduction=se_reduction))
                if i < 4:
                    self.down_convs[b].append(nn.Conv2d(out_ch, filters[i + 1], kernel_size=3, stride=2, padding=1))
        nb = self.num_branches
        self.up6 = nn.ConvTranspose2d(filters[4] * nb, filters[3], kernel_size=2, stride=2)
        self.up7 = nn.ConvTranspose2d(filters[3], filters[2], kernel_size=2, stride=2)
        self.up8 = nn.ConvTranspose2d(filters[2], filters[1], kernel_size=2, stride=2)
        self.up9 = nn.ConvTranspose2d(filters[1], filters[0], kernel_size=2, stride=2)
        self.dec_blocks = nn.ModuleList([
            ConvBlock(filters[3] + filters[3] * nb, filters[3], num_convs=3, dropout_rate=dropout_rate),
            ConvBlock(filters[2] + filters[2] * nb, filters[2], num_convs=3, dropout_rate=dropout_rate),
            ConvBlock(filters[1] + filters[1] * nb, filters[1], num_convs=2, dropout_rate=dropout_rate),
            ConvBlock(filters[0] + filters[0] * nb, filters[0], num_convs=2, dropout_rate=dropout_rate),
        ])
        self.dec_se_final = SEBlock(filters[0], reduction=se_reduction)
        self.final_conv = nn.Conv2d(filters[0], num_classes, kernel_size=1)
        if not 1 <= in_channels <= 64 or base_num_filters % 64:
            raise NotImplementedError("the B200 path implements 1 <= in_channels <= 64 and base_num_filters a multiple "
                                      "of 64")

    def invalidate_packed(self):
        """Forces the next forward to re-pack the bf16 GEMM operands (needed only after a weight update that bypasses
        torch's version counters, e.g. `p.data.copy_()` or an external kernel)."""
        self._pack_key = None

    def _pack_all_weights(self):
        """bf16 GEMM operands of every tensor-core conv / transposed-conv weight, refreshed in a few launches whenever
        a parameter changed (optimizer step, load_state_dict, .to()); registered for the autograd nodes."""
        from .. import ops
        ws = [(n, p) for n, p in self.named_parameters()
              if p.dim() == 4 and p.shape[1] >= 64 and p.shape[0] >= 64 and p.shape[0] % 64 == 0 and p.shape[1] % 64 == 0
              and ".fc" not in n]
        # self.training is part of the key: it decides whether the dgrad operands are packed (an eval forward between
        # two train forwards must not leave the train step without them)
        key = (self.training,) + tuple((id(p), p.data_ptr(), p._version) for _, p in ws)
        if not ws:
            return        # nn.DataParallel replica (named_parameters() is empty there): the nodes pack on the spot
        if getattr(self, "_pack_key", None) == key:
            return
        layout = tuple((p.data_ptr(), self.training) for _, p in ws)
        if getattr(self, "_pack_layout", None) != layout:
            items = [(n, p.detach(), n.startswith("up")) for n, p in ws]
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

    def _trunk(self, x):
        assert x.shape[1] == self.in_channels, f"Expected input with {self.in_channels} channel(s)"
        _require_cuda(x)
        self._pack_all_weights()
        if x.shape[2] % 16 or x.shape[3] % 16:
            raise RuntimeError("Sizes of tensors must match except in dimension 1: H and W must be multiples of 16")
        x = x.float().contiguous()
        if self.in_channels > 1:
            x = VF.ImageToAct.apply(x)
        nb = self.num_branches
        feats = [[None] * 5 for _ in range(nb)]
        for b in range(nb):
            e = x
            for i in range(5):
                e = self.enc_blocks[b][i].forward_nhwc(e)
                e = self.enc_ses[b][i].forward_nhwc(e)
                feats[b][i] = e
                if i < 4:
                    dc = self.down_convs[b][i]
                    e = VF.ConvS2.apply(e, dc.weight, dc.bias)
        d = VF.Cat.apply((False,) * nb, *[feats[b][4] for b in range(nb)])     # bottleneck features: one consumer each
        for lvl, up, blk in ((3, self.up6, 0), (2, self.up7, 1), (1, self.up8, 2), (0, self.up9, 3)):
            d = VF.ConvT2x2.apply(d, up.weight, up.bias)
            # the up-conv output is consumed here only; the skips also feed their branch's down conv
            d = VF.Cat.apply((False,) + (True,) * nb, d, *[feats[b][lvl] for b in range(nb)])
            d = self.dec_blocks[blk].forward_nhwc(d)
        return self.dec_se_final.forward_nhwc(d)

    @VF.ops.on_device_of_input
    def forward(self, x):
        return VF.Head.apply(self._trunk(x), self.final_conv.weight, self.final_conv.bias)

    @torch.no_grad()
    @VF.ops.on_device_of_input
    def predict_mask(self, x):
        """Inference entry (utils/trainer.py:216-217): (logits, uint8 mask) with mask = sigmoid(logits) > 0.5 fused
        into the head kernel."""
        from .. import ops
        d = VF.as_act(self._trunk(x))
        O = self.final_conv.weight.shape[0]
        logits = torch.empty((d.N, O, d.H, d.W), dtype=torch.float32, device=x.device)
        mask = torch.empty((d.N, O, d.H, d.W), dtype=torch.uint8, device=x.device)
        ops.head_fwd(d, None, None, self.final_conv.weight.detach().reshape(O, -1).contiguous(),
                     self.final_conv.bias.detach(), logits, mask)
        return logits, mask
