"""ORACLE (test infrastructure, not product code) — CPU restatement of the reference's models/mod.py UNet and ResUNet.

Only tests/ may import this module. Elementary tensor algebra from unet_oracle.py (shift + einsum convolutions, explicit
BatchNorm formulas); gradients come from autograd over these elementary ops. Pinned against the unmodified reference by
oracle/make_golden_mod.py -> tests/golden/mod_golden.pt.

Call sites restated (reference file:line):
  UNet._block / forward          models/mod.py:43-66   (conv(bias=False) -> BN -> ReLU, twice; cat([skip, x]))
  ResidualBlock.forward          models/mod.py:83-84   (relu(BN(conv(relu(BN(conv x)))) + skip_1x1(x)))
  ResUNet.forward                models/mod.py:119-131
  nn.MaxPool2d(2,2)              models/mod.py:27,104  (first maximum in row-major window order; odd sizes: floor)
  AttentionGate.forward          models/mod.py:229-234 (x * sigmoid(BN(conv1x1(relu(BN(W_g g) + BN(W_x x))))))
  AttentionUNet.forward          models/mod.py:281-295
  F.interpolate(bilinear)        models/mod.py:61-62,126-127,289-290 (align_corners=False; taken for odd sizes)
  in_channels > 1                models/mod.py:25      (the CUDA path stores the image and the first weight in bf16)
"""
import torch

from . import unet_oracle as O
from .vnet_oracle import bn_train_or_eval


def maxpool_first(x):
    """2x2/2 max-pool selecting the FIRST maximum of each window (differentiable through the selected element)"""
    N, C, H, W = x.shape
    x = x[:, :, :H // 2 * 2, :W // 2 * 2]          # floor: an odd last row / column is outside every window
    win = torch.stack([x[:, :, 0::2, 0::2], x[:, :, 0::2, 1::2], x[:, :, 1::2, 0::2], x[:, :, 1::2, 1::2]], dim=-1)
    best = win[..., 0]
    arg = torch.zeros_like(best, dtype=torch.long)
    for q in range(1, 4):
        better = win[..., q] > best
        best = torch.where(better, win[..., q], best)
        arg = torch.where(better, torch.full_like(arg, q), arg)
    return torch.gather(win, -1, arg.unsqueeze(-1)).squeeze(-1)


def _cbr(P, conv, bn, x, train, q, w_round, relu=True, stats_out=None):
    z = q(O.conv3x3(x, w_round(P[f"{conv}.weight"]), None))
    y = bn_train_or_eval(z, P, bn, train, stats_out)
    return torch.clamp_min(y, 0) if relu else y


def plain_block(P, prefix, x, train, q, first, stats_out=None):
    """models/mod.py:43-51"""
    wr = (lambda w: w) if first else q
    a = q(_cbr(P, f"{prefix}.0", f"{prefix}.1", x, train, q, wr, True, stats_out))
    return q(_cbr(P, f"{prefix}.3", f"{prefix}.4", a, train, q, q, True, stats_out))


def residual_block(P, prefix, x, train, q, first, stats_out=None):
    """models/mod.py:71-84"""
    wr = (lambda w: w) if first else q
    s = q(O.conv1x1(x, wr(P[f"{prefix}.skip.weight"]), None))
    a = q(_cbr(P, f"{prefix}.conv.0", f"{prefix}.conv.1", x, train, q, wr, True, stats_out))
    y = _cbr(P, f"{prefix}.conv.3", f"{prefix}.conv.4", a, train, q, q, False, stats_out)
    return q(torch.clamp_min(y + s, 0))


def bilinear_resize(x, Ho, Wo):
    """F.interpolate(x, size=(Ho, Wo), mode='bilinear', align_corners=False): source coordinate
    max(0, (o + 0.5) * in / out - 0.5), taps floor and min(floor + 1, in - 1), separable linear weights."""
    def taps(n_out, n_in):
        o = torch.arange(n_out, dtype=x.dtype, device=x.device)
        src = ((o + 0.5) * (n_in / n_out) - 0.5).clamp_min(0)
        i0 = src.floor().long().clamp_max(n_in - 1)
        i1 = (i0 + 1).clamp_max(n_in - 1)
        return i0, i1, src - i0.to(x.dtype)
    h0, h1, ah = taps(Ho, x.shape[2])
    w0, w1, aw = taps(Wo, x.shape[3])
    left, right = x.index_select(3, w0), x.index_select(3, w1)
    rows = left + aw.view(1, 1, 1, -1) * (right - left)
    top, bot = rows.index_select(2, h0), rows.index_select(2, h1)
    return top + ah.view(1, 1, -1, 1) * (bot - top)


def _c1bn(P, conv, bn, x, train, q, stats_out):
    """1x1 conv (+bias) -> BatchNorm (models/mod.py:214-221); the CUDA path stores the conv output in bf16"""
    z = q(O.conv1x1(x, q(P[f"{conv}.weight"]), P[f"{conv}.bias"]))
    return bn_train_or_eval(z, P, bn, train, stats_out)


def attention_gate(P, prefix, g, x, train, q, stats_out=None):
    """models/mod.py:229-234"""
    g1 = q(_c1bn(P, f"{prefix}.W_g.0", f"{prefix}.W_g.1", g, train, q, stats_out))
    s = q(torch.clamp_min(_c1bn(P, f"{prefix}.W_x.0", f"{prefix}.W_x.1", x, train, q, stats_out) + g1, 0))
    a = O.conv1x1(s, P[f"{prefix}.psi.0.weight"], P[f"{prefix}.psi.0.bias"])          # fp32 weights, fp32 map
    psi = O.sigmoid(bn_train_or_eval(a, P, f"{prefix}.psi.1", train, stats_out))
    return q(x * psi)


def _net(P, x, depth, train, q, block, stats_out, gates=False):
    if x.shape[1] > 1:
        x = q(x)                      # multi-channel images are stored as NHWC bf16 before the first conv
    skips = []
    for i in range(depth):
        x = block(P, f"encoders.{i}", x, train, q, i == 0 and x.shape[1] == 1, stats_out)
        skips.append(x)
        x = maxpool_first(x)
    x = block(P, "bottleneck", x, train, q, False, stats_out)
    for i, skip in enumerate(reversed(skips)):
        x = q(O.conv_transpose2x2(x, q(P[f"upconvs.{i}.weight"]), P[f"upconvs.{i}.bias"]))
        if x.shape[2:] != skip.shape[2:]:
            x = q(bilinear_resize(x, skip.shape[2], skip.shape[3]))
        if gates:
            skip = attention_gate(P, f"attn_gates.{i}", x, skip, train, q, stats_out)
        x = torch.cat([skip, x], dim=1)
        x = block(P, f"decoders.{i}", x, train, q, False, stats_out)
    return O.conv1x1(x, P["final_conv.weight"], P["final_conv.bias"])


def resunet_forward(P, x, depth, train=True, q=O.identity, stats_out=None):
    return _net(P, x, depth, train, q, residual_block, stats_out)


def unet_forward(P, x, depth, train=True, q=O.identity, stats_out=None):
    return _net(P, x, depth, train, q, plain_block, stats_out)


def attention_unet_forward(P, x, depth, train=True, q=O.identity, stats_out=None):
    return _net(P, x, depth, train, q, plain_block, stats_out, gates=True)
