"""ORACLE (test infrastructure, not product code) — CPU restatement of the reference's V-Net variant.

Only tests/ may import this module. Restates models/vnet.py with the elementary tensor algebra of unet_oracle.py
(shift + einsum convolutions, explicit BatchNorm formulas) — no F.conv2d / F.batch_norm / nn.Module calls. The
forward graph is differentiable torch tensor algebra, so gradients of the restatement come from autograd over these
elementary ops (not over the reference's layers). Pinned against the unmodified reference ImprovedVNet by
oracle/make_golden_vnet.py -> tests/golden/vnet_golden.pt.

Call sites restated (reference file:line):
  SEBlock.forward             models/vnet.py:18-26
  ConvBlock.forward           models/vnet.py:48-60   (conv -> BN -> ReLU -> Dropout, + residual / 1x1 projection)
  stride-2 down conv          models/vnet.py:97
  ImprovedVNet.forward        models/vnet.py:117-155 (3 branches, bottom concat, 4 x [up-conv, 4-way concat, block])

`q` emulates the CUDA path's storage rounding (bf16 activations and bf16 GEMM weight operands).
"""
import torch

from . import unet_oracle as O


def conv3x3_s2(x, w, b=None):
    """nn.Conv2d(k=3, stride=2, padding=1): the stride-1 result sampled at even output positions."""
    return O.conv3x3(x, w, b)[:, :, ::2, ::2]


def se_block(x, w1, b1, w2, b2):
    """models/vnet.py:18-26; w1 [C/r,C,1,1], w2 [C,C/r,1,1]"""
    z = x.mean(dim=(2, 3))
    h = torch.clamp_min(z @ w1[:, :, 0, 0].t() + b1, 0)
    g = O.sigmoid(h @ w2[:, :, 0, 0].t() + b2)
    return x * g[:, :, None, None]


def bn_train_or_eval(z, P, name, train, stats_out=None):
    if train:
        mean, var = O.batchnorm_stats(z)
        if stats_out is not None:
            stats_out[name] = (mean.detach(), var.detach(), z.shape[0] * z.shape[2] * z.shape[3])
    else:
        mean, var = P[f"{name}.running_mean"], P[f"{name}.running_var"]
    scale, shift, _ = O.batchnorm_affine(mean, var, P[f"{name}.weight"], P[f"{name}.bias"])
    return z * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


def conv_block(P, prefix, x, num_convs, train, q, wq, first, stats_out=None):
    """models/vnet.py:48-60 with dropout disabled (p = 0 or eval). `first`: x is the fp32 image (weights not rounded)."""
    residual = x
    for i in range(num_convs):
        w = P[f"{prefix}.convs.{i}.weight"]
        w = w if (first and i == 0) else wq(w)
        z = q(O.conv3x3(x, w, P[f"{prefix}.convs.{i}.bias"]))
        a = torch.clamp_min(bn_train_or_eval(z, P, f"{prefix}.bns.{i}", train, stats_out), 0)
        if i < num_convs - 1:
            x = q(a)
        else:
            x = a
    if f"{prefix}.res_proj.weight" in P:
        wr = P[f"{prefix}.res_proj.weight"]
        residual = q(O.conv1x1(residual, wr if first else wq(wr), P[f"{prefix}.res_proj.bias"]))
    return q(x + residual)


def vnet_forward(P, x, train=True, q=O.identity, stats_out=None):
    """P: name -> tensor with the reference's state_dict keys. Returns logits [N,num_classes,H,W]."""
    wq = q
    counts = [2, 2, 3, 3, 3]
    feats = [[None] * 5 for _ in range(3)]
    for b in range(3):
        e = x
        for i in range(5):
            e = conv_block(P, f"enc_blocks.{b}.{i}", e, counts[i], train, q, wq, first=(i == 0), stats_out=stats_out)
            s = f"enc_ses.{b}.{i}"
            e = q(se_block(e, P[f"{s}.fc1.weight"], P[f"{s}.fc1.bias"], P[f"{s}.fc2.weight"], P[f"{s}.fc2.bias"]))
            feats[b][i] = e
            if i < 4:
                e = q(conv3x3_s2(e, wq(P[f"down_convs.{b}.{i}.weight"]), P[f"down_convs.{b}.{i}.bias"]))
    d = torch.cat([feats[b][4] for b in range(3)], dim=1)
    for lvl, up, blk, n in ((3, "up6", 0, 3), (2, "up7", 1, 3), (1, "up8", 2, 2), (0, "up9", 3, 2)):
        d = q(O.conv_transpose2x2(d, wq(P[f"{up}.weight"]), P[f"{up}.bias"]))
        d = torch.cat([d] + [feats[b][lvl] for b in range(3)], dim=1)
        d = conv_block(P, f"dec_blocks.{blk}", d, n, train, q, wq, first=False, stats_out=stats_out)
    s = "dec_se_final"
    d = q(se_block(d, P[f"{s}.fc1.weight"], P[f"{s}.fc1.bias"], P[f"{s}.fc2.weight"], P[f"{s}.fc2.bias"]))
    return O.conv1x1(d, P["final_conv.weight"], P["final_conv.bias"])


def running_stats_update(P, stats, momentum=O.BN_MOMENTUM):
    """BatchNorm2d running statistics after one training forward (unbiased variance)."""
    out = {}
    for name, (mean, var, n) in stats.items():
        out[f"{name}.running_mean"] = (1 - momentum) * P[f"{name}.running_mean"] + momentum * mean
        out[f"{name}.running_var"] = (1 - momentum) * P[f"{name}.running_var"] + momentum * var * n / (n - 1)
    return out
