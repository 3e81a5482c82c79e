"""Generates tests/golden/vnet_golden.pt from the UNMODIFIED reference models/vnet.py (imported from /root/reference)
— run in the build container only. The V-Net has 160 M parameters, so the file holds digests and samples, not weights:
the tests rebuild the weights from the same seed through the drop-in module (bit-identical init, checked by digest).

    python oracle/make_golden_vnet.py
"""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle.make_golden import REF, OUT, tensor_digest, sample_idx  # noqa: E402


def main():
    sys.path.insert(0, REF)
    from models.vnet import ImprovedVNet, SEBlock, ConvBlock
    from models.loss import DiceLoss
    from oracle.unet_oracle import synth_batch
    torch.set_num_threads(os.cpu_count() or 1)
    out = {}
    g = torch.Generator().manual_seed(11)

    # op level: stride-2 conv (models/vnet.py:97), SEBlock, ConvBlock with and without projection (dropout 0)
    x = torch.randn((2, 6, 8, 12), generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn((10, 6, 3, 3), generator=g, dtype=torch.float64, requires_grad=True)
    b = torch.randn((10,), generator=g, dtype=torch.float64, requires_grad=True)
    z = F.conv2d(x, w, b, stride=2, padding=1)
    dz = torch.randn(z.shape, generator=g, dtype=torch.float64)
    z.backward(dz)
    out["conv_s2"] = dict(x=x.detach(), w=w.detach(), b=b.detach(), z=z.detach(), dz=dz, dx=x.grad, dw=w.grad, db=b.grad)

    torch.manual_seed(5)
    se = SEBlock(16, reduction=4).double()
    x = torch.randn((3, 16, 6, 5), generator=g, dtype=torch.float64, requires_grad=True)
    y = se(x)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    out["se"] = dict(x=x.detach(), y=y.detach(), dy=dy, dx=x.grad,
                     params={k: v.detach().clone() for k, v in se.state_dict().items()},
                     grads={k: p.grad.clone() for k, p in se.named_parameters()})

    for name, cin, cout, n in (("block_proj", 6, 8, 2), ("block_id", 8, 8, 3)):
        torch.manual_seed(6)
        blk = ConvBlock(cin, cout, n, 0.0).double().train()
        x = torch.randn((2, cin, 6, 6), generator=g, dtype=torch.float64, requires_grad=True)
        y = blk(x)
        dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
        y.backward(dy)
        out[name] = dict(x=x.detach(), y=y.detach(), dy=dy, dx=x.grad, num_convs=n,
                         params={k: v.detach().clone() for k, v in blk.state_dict().items()},
                         grads={k: p.grad.clone() for k, p in blk.named_parameters()})

    # whole net: defaults except dropout_rate = 0, seed 42, B = 2 @ 32x32, fp32 CPU, train mode
    torch.manual_seed(42)
    net = ImprovedVNet(dropout_rate=0.0)
    net.train()
    sd0 = net.state_dict()
    out["param_count"] = sum(p.numel() for p in net.parameters() if p.requires_grad)
    out["state_dict_keys"] = list(sd0.keys())
    out["state_dict_shapes"] = {k: tuple(v.shape) for k, v in sd0.items()}
    out["init_digest"] = {k: dict(sum=float(v.double().sum()), abs_sum=float(v.double().abs().sum()))
                          for k, v in sd0.items() if v.is_floating_point()}
    x, t = synth_batch(2, 32, 32, seed=1234)
    logits = net(x)
    bce = nn.BCEWithLogitsLoss()(logits, t)
    dice = DiceLoss()(logits, t)
    loss = bce + dice
    loss.backward()
    case = dict(x=x, t=t, logits=logits.detach().clone(), bce=float(bce), dice=float(dice), loss=float(loss))
    grads = {}
    for k, p in net.named_parameters():
        gr = p.grad.detach()
        idx = sample_idx(gr.numel())
        grads[k] = dict(norm=float(gr.double().norm()), idx=idx, vals=gr.flatten()[idx].clone())
    case["grads"] = grads
    sd1 = net.state_dict()
    case["running_digest"] = {k: dict(sum=float(sd1[k].double().sum()), abs_sum=float(sd1[k].double().abs().sum()))
                              for k in sd1 if "running" in k}
    net.eval()
    with torch.no_grad():
        le = net(x)
    case["eval_logits"] = le.clone()
    case["eval_mask"] = (torch.sigmoid(le) > 0.5)
    out["A"] = case
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "vnet_golden.pt")
    torch.save(out, path)
    print("vnet_golden.pt", os.path.getsize(path), "bytes; params", out["param_count"], "loss", case["loss"])


if __name__ == "__main__":
    main()
