"""Generates tests/golden/*.pt from the UNMODIFIED reference (imported from /root/reference) — run in the build
container only (the GPU box has no /root/reference). The reference ships no golden vectors; these
pin oracle/unet_oracle.py and, through it, the CUDA path.

    python oracle/make_golden.py            # writes tests/golden/ops_golden.pt, tests/golden/unet_golden.pt
"""
import hashlib
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def tensor_digest(t):
    """Order-sensitive fingerprint that travels without the tensor: sha1 of the raw bytes + a few moments."""
    t = t.detach().contiguous()
    return {"sha1": hashlib.sha1(t.numpy().tobytes()).hexdigest(), "sum": float(t.double().sum()),
            "abs_sum": float(t.double().abs().sum()), "shape": tuple(t.shape)}


def sample_idx(numel, k=257):
    g = torch.Generator().manual_seed(numel % 9973 + 17)
    return torch.randint(0, numel, (min(k, numel),), generator=g)


def ops_golden():
    g = torch.Generator().manual_seed(7)
    out = {}
    # 3x3 conv (models/model.py:36) forward + autograd backward
    x = torch.randn((2, 5, 8, 12), generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn((7, 5, 3, 3), generator=g, dtype=torch.float64, requires_grad=True)
    b = torch.randn((7,), generator=g, dtype=torch.float64, requires_grad=True)
    z = F.conv2d(x, w, b, padding=1)
    dz = torch.randn(z.shape, generator=g, dtype=torch.float64)
    z.backward(dz)
    out["conv3x3"] = dict(x=x.detach(), w=w.detach(), b=b.detach(), z=z.detach(), dz=dz, dx=x.grad, dw=w.grad, db=b.grad)
    # transposed conv (models/model.py:19)
    x = torch.randn((2, 6, 4, 5), generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn((6, 3, 2, 2), generator=g, dtype=torch.float64, requires_grad=True)
    b = torch.randn((3,), generator=g, dtype=torch.float64, requires_grad=True)
    y = F.conv_transpose2d(x, w, b, stride=2)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    out["convt"] = dict(x=x.detach(), w=w.detach(), b=b.detach(), y=y.detach(), dy=dy, dx=x.grad, dw=w.grad, db=b.grad)
    # BatchNorm2d training (models/model.py:38), incl. running stats
    bn = nn.BatchNorm2d(4).double()
    with torch.no_grad():
        bn.weight.copy_(torch.randn(4, generator=g, dtype=torch.float64))
        bn.bias.copy_(torch.randn(4, generator=g, dtype=torch.float64))
    r = torch.randn((3, 4, 6, 6), generator=g, dtype=torch.float64).clamp_min(0).requires_grad_(True)
    y = bn(r)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    out["bn"] = dict(r=r.detach(), gamma=bn.weight.detach().clone(), beta=bn.bias.detach().clone(), y=y.detach(), dy=dy,
                     dr=r.grad, dgamma=bn.weight.grad, dbeta=bn.bias.grad, running_mean=bn.running_mean.clone(),
                     running_var=bn.running_var.clone(), nbt=bn.num_batches_tracked.clone())
    # max-pool with ties (F.max_pool2d, models/model.py:56): quantised values force many ties
    y = (torch.randint(0, 3, (2, 3, 6, 8), generator=g).double()).requires_grad_(True)
    p = F.max_pool2d(y, 2)
    dp = torch.randn(p.shape, generator=g, dtype=torch.float64)
    p.backward(dp)
    out["pool"] = dict(y=y.detach(), p=p.detach(), dp=dp, dy=y.grad)
    # losses: reference classes (models/loss.py) + torch BCEWithLogits (utils/trainer.py:37), soft targets
    sys.path.insert(0, REF)
    from models.loss import DiceLoss, FocalTverskyLoss
    logits = (torch.randn((3, 1, 16, 16), generator=g, dtype=torch.float64) * 3).requires_grad_(True)
    targets = torch.rand((3, 1, 16, 16), generator=g, dtype=torch.float64)
    targets = torch.where(targets > 0.6, torch.ones_like(targets), targets * 0.3)
    bce = nn.BCEWithLogitsLoss()(logits, targets)
    dice = DiceLoss()(logits, targets)
    ft = FocalTverskyLoss()(logits, targets)
    total = 1.0 * bce + 1.0 * dice + 0.5 * ft
    total.backward()
    out["loss"] = dict(logits=logits.detach(), targets=targets, bce=bce.detach(), dice=dice.detach(), ft=ft.detach(),
                       total=total.detach(), w=(1.0, 1.0, 0.5), dlogits=logits.grad)
    # threshold semantics (utils/trainer.py:217) around zero in fp32
    lg = torch.tensor([-1.0, -1e-7, 0.0, 5e-8, 8.9e-8, 9.0e-8, 1.2e-7, 1e-6, 1.0], dtype=torch.float32)
    out["threshold"] = dict(logits=lg, mask=(torch.sigmoid(lg) > 0.5))
    # AdamW (utils/trainer.py:41): 3 steps, lr 1e-3, default betas/eps/wd
    p = torch.randn(64, generator=g, dtype=torch.float64).requires_grad_(True)
    p0 = p.detach().clone()
    opt = torch.optim.AdamW([p], lr=1e-3)
    grads = []
    for _ in range(3):
        gr = torch.randn(64, generator=g, dtype=torch.float64)
        grads.append(gr)
        p.grad = gr.clone()
        opt.step()
    out["adamw"] = dict(p0=p0, grads=torch.stack(grads), p3=p.detach().clone(), lr=1e-3)
    return out


def unet_golden():
    sys.path.insert(0, REF)
    from models.model import UNet
    from models.loss import DiceLoss
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle.unet_oracle import synth_batch

    out = {}
    torch.manual_seed(42)
    net = UNet()
    net.train()
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    out["init_digest"] = {k: tensor_digest(v) for k, v in sd0.items()}
    out["param_count"] = sum(p.numel() for p in net.parameters() if p.requires_grad)
    out["state_dict_keys"] = list(sd0.keys())
    out["state_dict_shapes"] = {k: tuple(v.shape) for k, v in sd0.items()}

    # case A: structured synthetic batch, B=2 @ 32x32, fp32 CPU, train mode, BCE + Dice
    x, t = synth_batch(2, 32, 32, seed=1234)
    logits = net(x)
    bce = nn.BCEWithLogitsLoss()(logits, t)
    dice = DiceLoss()(logits, t)
    loss = bce + dice
    loss.backward()
    case = dict(x=x, t=t, logits=logits.detach().clone(), bce=float(bce), dice=float(dice), loss=float(loss))
    grads = {}
    for k, p in net.named_parameters():
        g = p.grad.detach()
        idx = sample_idx(g.numel())
        grads[k] = dict(norm=float(g.double().norm()), idx=idx, vals=g.flatten()[idx].clone(),
                        full=g.clone() if g.numel() <= 4096 else None)
    case["grads"] = grads
    sd1 = net.state_dict()
    case["running"] = {k: sd1[k].detach().clone() for k in sd1 if "running" in k or "num_batches" in k}
    # eval-mode logits + mask with the updated running stats
    net.eval()
    with torch.no_grad():
        le = net(x)
    case["eval_logits"] = le.clone()
    case["eval_mask"] = (torch.sigmoid(le) > 0.5)
    out["A"] = case

    # case B: the BASELINE.md probe (known answers BCE 0.746483, Dice 0.620441): B=4 @ 256x256 pure noise
    torch.manual_seed(42)
    net = UNet()
    net.train()
    g = torch.Generator().manual_seed(1234)
    x = torch.rand((4, 1, 256, 256), generator=g)
    t = (torch.rand((4, 1, 256, 256), generator=g) > 0.7).float()
    logits = net(x)
    bce = nn.BCEWithLogitsLoss()(logits, t)
    dice = DiceLoss()(logits, t)
    (bce + dice).backward()
    idx = sample_idx(logits.numel(), 4099)
    caseB = dict(bce=float(bce), dice=float(dice), logits_idx=idx, logits_vals=logits.detach().flatten()[idx].clone(),
                 logits_std=float(logits.std()), grad_norms={k: float(p.grad.double().norm()) for k, p in net.named_parameters()})
    out["B"] = caseB
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    torch.save(ops_golden(), os.path.join(OUT, "ops_golden.pt"))
    torch.save(unet_golden(), os.path.join(OUT, "unet_golden.pt"))
    for f in ("ops_golden.pt", "unet_golden.pt"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
