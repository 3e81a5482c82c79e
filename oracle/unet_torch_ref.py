"""ORACLE (test infrastructure / CPU baseline, not product code) — functional restatement of the reference's UNet
forward graph with the SAME torch library ops the reference's nn.Modules dispatch to (F.conv2d, F.relu,
F.batch_norm, F.max_pool2d, F.conv_transpose2d, torch.cat) and torch autograd for the backward pass.

Used (a) as the timed CPU baseline of bench.py (`cpu_baseline`, `--impl reference`): it executes exactly the
oneDNN/ATen CPU kernels the reference would, on all host cores; (b) as a second pin of unet_oracle.py.
Graph: reference models/model.py:53-73; block order conv -> ReLU -> BatchNorm: models/model.py:33-43;
losses: models/loss.py:13-24 and nn.BCEWithLogitsLoss (utils/trainer.py:37).
"""
import torch
import torch.nn.functional as F


def _block(P, name, x, train):
    for idx in (0, 3):
        x = F.conv2d(x, P[f"{name}.{idx}.weight"], P[f"{name}.{idx}.bias"], padding=1)
        x = F.relu(x)
        bn = f"{name}.{idx + 2}"
        x = F.batch_norm(x, P[f"{bn}.running_mean"], P[f"{bn}.running_var"], P[f"{bn}.weight"], P[f"{bn}.bias"],
                         training=train, momentum=0.1, eps=1e-5)
    return x


def unet_forward(P, x, train=True):
    e1 = _block(P, "encoder1", x, train)
    e2 = _block(P, "encoder2", F.max_pool2d(e1, 2), train)
    e3 = _block(P, "encoder3", F.max_pool2d(e2, 2), train)
    e4 = _block(P, "encoder4", F.max_pool2d(e3, 2), train)
    d = _block(P, "middle.1", F.max_pool2d(e4, 2), train)
    d = F.conv_transpose2d(d, P["middle.2.weight"], P["middle.2.bias"], stride=2)
    for name, ct, skip in (("decoder3.0", "decoder3.1", e4), ("decoder2.0", "decoder2.1", e3),
                           ("decoder1.0", "decoder1.1", e2)):
        d = _block(P, name, torch.cat([d, skip], dim=1), train)
        d = F.conv_transpose2d(d, P[f"{ct}.weight"], P[f"{ct}.bias"], stride=2)
    d = _block(P, "final.0", torch.cat([d, e1], dim=1), train)
    return F.conv2d(d, P["final.1.weight"], P["final.1.bias"])


def dice_loss(logits, targets, smooth=1.0):
    p = torch.sigmoid(logits).reshape(logits.shape[0], -1)
    t = targets.reshape(targets.shape[0], -1).float()
    inter = (p * t).sum(dim=1)
    union = p.sum(dim=1) + t.sum(dim=1)
    return 1 - ((2.0 * inter + smooth) / (union + smooth)).mean()


def train_step(P, x, t):
    """Forward + BCE + Dice + backward (BASELINE config 1). P: leaf tensors with requires_grad on parameters."""
    logits = unet_forward(P, x, train=True)
    loss = F.binary_cross_entropy_with_logits(logits, t) + dice_loss(logits, t)
    params = [v for v in P.values() if v.requires_grad]
    grads = torch.autograd.grad(loss, params)
    return loss.detach(), logits.detach(), grads


def make_params(state_dict, requires_grad=True):
    P = {}
    for k, v in state_dict.items():
        v = v.detach().clone()
        if requires_grad and v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
        P[k] = v
    return P
