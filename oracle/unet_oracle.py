"""ORACLE (test infrastructure, not product code) — CPU restatement of the reference's UNet hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

The reference's arithmetic lives in a third-party dependency that is not vendored under /root/reference:
PyTorch (unpinned upstream; this project pins torch 2.11.0+cu128, the container's). This file restates the
published semantics of the torch ops at the reference's call sites with elementary tensor algebra (pad, slice,
matmul/einsum, sums) and explicit hand-derived backward formulas — it does not call F.conv2d, F.batch_norm,
F.max_pool2d, nn.BCEWithLogitsLoss or autograd. It is pinned against the real reference modules
(models/model.py UNet, models/loss.py DiceLoss/FocalTverskyLoss, nn.BCEWithLogitsLoss) by
oracle/make_golden.py, which imports /root/reference in the build container and commits the resulting vectors to
tests/golden/ (the reference itself ships no golden vectors: SURVEY.md §8c).

Call sites restated (reference file:line):
  conv3x3 / conv1x1      models/model.py:36,39,30        (nn.Conv2d, padding=1 / kernel 1)
  relu                   models/model.py:37,40           (nn.ReLU)
  batchnorm              models/model.py:38,41           (nn.BatchNorm2d: eps 1e-5, momentum 0.1, biased batch var)
  maxpool2x2             models/model.py:17,56-58        (first maximum in row-major window order gets the gradient)
  conv_transpose2x2      models/model.py:19,49           (nn.ConvTranspose2d k=2 s=2)
  concat order           models/model.py:64,66,68,70     ([decoder, encoder])
  unet forward graph     models/model.py:53-73
  dice                   models/loss.py:13-24
  focal tversky          models/loss.py:34-46
  bce with logits        utils/trainer.py:37,85          (mean over all elements)
  threshold              utils/trainer.py:101,152,217    (sigmoid(logits) > 0.5 in the logits dtype)

Every function is device-agnostic elementary tensor algebra: the GPU tests run the SAME restatement in fp64 on the
device for the BASELINE-sized cases (4 x 256^2), where the CPU would need minutes.

`q` (quantiser) emulates the CUDA path's storage precision: q = bf16_round reproduces every point where the
B200 kernels round an activation / gradient / weight operand to bf16; q = identity is the exact-arithmetic oracle.
"""
import math

import torch

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def identity(t):
    return t


def bf16_round(t):
    return t.to(torch.bfloat16).to(t.dtype)


# ------------------------------------------------------------------------------------------------------------
# elementary ops (NCHW), forward and explicit backward
# ------------------------------------------------------------------------------------------------------------
def _shift(x, dh, dw):
    """out[n,c,h,w] = x[n,c,h+dh,w+dw] with zero padding."""
    N, C, H, W = x.shape
    out = torch.zeros_like(x)
    hs, he = max(0, -dh), min(H, H - dh)
    ws, we = max(0, -dw), min(W, W - dw)
    if hs < he and ws < we:
        out[:, :, hs:he, ws:we] = x[:, :, hs + dh:he + dh, ws + dw:we + dw]
    return out


def conv3x3(x, w, b=None):
    """Cross-correlation, padding 1 (models/model.py:36): sum over the 9 taps of a channel contraction."""
    out = None
    for t in range(9):
        dh, dw = t // 3 - 1, t % 3 - 1
        term = torch.einsum("nchw,oc->nohw", _shift(x, dh, dw), w[:, :, t // 3, t % 3])
        out = term if out is None else out + term
    if b is not None:
        out = out + b.view(1, -1, 1, 1)
    return out


def conv3x3_bwd(x, w, dz):
    """Returns (dx, dw, db) for z = conv3x3(x, w, b)."""
    dx = torch.zeros_like(x)
    dw = torch.zeros_like(w)
    for t in range(9):
        dh, dw_ = t // 3 - 1, t % 3 - 1
        # z[h,w] += W_t x[h+dh, w+dw]  =>  dx[h+dh, w+dw] += W_t^T dz[h,w] ; dW_t += dz (x) x shifted
        dx = dx + _shift(torch.einsum("nohw,oc->nchw", dz, w[:, :, t // 3, t % 3]), -dh, -dw_)
        dw[:, :, t // 3, t % 3] = torch.einsum("nohw,nchw->oc", dz, _shift(x, dh, dw_))
    db = dz.sum(dim=(0, 2, 3))
    return dx, dw, db


def conv1x1(x, w, b=None):
    out = torch.einsum("nchw,oc->nohw", x, w[:, :, 0, 0])
    if b is not None:
        out = out + b.view(1, -1, 1, 1)
    return out


def conv1x1_bwd(x, w, dz):
    dx = torch.einsum("nohw,oc->nchw", dz, w[:, :, 0, 0])
    dw = torch.einsum("nohw,nchw->oc", dz, x).view_as(w)
    return dx, dw, dz.sum(dim=(0, 2, 3))


def conv_transpose2x2(x, w, b=None):
    """nn.ConvTranspose2d(k=2,s=2) (models/model.py:19): out[n,o,2i+a,2j+b] = sum_c x[n,c,i,j] w[c,o,a,b] + bias."""
    N, C, H, W = x.shape
    O = w.shape[1]
    out = torch.zeros((N, O, 2 * H, 2 * W), dtype=x.dtype, device=x.device)
    for a in range(2):
        for bb in range(2):
            out[:, :, a::2, bb::2] = torch.einsum("nchw,co->nohw", x, w[:, :, a, bb])
    if b is not None:
        out = out + b.view(1, -1, 1, 1)
    return out


def conv_transpose2x2_bwd(x, w, dy):
    dx = torch.zeros_like(x)
    dw = torch.zeros_like(w)
    for a in range(2):
        for bb in range(2):
            g = dy[:, :, a::2, bb::2]
            dx = dx + torch.einsum("nohw,co->nchw", g, w[:, :, a, bb])
            dw[:, :, a, bb] = torch.einsum("nchw,nohw->co", x, g)
    return dx, dw, dy.sum(dim=(0, 2, 3))


def batchnorm_stats(r):
    """Batch mean and BIASED variance over (N,H,W) (models/model.py:38 in train mode)."""
    mean = r.mean(dim=(0, 2, 3))
    var = ((r - mean.view(1, -1, 1, 1)) ** 2).mean(dim=(0, 2, 3))
    return mean, var


def batchnorm_affine(mean, var, gamma, beta, eps=BN_EPS):
    invstd = 1.0 / torch.sqrt(var + eps)
    scale = gamma * invstd
    shift = beta - mean * scale
    return scale, shift, invstd


def batchnorm_bwd(dy, r, mean, invstd, gamma):
    """Train-mode BatchNorm backward: returns (dr, dgamma, dbeta)."""
    n = r.shape[0] * r.shape[2] * r.shape[3]
    xhat = (r - mean.view(1, -1, 1, 1)) * invstd.view(1, -1, 1, 1)
    dbeta = dy.sum(dim=(0, 2, 3))
    dgamma = (dy * xhat).sum(dim=(0, 2, 3))
    dr = (gamma * invstd).view(1, -1, 1, 1) * (dy - (dbeta / n).view(1, -1, 1, 1) - xhat * (dgamma / n).view(1, -1, 1, 1))
    return dr, dgamma, dbeta


def maxpool2x2(y):
    """F.max_pool2d(y, 2) (models/model.py:56): returns (pooled, argmax in {0,1,2,3} row-major, first max wins)."""
    N, C, H, W = y.shape
    win = torch.stack([y[:, :, 0::2, 0::2], y[:, :, 0::2, 1::2], y[:, :, 1::2, 0::2], y[:, :, 1::2, 1::2]], dim=-1)
    best = win[..., 0].clone()
    arg = torch.zeros_like(best, dtype=torch.long)
    for qi in range(1, 4):
        better = win[..., qi] > best
        best = torch.where(better, win[..., qi], best)
        arg = torch.where(better, torch.full_like(arg, qi), arg)
    return best, arg


def maxpool2x2_bwd(dpool, arg, shape):
    dy = torch.zeros(shape, dtype=dpool.dtype, device=dpool.device)
    for qi in range(4):
        dy[:, :, (qi >> 1)::2, (qi & 1)::2] = torch.where(arg == qi, dpool, torch.zeros_like(dpool))
    return dy


def sigmoid(x):
    e = torch.exp(-x.abs())
    return torch.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e))


def threshold_mask(logits):
    """(torch.sigmoid(logits) > 0.5) evaluated in the logits dtype (utils/trainer.py:217)."""
    return (1.0 / (1.0 + torch.exp(-logits))) > 0.5


# ------------------------------------------------------------------------------------------------------------
# losses (models/loss.py:13-24, 34-46; nn.BCEWithLogitsLoss)
# ------------------------------------------------------------------------------------------------------------
def seg_loss(logits, targets, w_bce=1.0, w_dice=1.0, w_ft=0.0, dice_smooth=1.0, ft_alpha=0.4, ft_beta=0.6,
             ft_gamma=2.0, ft_smooth=1e-6):
    """Returns dict(total, bce, dice, ft, dlogits) with the analytic gradient of `total`."""
    B = logits.shape[0]
    x = logits.reshape(B, -1)
    t = targets.reshape(B, -1).to(x.dtype)
    n = x.numel()
    p = sigmoid(x)
    bce_el = torch.clamp(x, min=0) - x * t + torch.log1p(torch.exp(-x.abs()))
    bce = bce_el.sum() / n
    td = t.float().to(x.dtype)           # models/loss.py:19 casts the Dice targets with .float()
    I = (p * td).sum(dim=1)
    tsum = t.float().sum(dim=1).to(x.dtype)   # ... and therefore sums them in fp32 (models/loss.py:22)
    U = p.sum(dim=1) + tsum
    dice_b = (2.0 * I + dice_smooth) / (U + dice_smooth)
    dice = 1.0 - dice_b.mean()
    TP = (p * t).sum()
    FP = (p * (1 - t)).sum()
    FN = ((1 - p) * t).sum()
    D = TP + ft_alpha * FP + ft_beta * FN + ft_smooth
    ti = (TP + ft_smooth) / D
    ft = (1 - ti) ** ft_gamma
    total = w_bce * bce + w_dice * dice + w_ft * ft
    # gradients
    dp = p * (1 - p)
    g = w_bce * (p - t) / n
    den = (U + dice_smooth).unsqueeze(1)
    g = g + w_dice * (-(1.0 / B)) * (2.0 * td * den - (2.0 * I + dice_smooth).unsqueeze(1)) / (den * den) * dp
    if w_ft != 0.0:
        dti = (t * D - (TP + ft_smooth) * (t + ft_alpha * (1 - t) - ft_beta * t)) / (D * D)
        dL_dti = -ft_gamma * (1 - ti) ** (ft_gamma - 1.0) if float(1 - ti) > 0 else torch.zeros((), dtype=x.dtype, device=x.device)
        g = g + w_ft * dL_dti * dti * dp
    return {"total": total, "bce": bce, "dice": dice, "ft": ft, "dlogits": g.reshape(logits.shape),
            "sums": torch.stack([I, p.sum(dim=1), tsum, bce_el.sum(dim=1)], dim=1)}


# ------------------------------------------------------------------------------------------------------------
# UNet (models/model.py:5-73): parameter naming follows the reference state_dict
# ------------------------------------------------------------------------------------------------------------
BLOCKS = ["encoder1", "encoder2", "encoder3", "encoder4", "middle.1", "decoder3.0", "decoder2.0", "decoder1.0",
          "final.0"]
CONVT = {"middle.1": "middle.2", "decoder3.0": "decoder3.1", "decoder2.0": "decoder2.1", "decoder1.0": "decoder1.1"}


def _conv_bn(P, name, idx, x, cache, train, q, in_is_image):
    """conv(+bias) -> ReLU -> BatchNorm, one of the two stages of conv_block (models/model.py:33-43)."""
    w = P[f"{name}.{idx}.weight"]
    b = P[f"{name}.{idx}.bias"]
    wq = w if in_is_image else q(w)          # the Cin=1 first conv runs on fp32 weights and fp32 image
    r = q(torch.clamp(conv3x3(x, wq, b), min=0))
    bn = f"{name}.{idx + 2}"
    gamma, beta = P[f"{bn}.weight"], P[f"{bn}.bias"]
    if train:
        mean, var = batchnorm_stats(r)
    else:
        mean, var = P[f"{bn}.running_mean"], P[f"{bn}.running_var"]
    scale, shift, invstd = batchnorm_affine(mean, var, gamma, beta)
    if cache is not None:
        cache[f"{name}.{idx}"] = dict(x=x, r=r, mean=mean, var=var, invstd=invstd, scale=scale, shift=shift)
    return r, scale, shift


def _bn_apply(r, scale, shift, q):
    return q(r * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))


def unet_forward(P, x, train=True, q=identity, cache=None):
    """Reference forward graph (models/model.py:53-73). P: dict name -> tensor (reference state_dict keys).
    Returns logits [N,O,H,W]. If `cache` is a dict it is filled for unet_backward."""
    if x.shape[2] % 16 or x.shape[3] % 16:
        raise RuntimeError("Sizes of tensors must match: H and W must be multiples of 16 (models/model.py:64)")

    def block(name, inp, first=False):
        r0, s0, h0 = _conv_bn(P, name, 0, inp, cache, train, q, first)
        y0 = _bn_apply(r0, s0, h0, q)
        r1, s1, h1 = _conv_bn(P, name, 3, y0, cache, train, q, False)
        return r1, s1, h1

    skips = []
    cur = x
    for i, name in enumerate(["encoder1", "encoder2", "encoder3", "encoder4"]):
        r1, s1, h1 = block(name, cur, first=(i == 0))
        y = _bn_apply(r1, s1, h1, q)
        pooled, arg = maxpool2x2(y)
        if cache is not None:
            cache[f"{name}.pool"] = dict(arg=arg, shape=y.shape)
        skips.append(y)
        cur = pooled
    # middle: (pool already applied) conv_block(512,1024) -> convT
    r1, s1, h1 = block("middle.1", cur)
    y = _bn_apply(r1, s1, h1, q)
    up_in = {"middle.1": y}
    d = q(conv_transpose2x2(y, q(P["middle.2.weight"]), P["middle.2.bias"]))
    if cache is not None:
        cache["middle.2"] = dict(x=y)
    for name, skip in zip(["decoder3.0", "decoder2.0", "decoder1.0"], [skips[3], skips[2], skips[1]]):
        cat = torch.cat([d, skip], dim=1)
        r1, s1, h1 = block(name, cat)
        y = _bn_apply(r1, s1, h1, q)
        ct = CONVT[name]
        d = q(conv_transpose2x2(y, q(P[f"{ct}.weight"]), P[f"{ct}.bias"]))
        if cache is not None:
            cache[ct] = dict(x=y)
    cat = torch.cat([d, skips[0]], dim=1)
    r1, s1, h1 = block("final.0", cat)
    # head: BN apply folded into the 1x1 conv in fp32 (the CUDA path never materialises this BN output)
    wf = P["final.1.weight"][:, :, 0, 0] * s1.view(1, -1)
    bf = P["final.1.bias"] + (P["final.1.weight"][:, :, 0, 0] * h1.view(1, -1)).sum(dim=1)
    logits = torch.einsum("nchw,oc->nohw", r1, wf) + bf.view(1, -1, 1, 1)
    if cache is not None:
        cache["head"] = dict(r=r1, scale=s1, shift=h1)
    del up_in
    return logits


def unet_backward(P, cache, dlogits, q=identity):
    """Explicit reverse pass of unet_forward (train mode). Returns dict name -> gradient for every parameter."""
    G = {}

    def conv_bn_bwd(name, idx, dy, first=False, need_dx=True):
        c = cache[f"{name}.{idx}"]
        bn = f"{name}.{idx + 2}"
        dr, dgamma, dbeta = batchnorm_bwd(dy, c["r"], c["mean"], c["invstd"], P[f"{bn}.weight"])
        G[f"{bn}.weight"], G[f"{bn}.bias"] = dgamma, dbeta
        dz = q(torch.where(c["r"] > 0, dr, torch.zeros_like(dr)))
        w = P[f"{name}.{idx}.weight"]
        wq = w if first else q(w)
        dx, dw, db = conv3x3_bwd(c["x"], wq, dz)
        G[f"{name}.{idx}.weight"], G[f"{name}.{idx}.bias"] = dw, db
        return q(dx) if need_dx else None

    def block_bwd(name, dy1, first=False):
        dy0 = conv_bn_bwd(name, 3, dy1)
        return conv_bn_bwd(name, 0, dy0, first=first, need_dx=not first)

    # head
    h = cache["head"]
    w1 = P["final.1.weight"][:, :, 0, 0]
    y_last = _bn_apply(h["r"], h["scale"], h["shift"], q)
    G["final.1.weight"] = torch.einsum("nohw,nchw->oc", dlogits, y_last).view_as(P["final.1.weight"])
    G["final.1.bias"] = dlogits.sum(dim=(0, 2, 3))
    dy = q(torch.einsum("nohw,oc->nchw", dlogits, w1))
    dcat = block_bwd("final.0", dy)
    skip_grads = {}
    for name, enc in zip(["decoder1.0", "decoder2.0", "decoder3.0"], ["encoder1", "encoder2", "encoder3"]):
        ct = CONVT[name]
        half = dcat.shape[1] // 2
        dd, skip_grads[enc] = dcat[:, :half], dcat[:, half:]
        dxt, dwt, dbt = conv_transpose2x2_bwd(cache[ct]["x"], q(P[f"{ct}.weight"]), dd)
        G[f"{ct}.weight"], G[f"{ct}.bias"] = dwt, dbt
        dcat = block_bwd(name, q(dxt))
    half = dcat.shape[1] // 2
    dd, skip_grads["encoder4"] = dcat[:, :half], dcat[:, half:]
    dxt, dwt, dbt = conv_transpose2x2_bwd(cache["middle.2"]["x"], q(P["middle.2.weight"]), dd)
    G["middle.2.weight"], G["middle.2.bias"] = dwt, dbt
    dpool = block_bwd("middle.1", q(dxt))
    for i, enc in enumerate(["encoder4", "encoder3", "encoder2", "encoder1"]):
        pc = cache[f"{enc}.pool"]
        dy = skip_grads[enc] + maxpool2x2_bwd(dpool, pc["arg"], pc["shape"])
        dpool = block_bwd(enc, dy, first=(enc == "encoder1"))
    return G


def running_stats_update(P, cache, momentum=BN_MOMENTUM):
    """In-place BatchNorm running-stat update of a train-mode forward (models/model.py:38; unbiased variance)."""
    for key, c in cache.items():
        if "mean" not in c:
            continue
        name, idx = key.rsplit(".", 1)
        bn = f"{name}.{int(idx) + 2}"
        n = c["r"].shape[0] * c["r"].shape[2] * c["r"].shape[3]
        P[f"{bn}.running_mean"].mul_(1 - momentum).add_(momentum * c["mean"])
        P[f"{bn}.running_var"].mul_(1 - momentum).add_(momentum * c["var"] * n / max(n - 1, 1))
        P[f"{bn}.num_batches_tracked"] += 1


def adamw_step(p, g, m, v, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01, step=1):
    """torch.optim.AdamW single-tensor update (utils/trainer.py:41), in place."""
    p.mul_(1 - lr * weight_decay)
    m.mul_(beta1).add_((1 - beta1) * g)
    v.mul_(beta2).add_((1 - beta2) * g * g)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p.sub_((lr / bc1) * m / denom)


# ------------------------------------------------------------------------------------------------------------
# synthetic "ultrasound-shaped" data and reference-default initialisation
# ------------------------------------------------------------------------------------------------------------
def synth_batch(B, H, W, seed=1234, device="cpu"):
    """Low-frequency tissue field x (1 - 0.6 nodule) x Rayleigh speckle, clamped to [0,1]; mask = rotated ellipse."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand((B, 1, 8, 8), generator=g)
    tissue = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False) * 0.5 + 0.25
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    cx = (torch.rand(B, generator=g) * 0.8 - 0.4).view(B, 1, 1)
    cy = (torch.rand(B, generator=g) * 0.8 - 0.4).view(B, 1, 1)
    ax = (torch.rand(B, generator=g) * 0.35 + 0.15).view(B, 1, 1)
    ay = (torch.rand(B, generator=g) * 0.35 + 0.15).view(B, 1, 1)
    th = (torch.rand(B, generator=g) * math.pi).view(B, 1, 1)
    xr = (xx - cx) * torch.cos(th) + (yy - cy) * torch.sin(th)
    yr = -(xx - cx) * torch.sin(th) + (yy - cy) * torch.cos(th)
    nodule = (((xr / ax) ** 2 + (yr / ay) ** 2) <= 1.0).float().unsqueeze(1)
    u = torch.rand((B, 1, H, W), generator=g).clamp_min(1e-6)
    speckle = torch.sqrt(-2.0 * torch.log(u)) / 1.2533
    img = (tissue * (1.0 - 0.6 * nodule) * speckle).clamp(0.0, 1.0)
    return img.to(device), nodule.to(device)


def unet_param_shapes(in_channels=1, out_channels=1):
    """Reference state_dict layout: name -> shape, in registration order."""
    shapes = {}

    def conv_block(prefix, cin, cout):
        for idx, ci in ((0, cin), (3, cout)):
            shapes[f"{prefix}.{idx}.weight"] = (cout, ci, 3, 3)
            shapes[f"{prefix}.{idx}.bias"] = (cout,)
            bn = f"{prefix}.{idx + 2}"
            shapes[f"{bn}.weight"] = (cout,)
            shapes[f"{bn}.bias"] = (cout,)
            shapes[f"{bn}.running_mean"] = (cout,)
            shapes[f"{bn}.running_var"] = (cout,)
            shapes[f"{bn}.num_batches_tracked"] = ()

    conv_block("encoder1", in_channels, 64)
    conv_block("encoder2", 64, 128)
    conv_block("encoder3", 128, 256)
    conv_block("encoder4", 256, 512)
    conv_block("middle.1", 512, 1024)
    shapes["middle.2.weight"] = (1024, 512, 2, 2)
    shapes["middle.2.bias"] = (512,)
    for name, cin, cout in (("decoder3", 1024, 256), ("decoder2", 512, 128), ("decoder1", 256, 64)):
        conv_block(f"{name}.0", cin, cin // 2)
        shapes[f"{name}.1.weight"] = (cin // 2, cout, 2, 2)
        shapes[f"{name}.1.bias"] = (cout,)
    conv_block("final.0", 128, 64)
    shapes["final.1.weight"] = (out_channels, 64, 1, 1)
    shapes["final.1.bias"] = (out_channels,)
    return shapes


def unet_init(seed=42, in_channels=1, out_channels=1, dtype=torch.float32):
    """torch default init restated: conv/convT weight and bias ~ U(+-1/sqrt(fan_in)) (kaiming
    uniform with a=sqrt(5) reduces to that bound), BN gamma=1 beta=0 rm=0 rv=1. Not bit-identical to
    torch.manual_seed(seed); UNet() — goldens carry the reference's own weights."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, shp in unet_param_shapes(in_channels, out_channels).items():
        if name.endswith("num_batches_tracked"):
            P[name] = torch.zeros((), dtype=torch.long)
        elif name.endswith("running_mean"):
            P[name] = torch.zeros(shp, dtype=dtype)
        elif name.endswith("running_var"):
            P[name] = torch.ones(shp, dtype=dtype)
        elif len(shp) == 4:
            # ConvTranspose2d weight is [Cin, Cout, k, k]; torch computes fan_in from dim 1 for both layouts
            fan_in = shp[1] * shp[2] * shp[3]
            bound = 1.0 / math.sqrt(fan_in)
            P[name] = ((torch.rand(shp, generator=g) * 2 - 1) * bound).to(dtype)
            P["__last_bound__"] = bound
        elif name.rsplit(".", 1)[0] + ".running_mean" in unet_param_shapes(in_channels, out_channels):
            P[name] = torch.ones(shp, dtype=dtype) if name.endswith("weight") else torch.zeros(shp, dtype=dtype)
        else:  # conv / convT bias
            bound = P["__last_bound__"]
            P[name] = ((torch.rand(shp, generator=g) * 2 - 1) * bound).to(dtype)
    P.pop("__last_bound__", None)
    return P
