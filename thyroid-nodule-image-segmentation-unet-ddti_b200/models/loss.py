"""Drop-in replacement for the reference's ``models/loss.py``: DiceLoss, FocalTverskyLoss, BoundaryLoss,
CompositeLoss with the same constructor signatures (reference models/loss.py:9,27,49,69). Dice / FocalTversky / BCE
are evaluated by ONE fused libb2s reduction kernel plus one gradient kernel (CUDA only). BoundaryLoss is the host
scipy distance-transform loss of the reference and is out of scope for the B200 path: it is kept
only so that utils/trainer.py:11,39 imports and constructs it; it is evaluated with stock torch ops.
"""
import torch
import torch.nn as nn

from .. import ops


class _SegLossFunction(torch.autograd.Function):
    @staticmethod
    @ops.on_device_of_input
    def forward(ctx, logits, targets, cfg):
        if not logits.is_cuda:
            raise RuntimeError("b200seg losses run on CUDA (sm_100a) only; there is no CPU fallback")
        lg = logits.detach().float().contiguous()
        tg = targets.detach().float().contiguous()
        B = lg.shape[0]
        per = lg.numel() // B
        dev = lg.device
        partial = torch.empty(B * ops.loss_chunks(per) * 4, dtype=torch.float32, device=dev)
        sums = torch.empty(B * 4, dtype=torch.float32, device=dev)
        out = torch.empty(8, dtype=torch.float32, device=dev)
        ops.seg_loss_fwd(lg, tg, partial, sums, out, **cfg)
        ctx.save_for_backward(lg, tg, sums, out)
        ctx.cfg = cfg
        ctx.in_dtype = logits.dtype
        return out[0].clone()

    @staticmethod
    @ops.on_device_of_input
    def backward(ctx, grad_out):
        lg, tg, sums, out = ctx.saved_tensors
        dl = torch.empty_like(lg)
        go = grad_out.detach().float().contiguous().view(1)
        ops.seg_loss_bwd(lg, tg, sums, out[4:7], go, dl, **ctx.cfg)
        return dl.to(ctx.in_dtype), None, None


def _cfg(w_bce=0.0, w_dice=0.0, w_ft=0.0, dice_smooth=1.0, ft_alpha=0.4, ft_beta=0.6, ft_gamma=2.0, ft_smooth=1e-6):
    return dict(w_bce=float(w_bce), w_dice=float(w_dice), w_ft=float(w_ft), dice_smooth=float(dice_smooth),
                ft_alpha=float(ft_alpha), ft_beta=float(ft_beta), ft_gamma=float(ft_gamma), ft_smooth=float(ft_smooth))


class DiceLoss(nn.Module):
    """Soft Dice on sigmoid(logits), per-sample, smooth=1 (reference models/loss.py:7-24)."""

    def __init__(self, smooth=1.0):
        super().__init__()
        self.smooth = smooth

    def forward(self, logits, targets):
        return _SegLossFunction.apply(logits, targets, _cfg(w_dice=1.0, dice_smooth=self.smooth))


class FocalTverskyLoss(nn.Module):
    """Batch-global focal Tversky (reference models/loss.py:26-46)."""

    def __init__(self, alpha=0.4, beta=0.6, gamma=2.0, smooth=1e-6):
        super().__init__()
        self.alpha, self.beta, self.gamma, self.smooth = alpha, beta, gamma, smooth

    def forward(self, logits, targets):
        return _SegLossFunction.apply(logits, targets, _cfg(w_ft=1.0, ft_alpha=self.alpha, ft_beta=self.beta,
                                                            ft_gamma=self.gamma, ft_smooth=self.smooth))


class BCEDiceLoss(nn.Module):
    """w_bce * BCEWithLogits + w_dice * Dice in one pass — the Dice+BCE objective of BASELINE.json
    (utils/trainer.py:85-86,90 with bce_ratio = dice_ratio = 1)."""

    def __init__(self, w_bce=1.0, w_dice=1.0, smooth=1.0):
        super().__init__()
        self.w_bce, self.w_dice, self.smooth = w_bce, w_dice, smooth

    def forward(self, logits, targets):
        return _SegLossFunction.apply(logits, targets, _cfg(w_bce=self.w_bce, w_dice=self.w_dice,
                                                            dice_smooth=self.smooth))


class BoundaryLoss(nn.Module):
    """Host-side scipy EDT boundary loss (reference models/loss.py:48-66). OUT OF SCOPE for the CUDA path; stock ops."""

    def forward(self, logits, targets):
        import numpy as np
        import scipy.ndimage as nd
        probs = torch.sigmoid(logits)
        t_np = targets.detach().cpu().numpy().astype(np.uint8)
        loss = 0.0
        for b in range(targets.shape[0]):
            dist = torch.from_numpy(nd.distance_transform_edt(1 - t_np[b, 0])).to(logits.device).float()
            loss = loss + torch.mean(torch.abs(probs[b, 0] - targets[b, 0]) * dist)
        return loss / targets.shape[0]


class CompositeLoss(nn.Module):
    """Weighted sum (reference models/loss.py:68-83); the λ names are the reference's literal keyword names.
    The FocalTversky (alpha .3, beta .7, gamma .75), BCE and Dice terms share one fused kernel pass."""

    def __init__(self, λ_ft=1.0, λ_b=0.5, λ_bce=0.0, λ_dice=0.0):
        super().__init__()
        self.bl = BoundaryLoss()
        self.λ_ft, self.λ_b, self.λ_bce, self.λ_dice = λ_ft, λ_b, λ_bce, λ_dice

    def forward(self, logits, targets):
        cfg = _cfg(w_bce=self.λ_bce if self.λ_bce > 0 else 0.0, w_dice=self.λ_dice if self.λ_dice > 0 else 0.0,
                   w_ft=self.λ_ft, ft_alpha=0.3, ft_beta=0.7, ft_gamma=0.75)
        loss = _SegLossFunction.apply(logits, targets, cfg)
        if self.λ_b != 0:
            loss = loss + self.λ_b * self.bl(logits, targets)
        return loss
