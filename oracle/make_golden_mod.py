"""Generates tests/golden/mod_golden.pt from the UNMODIFIED reference models/mod.py (UNet and ResUNet, depth 3 so the
weights stay small enough to rebuild from the seed) — run in the build container only."""
import os
import sys

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle.make_golden import REF, OUT, sample_idx  # noqa: E402


def main():
    sys.path.insert(0, REF)
    from models import mod as R
    from models.loss import DiceLoss
    from oracle.unet_oracle import synth_batch
    torch.set_num_threads(os.cpu_count() or 1)
    out = {"param_count_default": {"ResUNet": sum(p.numel() for p in R.ResUNet().parameters()),
                                   "UNet": sum(p.numel() for p in R.UNet().parameters())}}
    out["param_count_default"]["AttentionUNet"] = sum(p.numel() for p in R.AttentionUNet().parameters())
    x, t = synth_batch(2, 32, 32, seed=1234)
    xo, to = synth_batch(2, 36, 44, seed=1235)          # 36 -> 18 -> 9 -> 4: the up-sampled 8 != 9 takes the bilinear branch
    g3 = torch.Generator().manual_seed(9)
    x3 = torch.cat([x, torch.rand((2, 2, 32, 32), generator=g3)], dim=1)      # a 3-channel image
    cases = (("ResUNet", R.ResUNet, {}, x, t), ("UNet", R.UNet, {}, x, t), ("AttentionUNet", R.AttentionUNet, {}, x, t),
             ("UNet_odd", R.UNet, {}, xo, to), ("AttentionUNet_odd", R.AttentionUNet, {}, xo, to),
             ("ResUNet_rgb", R.ResUNet, {"in_channels": 3}, x3, t), ("UNet_rgb", R.UNet, {"in_channels": 3}, x3, t))
    for name, cls, kw, x, t in cases:
        torch.manual_seed(42)
        net = cls(depth=3, **kw)
        net.train()
        sd0 = net.state_dict()
        case = dict(depth=3, x=x, t=t, state_dict_keys=list(sd0.keys()),
                    init_digest={k: dict(sum=float(v.double().sum()), abs_sum=float(v.double().abs().sum()))
                                 for k, v in sd0.items() if v.is_floating_point()})
        logits = net(x)
        bce = nn.BCEWithLogitsLoss()(logits, t)
        dice = DiceLoss()(logits, t)
        (bce + dice).backward()
        case.update(logits=logits.detach().clone(), bce=float(bce), dice=float(dice))
        case["grads"] = {}
        for k, p in net.named_parameters():
            g = p.grad.detach()
            idx = sample_idx(g.numel())
            case["grads"][k] = dict(norm=float(g.double().norm()), idx=idx, vals=g.flatten()[idx].clone())
        net.eval()
        with torch.no_grad():
            le = net(x)
        case["eval_logits"] = le.clone()
        case["eval_mask"] = torch.sigmoid(le) > 0.5
        out[name] = case
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "mod_golden.pt")
    torch.save(out, path)
    print("mod_golden.pt", os.path.getsize(path), "bytes", out["param_count_default"])


if __name__ == "__main__":
    main()
