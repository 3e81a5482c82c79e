"""Synthetic 1-channel "ultrasound-shaped" frames and nodule masks: the workload generator of bench.py
and the tools. The reference trains on DDTI images it does not ship (data/data_loader.py); there is no dataset here.

tests/test_oracle_cpu.py checks that this generator and the oracle's own copy produce identical tensors."""
import math

import torch


def synth_batch(B, H, W, seed=1234, device="cpu"):
    """(image [B,1,H,W] fp32 in [0,1], mask [B,1,H,W] fp32 in {0,1}).

    image = low-frequency tissue field x (1 - 0.6 nodule) x Rayleigh speckle, clamped; mask = a rotated ellipse."""
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
