#!/usr/bin/env python
"""Quick check of the three conv kernel variants against torch fp32 (first used to establish that row-shifted
SWIZZLE_128B operands need descriptor base_offset = 0 on B200; the (addr >> 7) & 7 convention gave wrong results)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg  # noqa
from b200seg import ops

PAIR, HALO, LEGACY = 1 << 10, 1 << 11, 1 << 12
torch.manual_seed(0)
N, H, W, Cin, Cout = 2, 8, 128, 64, 64
x = torch.randn(N, Cin, H, W, device="cuda").to(torch.bfloat16).float()
w = (torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05).to(torch.bfloat16).float()
ref = torch.nn.functional.conv2d(x, w, None, padding=1)
wf, _ = ops.pack_conv_weight(w, want_dgrad=False)
for name, t in (("legacy", LEGACY), ("pair", PAIR), ("halo", HALO)):
    y = ops.Act.empty(N, H, W, Cout, "cuda")
    try:
        ops.conv_fwd(ops.Act.from_nchw(x), wf, None, y, ksize=3, tile_n=t)
        torch.cuda.synchronize()
        err = (y.to_nchw_float() - ref).abs()
        print(f"{name:32s} max err {float(err.max()):.4f}  mean err {float(err.mean()):.5f}  (max|ref| {float(ref.abs().max()):.3f})", flush=True)
    except Exception as e:  # noqa
        print(f"{name:32s} FAILED: {e}", flush=True)
