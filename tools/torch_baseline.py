#!/usr/bin/env python
"""'Kernel to beat' (BASELINE.md §4): the reference UNet graph on STOCK PyTorch eager (cuDNN/cuBLAS) on one B200.
The drop-in module's parameter containers are real torch layers, so calling them in the reference's order IS the
reference's stock path (models/model.py:53-73). Prints one JSON line per variant. Not part of the product path.

    python tools/torch_baseline.py [--batch 64] [--size 256] [--steps 10]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import b200seg  # noqa: E402,F401
from b200seg.models.model import UNet  # noqa: E402
from b200seg.synth import synth_batch


def stock_forward(m, x):
    e1 = m.encoder1(x)
    e2 = m.encoder2(F.max_pool2d(e1, 2))
    e3 = m.encoder3(F.max_pool2d(e2, 2))
    e4 = m.encoder4(F.max_pool2d(e3, 2))
    d = m.middle(e4)
    d = m.decoder3(torch.cat([d, e4], 1))
    d = m.decoder2(torch.cat([d, e3], 1))
    d = m.decoder1(torch.cat([d, e2], 1))
    return m.final(torch.cat([d, e1], 1))


def dice(logits, t):
    p = torch.sigmoid(logits).reshape(logits.shape[0], -1)
    t = t.reshape(t.shape[0], -1).float()
    return 1 - ((2 * (p * t).sum(1) + 1) / (p.sum(1) + t.sum(1) + 1)).mean()


def run(variant, B, S, steps, warmup=3):
    torch.manual_seed(42)
    m = UNet().cuda().train()
    x, t = synth_batch(B, S, S)
    x, t = x.cuda(), t.cuda()
    if variant == "bf16_channels_last":
        m = m.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-5, fused=True)
    amp = variant != "fp32_tf32"

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            lg = stock_forward(m, x)
            loss = F.binary_cross_entropy_with_logits(lg.float(), t) + dice(lg.float(), t)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"variant": variant, "batch": B, "size": S, "ms_per_step": ms, "images_per_s": B / ms * 1e3,
                      "tflops_algorithmic": B * 288.476e9 * (S / 256) ** 2 / (ms * 1e-3) / 1e12, "loss": float(loss),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--variants", default="fp32_tf32,bf16_nchw,bf16_channels_last")
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    for v in a.variants.split(","):
        try:
            run(v, a.batch, a.size, a.steps)
        except Exception as e:  # keep going: a variant may run out of memory
            print(json.dumps({"variant": v, "error": repr(e)[:300]}), flush=True)
        torch.cuda.empty_cache()
