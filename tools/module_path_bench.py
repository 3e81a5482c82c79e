#!/usr/bin/env python
"""Throughput of the DROP-IN module path (what the reference's unchanged utils/trainer.py exercises): b200seg UNet as an
nn.Module + BCEDiceLoss + torch.optim.AdamW, host-launched, batch 64 @256^2 — next to TrainStep (flat buckets, CUDA graph)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg  # noqa
from b200seg.models.model import UNet
from b200seg.models.loss import BCEDiceLoss
from b200seg.synth import synth_batch
dev = "cuda"
torch.manual_seed(42)
net = UNet().to(dev).train()
opt = torch.optim.AdamW(net.parameters(), lr=1e-5, fused=True)
crit = BCEDiceLoss()
x, t = synth_batch(64, 256, 256, seed=1234)
x, t = x.to(dev), t.to(dev)
def step():
    opt.zero_grad(set_to_none=True)
    loss = crit(net(x), t)
    loss.backward()
    opt.step()
    return loss
for _ in range(4): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): loss = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(json.dumps({"path": "nn.Module UNet + BCEDiceLoss + torch AdamW(fused), host-launched", "ms_per_step": ms, "images_per_s": 64 / (ms * 1e-3), "loss": float(loss.detach())}))
