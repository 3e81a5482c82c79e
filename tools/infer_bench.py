#!/usr/bin/env python
"""UNet inference (BASELINE configs[3]): eval-mode forward + fused threshold masks at 1x512x512.

    python tools/infer_bench.py [--batch 256] [--chunk 64] [--steps 5]

The batch is processed in chunks of `--chunk` frames through UNet.predict_mask (eval BatchNorm folded to a per-channel
affine; the mask = sigmoid(logit) > 0.5 is produced by the head kernel). Inputs are resident on the device; an e2e
variant copies each chunk from pinned host memory and the uint8 masks back. Prints one JSON line."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import b200seg  # noqa
from b200seg import _lib
from b200seg.models.model import UNet
from b200seg.synth import synth_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--chunk", type=int, default=64)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(42)
net = UNet().to(dev).eval()
with torch.no_grad():   # non-degenerate running statistics
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.uniform_(0.0, 0.2); m.running_var.uniform_(0.5, 1.5)
x_cpu, _ = synth_batch(args.chunk, args.size, args.size, seed=1234)
x_pin = x_cpu.pin_memory()
x_dev = x_pin.to(dev)
x_in = torch.empty_like(x_dev)
mask_host = torch.empty((args.chunk, 1, args.size, args.size), dtype=torch.uint8).pin_memory()
chunks = args.batch // args.chunk

def run(e2e):
    for _ in range(chunks):
        if e2e:
            x_in.copy_(x_pin, non_blocking=True)
            _, mask = net.predict_mask(x_in)
            mask_host.copy_(mask, non_blocking=True)
        else:
            net.predict_mask(x_dev)

res = {}
for mode in (False, True):
    for _ in range(2):
        run(mode)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run(mode)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    res["e2e" if mode else "resident"] = {"ms_per_batch": ms, "frames_per_s": chunks * args.chunk / (ms * 1e-3),
                                          "launches_per_batch": (_lib.launch_count() - l0) / args.steps}
flops = 96.184e9 * (args.size / 256) ** 2
res["algorithmic_tflops"] = chunks * args.chunk * flops / (res["resident"]["ms_per_batch"] * 1e-3) / 1e12
res.update({"model": "UNet eval + masks", "batch": chunks * args.chunk, "chunk": args.chunk, "image": f"1x{args.size}x{args.size}",
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30})
print(json.dumps(res))
