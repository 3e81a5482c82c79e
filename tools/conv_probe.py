#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 conv kernels at the UNet layer shapes (batch 64 @256^2 by default).

    python tools/conv_probe.py [--batch 64] [--only REGEX] [--iters 10] [--tile-n N] [--json out.json]

Each case is timed with CUDA events over `iters` launches that rotate through enough distinct buffers to exceed
the 126 MB L2 (or all of them when the tensors are larger than L2 anyway). Prints achieved TFLOP/s per case.
Used to iterate on csrc/conv_tc.cu and as the short command for `ncu --set full -k regex:...` captures.
"""
import argparse
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import b200seg  # noqa: E402,F401
from b200seg import ops  # noqa: E402

DEV = "cuda"


def unet_cases(B, S):
    """(kind, Cin, Cout, H) of the distinct tensor-core launches of one UNet training step."""
    cs = []
    enc = [(64, 64, S), (64, 128, S // 2), (128, 128, S // 2), (128, 256, S // 4), (256, 256, S // 4),
           (256, 512, S // 8), (512, 512, S // 8), (512, 1024, S // 16), (1024, 1024, S // 16)]
    dec = [(1024, 512, S // 8), (512, 256, S // 4), (256, 128, S // 2), (128, 64, S)]
    for cin, cout, h in enc + dec:
        cs.append(("fwd", cin, cout, h))
    for cin, cout, h in enc + dec:
        cs.append(("wgrad", cin, cout, h))
    for cin, cout, h in [(128, 64, S // 2), (256, 128, S // 4), (512, 256, S // 8), (1024, 512, S // 16)]:
        cs.append(("convT", cin, cout, h))
    return cs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--tile-n", type=int, default=0)
    ap.add_argument("--splits", type=int, default=0)
    ap.add_argument("--json", default="")
    ap.add_argument("--no-bias", action="store_true")
    ap.add_argument("--nbuf", type=int, default=0, help="number of rotating buffer sets (1: L2-resident operands)")
    ap.add_argument("--concat", action="store_true", help="convT: write into the first half of a concat buffer")
    ap.add_argument("--convt-mode", default="fwd", choices=["fwd", "dgrad", "wgrad"])
    args = ap.parse_args()
    B = args.batch
    results = []
    for kind, cin, cout, h in unet_cases(B, args.size):
        name = f"{kind}[{cin}->{cout}@{h}x{h}]"
        if args.only and not re.search(args.only, name):
            continue
        torch.manual_seed(0)
        in_bytes = B * h * h * cin * 2
        nbuf = args.nbuf if args.nbuf > 0 else max(2, min(8, int(200e6 // max(in_bytes, 1)) + 1))
        xs = [ops.Act(torch.randn((B, h, h, cin), device=DEV).to(torch.bfloat16)) for _ in range(nbuf)]
        if kind == "fwd":
            w = torch.randn((cout, cin, 3, 3), device=DEV) * 0.05
            wf, _ = ops.pack_conv_weight(w, want_dgrad=False)
            bias = torch.randn(cout, device=DEV)
            ys = [ops.Act.empty(B, h, h, cout, DEV) for _ in range(nbuf)]
            rows = ops.conv_stats_rows(B, h, h, cout, args.tile_n)
            stats = torch.empty((rows, 2, cout), dtype=torch.float32, device=DEV)
            flops = 18.0 * B * h * h * cin * cout
            fn = lambda i: ops.conv_fwd(xs[i % nbuf], wf, bias, ys[i % nbuf], ksize=3, relu=True, stats=stats,
                                        tile_n=args.tile_n)
        elif kind == "wgrad":
            dzs = [ops.Act(torch.randn((B, h, h, cout), device=DEV).to(torch.bfloat16)) for _ in range(nbuf)]
            nbytes, _ = ops.wgrad_workspace(B, h, h, cin, cout, 9, args.tile_n, args.splits)
            ws = torch.empty(nbytes // 4, dtype=torch.float32, device=DEV)
            dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=DEV)
            flops = 18.0 * B * h * h * cin * cout
            fn = lambda i: ops.conv3x3_wgrad(xs[i % nbuf], dzs[i % nbuf], ws, dw, tile_n=args.tile_n,
                                             splits=args.splits)
        else:
            w = torch.randn((cin, cout, 2, 2), device=DEV) * 0.05
            wf, _ = ops.pack_convt_weight(w)
            bias = torch.randn(cout, device=DEV)
            if args.concat:   # the engine's layout: the up-conv writes the first half of a [.., 2*Cout] concat buffer
                ys = [ops.Act.empty(B, 2 * h, 2 * h, 2 * cout, DEV).slice(0, cout) for _ in range(nbuf)]
            else:
                ys = [ops.Act.empty(B, 2 * h, 2 * h, cout, DEV) for _ in range(nbuf)]
            flops = 8.0 * B * h * h * cin * cout
            gb = (2.0 * B * h * h * (cin + 4 * cout)) / 1e9
            if args.convt_mode == "dgrad":
                _, wdp = ops.pack_convt_weight(w)
                fn = lambda i: ops.convt_dgrad(ys[i % nbuf], wdp, xs[i % nbuf], tile_n=args.tile_n)
            elif args.convt_mode == "wgrad":
                nbytes, _ = ops.wgrad_workspace(B, h, h, cin, cout, 4, args.tile_n, args.splits)
                ws = torch.empty(nbytes // 4, dtype=torch.float32, device=DEV)
                dw = torch.empty((cin, cout, 2, 2), dtype=torch.float32, device=DEV)
                fn = lambda i: ops.convt_wgrad(xs[i % nbuf], ys[i % nbuf], ws, dw, tile_n=args.tile_n, splits=args.splits)
            else:
                fn = lambda i: ops.convt_fwd(xs[i % nbuf], wf, None if args.no_bias else bias, ys[i % nbuf], tile_n=args.tile_n)
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        tf = flops / (ms * 1e-3) / 1e12
        extra = f"  {gb / (ms * 1e-3):8.1f} GB/s" if kind == "convT" else ""
        print(f"{name:34s} {ms * 1e3:9.1f} us  {tf:8.1f} TFLOP/s{extra}", flush=True)
        results.append({"name": name, "us": ms * 1e3, "tflops": tf})
        del xs
        torch.cuda.empty_cache()
    if args.json:
        with open(args.json, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
