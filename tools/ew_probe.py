#!/usr/bin/env python
"""Per-shape timing of the V-Net bandwidth kernels (SE pool / scale, channel sums, slice copy, BN-act) at the five
resolution levels of a 512^2 batch: python tools/ew_probe.py [--batch 16]. CUDA events, L2 flushed between launches."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import b200seg  # noqa
from b200seg import ops, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--c1", action="store_true", help="time the Cin = 1 first-conv kernels instead")
args = ap.parse_args()
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L = _lib.lib()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps + 1):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


if args.c1:
    for (N, S) in ((64, 256), (16, 512)):
        x = torch.rand((N, 1, S, S), device=dev)
        w, b = torch.randn((64, 1, 3, 3), device=dev) * 0.3, torch.randn(64, device=dev) * 0.1
        r = ops.Act.empty(N, S, S, 64, dev)
        stats = torch.empty(ops.c1_rows(N, S, S) * 2 * 64, dtype=torch.float32, device=dev)
        partial = torch.empty(ops.c1_rows(N, S, S) * 64 * 9, dtype=torch.float32, device=dev)
        scratch = torch.empty(128 * 64 * 9, dtype=torch.float32, device=dev)
        dw = torch.empty((64, 1, 3, 3), dtype=torch.float32, device=dev)
        tf = timed(lambda: ops.conv3x3_c1_fwd(x, w, b, r, relu=True, stats=stats))
        tw = timed(lambda: L.b2s_conv3x3_c1_wgrad(ops._p(x), r.ptr, ops._p(partial), N, S, S, 64, ops._stream()))
        print(json.dumps({"shape": [N, 1, S, S], "c1_fwd_us": round(tf * 1e3, 1),
                          "c1_wgrad_us": round(tw * 1e3, 1), "checksum": float(r.buf.float().sum())}), flush=True)
    sys.exit(0)

rows = []
for l in range(5):
    H = args.size >> l
    C = 64 << l
    N = args.batch
    x = ops.Act(torch.randn((N, H, H, C), device=dev).to(torch.bfloat16))
    y = ops.Act(torch.randn((N, H, H, C), device=dev).to(torch.bfloat16))
    out = ops.Act.empty(N, H, H, C, dev)
    nbytes = N * H * H * C * 2
    HW = H * H
    chunks = L.b2s_se_chunks(HW)
    partial = torch.empty(N * chunks * C, dtype=torch.float32, device=dev)
    gate = torch.rand((N, C), device=dev)
    sums = torch.empty(C, dtype=torch.float32, device=dev)
    st = ops._stream
    r = {"level": l, "shape": [N, H, H, C], "MB": nbytes / 1e6}
    t = timed(lambda: L.b2s_se_pool(x.ptr, x.cstride, None, 0, ops._p(partial), N, HW, C, st()))
    r["se_pool"] = [round(t * 1e3, 1), round(nbytes / t / 1e6)]
    t = timed(lambda: L.b2s_se_pool(x.ptr, x.cstride, y.ptr, y.cstride, ops._p(partial), N, HW, C, st()))
    r["se_pool_dot"] = [round(t * 1e3, 1), round(2 * nbytes / t / 1e6)]
    t = timed(lambda: L.b2s_se_scale(x.ptr, x.cstride, ops._p(gate), None, 0.0, out.ptr, out.cstride, N, HW, C, st()))
    r["se_scale"] = [round(t * 1e3, 1), round(2 * nbytes / t / 1e6)]
    t = timed(lambda: ops.channel_sums(x, sums))
    r["channel_sums(+reduce)"] = [round(t * 1e3, 1), round(nbytes / t / 1e6)]
    t = timed(lambda: ops.copy_channels(x, out))
    r["copy_channels"] = [round(t * 1e3, 1), round(2 * nbytes / t / 1e6)]
    rows.append(r)
    print(json.dumps(r), flush=True)
print("columns: [microseconds, GB/s]")
