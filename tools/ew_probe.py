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
