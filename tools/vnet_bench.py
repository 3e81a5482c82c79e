#!/usr/bin/env python
"""V-Net variant (BASELINE configs[4]): training throughput of b200seg.models.vnet.ImprovedVNet at 1x512x512 on N GPUs.

    python tools/vnet_bench.py [--batch 4] [--steps 5] [--size 512]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/vnet_bench.py --batch 4

One process per GPU; gradients are all-reduced by torch DistributedDataParallel (NCCL buckets overlapped with the
autograd-driven backward of the libb2s nodes); loss = fused BCE + Dice; optimiser = torch AdamW (fused). Prints one JSON
line (rank 0). This is a parity-test configuration, not the bench.py headline."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import b200seg  # noqa
from b200seg import _lib
from b200seg.models.vnet import ImprovedVNet
from b200seg.models.loss import BCEDiceLoss
from b200seg.synth import synth_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--dropout", type=float, default=0.05)
ap.add_argument("--graph", action="store_true", help="capture forward + loss + backward + AdamW in one CUDA graph (1 GPU)")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(42)
net = ImprovedVNet(dropout_rate=args.dropout).to(dev).train()
model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
opt = torch.optim.AdamW(net.parameters(), lr=1e-5, fused=True, capturable=args.graph)
crit = BCEDiceLoss()
x, t = synth_batch(args.batch, args.size, args.size, seed=1234 + rank)
x, t = x.to(dev), t.to(dev)

def step():
    opt.zero_grad(set_to_none=True)
    loss = crit(model(x), t)
    loss.backward()
    opt.step()
    return loss

graph = None
if args.graph and world == 1:
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    opt.zero_grad(set_to_none=True)
    l_cap = _lib.launch_count()
    with torch.cuda.graph(graph):
        static_loss = crit(model(x), t)
        static_loss.backward()
        opt.step()
    captured_launches = _lib.launch_count() - l_cap
    eager_step = step

    def step():
        graph.replay()
        return static_loss
for _ in range(args.warmup):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
l0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms = float(ms) / args.steps
if rank == 0 and os.environ.get("VNET_PROFILE"):
    from b200seg import ops as _ops
    _ops.PROFILE = []
    step(); torch.cuda.synchronize()
    agg = {}
    for n, k, w, s_, e_, _nb in _ops.PROFILE:
        a = agg.setdefault(n, [k, 0, 0.0, 0.0]); a[1] += 1; a[2] += s_.elapsed_time(e_); a[3] += w
    _ops.PROFILE = None
    tot = sum(a[2] for a in agg.values())
    print(f"profiled kernel time {tot:.2f} ms", file=sys.stderr)
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:45]:
        rate = a[3] / (a[2] * 1e-3) / (1e12 if a[0] == "tensor" else 1e9)
        print(f"{n:44s} {a[0]:6s} n={a[1]:3d} ms={a[2]:8.3f} {rate:8.1f} {'TF/s' if a[0]=='tensor' else 'GB/s'}", file=sys.stderr)
if rank == 0:
    flops = 3894e9 * (args.size / 512) ** 2     # ~3 894 GFLOP per image per training step @512^2
    print(json.dumps({"model": "ImprovedVNet (models/vnet.py)", "n_gpus": world, "batch_per_gpu": args.batch, "image": f"1x{args.size}x{args.size}",
                      "ms_per_step": ms, "images_per_s": world * args.batch / (ms * 1e-3), "algorithmic_tflops_per_gpu": args.batch * flops / (ms * 1e-3) / 1e12,
                      "loss": float(loss), "libb2s_launches_per_step": captured_launches if graph is not None else (_lib.launch_count() - l0) / args.steps, "cuda_graph": graph is not None,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
if world > 1:
    dist.destroy_process_group()
