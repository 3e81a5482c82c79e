#!/usr/bin/env python
"""N-rank NCCL correctness check of the data-parallel step, launched by tests/test_ddp_gpu.py (or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        tests/run_ddp_check.py [--size 64]

(1) bench.dp_check: all-reduced gradients == sum over single-GPU chunked replicas (per-replica BatchNorm statistics, the
    reference's nn.DataParallel semantics, utils/trainer.py:28-30); one captured step from identical parameters gives
    bit-identical parameters on every rank and matches the chunked emulation.
(2) The CUDA-graph step with in-graph bucketed all-reduces == the host-launched bucketed step over several steps with a
    changing learning rate (parameters, BatchNorm buffers, losses), on every rank.
Rank 0 prints one line: DDP_JSON {...}.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("B2S_HANG_DUMP"):          # debugging aid: dump every thread's stack if the run is still alive after N s
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ["B2S_HANG_DUMP"]), exit=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--batch", type=int, default=4)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import b200seg  # noqa: F401
    from b200seg.models.model import UNet
    from b200seg.train import TrainStep, ranks_agree
    from b200seg.synth import synth_batch
    import bench

    world, rank, local = (int(os.environ[k]) for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(42)
    sd = UNet().state_dict()
    S = args.size
    report = {"world": world, "size": S}
    report["dp_check"] = bench.dp_check(world, rank, dev, sd, S, vb=args.batch)

    # (2) graph (in-graph NCCL buckets) vs host-launched bucketed step; tiny buckets so that several are in flight
    x, t = synth_batch(args.batch, S, S, seed=100 + rank)
    x2, t2 = synth_batch(args.batch, S, S, seed=200 + rank)
    x, t, x2, t2 = (v.to(dev) for v in (x, t, x2, t2))
    eager = TrainStep(sd, dev, lr=1e-3, bucket_mb=4.0)
    graph = TrainStep(sd, dev, lr=1e-3, bucket_mb=4.0)
    ok = graph.capture(x, t)
    for _ in range(2):
        eager.step(x, t)
    losses = []
    for i, lr in enumerate([1e-3, 5e-4, 2e-3]):
        xb, tb = (x, t) if i % 2 == 0 else (x2, t2)
        le = eager.step(xb, tb, lr=lr).clone()
        lg = graph.step_graphed(xb, tb, lr=lr).clone()
        losses.append((float(le[0]), float(lg[0])))
    torch.cuda.synchronize()
    se, sg = eager.state_dict(), graph.state_dict()
    worst = 0.0
    for k in se:
        a, b = se[k].double(), sg[k].double()
        worst = max(worst, float((a - b).abs().max() / a.abs().max().clamp_min(1e-12)))
    flag = torch.tensor([worst], device=dev, dtype=torch.float64)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    report["graph_vs_eager"] = {"captured": bool(ok), "comm_in_graph": bool(getattr(graph, "graph_comm", False)),
                                "capture_error": getattr(graph, "capture_error", ""), "buckets": len(graph.buckets),
                                "losses": losses, "max_rel_param_diff_over_ranks": float(flag.item()),
                                "graph_ranks_agree": bool(ranks_agree(graph.flat_p)),
                                "eager_ranks_agree": bool(ranks_agree(eager.flat_p)),
                                "bit_identical_rank0": bool(torch.equal(eager.flat_p, graph.flat_p))}
    dist.barrier()
    if rank == 0:
        print("DDP_JSON " + json.dumps(report), flush=True)
    graph.release_graph()          # a live graph with captured collectives blocks the communicator teardown
    del graph, eager
    bench.shutdown_process_group()


if __name__ == "__main__":
    main()
