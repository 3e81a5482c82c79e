#!/usr/bin/env python
"""Benchmark of the B200 UNet hot path (contract: see the task's bench.py section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step = forward + fused BCE/Dice loss + backward + AdamW of the reference UNet (models/model.py) on one batch of
synthetic 1x256x256 ultrasound-shaped frames, bf16 storage / fp32 accumulate, per-GPU batch 64 (BASELINE
configs[1]); with N GPUs the batch is sharded (weak scaling: global batch 64*N, 512 at N=8 = configs[2]) and
gradients are all-reduced in buckets over NCCL, overlapped with backward.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_IMG_TRAIN_256 = 288.476e9   # fwd + dgrad + wgrad at 1x256x256
METRIC = "UNet train images/sec @256^2 bf16"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms; samples are attributed to a timed region by their
    timestamps (the sampler is started before warm-up because nvidia-smi takes ~1 s to emit its first line)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = f"/tmp/b2s_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()

    def summary(self, t0, t1):
        """Samples with t0 <= timestamp <= t1 (time.time() seconds, local clock)."""
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            lines = open(self.path).read().splitlines()
        except OSError:
            lines = []
        for line in lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if ts < t0 - 0.05 or ts > t1 + 0.05:
                    continue
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for n, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in the timed region"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def workload_config(B, S, world):
    """config block shared by both arms (the reference arm adds what it sampled)"""
    return {"workload": f"UNet bf16 training, batch {B} at {S}x{S} per GPU (BASELINE configs[1]; "
                        f"global batch {B * world})", "batch_per_gpu": B, "global_batch": B * world,
            "image": f"1x{S}x{S}", "loss": "BCE+Dice (fused)", "optimizer": "AdamW lr 1e-5 (fused, flat buckets)",
            "parallelism": f"dp{world}", "precision": "bf16 storage / fp32 accumulate, fp32 master weights",
            "l2": "working set (>= 6 GB of activations per step) is far larger than the 126 MB L2"}


def reference_config_note():
    return ("BASELINE configs[0]: the reference's CPU path, fp32, batch 4 at 256x256, train-mode forward + BCE + Dice + "
            "backward (no optimiser step), SURVEY §8(d)")


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path — its unmodified nn.Modules from
    baseline/_ref (tools/stage_reference.py) on all host cores. Runs on rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tools import bench_legs
    threads = os.cpu_count() or 1
    batch = 4   # bounded sample of the batch-64 workload: 4 images per step (= BASELINE configs[0]'s batch)
    steps = max(1, min(args.steps, 6))
    warm = 1
    times, kind, note = bench_legs.cpu_reference_times(batch, args.size, steps, warm, threads)
    ms = 1e3 * sum(times) / len(times)
    value = batch / (ms / 1e3)
    sample = (f"each step = {note}, on {batch} of the {args.batch} images of a batch ({args.size}x{args.size}), "
              f"{threads} threads; {steps} timed steps after {warm} warm-up")
    cfg = workload_config(args.batch, args.size, 1)
    cfg["sample"] = sample
    cfg["precision"] = "fp32 (the reference's CPU path)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# kernel families of the per-launch roofline table: (family, predicate on the ops.py record name, bound, what limits it)
def _dims(name):
    """'conv3x3[64->128@256x256]' -> (64, 128, 256, 256)"""
    import re
    m = re.search(r"\[(\d+)(?:->|<-)(\d+)@(\d+)x(\d+)\]", name)
    return tuple(int(v) for v in m.groups()) if m else None


def family_of(name, kind, nbytes):
    d = _dims(name)
    if name.startswith("conv3x3["):
        if d[3] >= 256:
            return "conv3x3 fwd+dgrad @256^2 (N<=128 tiles)"
        if d[3] >= 128:
            return "conv3x3 fwd+dgrad @128^2"
        return "conv3x3 fwd+dgrad deep (<=64^2)"
    if name.startswith("wgrad3x3["):
        return "wgrad3x3 @256^2/128^2 (halo kernel)" if d[3] >= 128 else "wgrad3x3 deep (<=64^2)"
    if name.startswith("convT_"):
        # arithmetic intensity below the ridge (~250 FLOP/B) => HBM-bound: reported in GB/s
        return "convT high-res (HBM-bound, GB/s)" if d[3] >= 64 else "convT deep (tensor-bound)"
    if name.startswith(("bn_apply", "bn_bwd")):
        return "BatchNorm apply / backward passes"
    if name.startswith(("conv3x3_c1", "head_")):
        return "first conv (Cin=1) + 1x1 head"
    if name.startswith("seg_loss"):
        return "Dice+BCE loss fwd/bwd"
    if name.startswith("adamw"):
        return "AdamW"
    if name.startswith(("wgrad_reduce", "wgradT_reduce", "pack_")):
        return "split-K reduce + weight packing"
    return "other"


FAMILY_NOTE = {
    "conv3x3 fwd+dgrad @256^2 (N<=128 tiles)": "tensor pipe 46 % active under ncu (N=64 MMAs: shared-memory bandwidth, 668 KB per item)",
    "conv3x3 fwd+dgrad @128^2": "tensor pipe; halo kernel",
    "conv3x3 fwd+dgrad deep (<=64^2)": "tensor pipe 85 % active under ncu (tile-pair kernel, L2 -> smem)",
    "wgrad3x3 @256^2/128^2 (halo kernel)": "tensor pipe 50-59 % active under ncu",
    "wgrad3x3 deep (<=64^2)": "tensor pipe 91 % active under ncu",
    "convT high-res (HBM-bound, GB/s)": "dram throughput (arithmetic intensity 85-171 FLOP/B < ridge)",
    "convT deep (tensor-bound)": "tensor pipe",
    "BatchNorm apply / backward passes": "dram throughput 72-79 % under ncu (bytes in flight)",
    "first conv (Cin=1) + 1x1 head": "issue slots 64-68 % under ncu (instruction-issue bound)",
    "Dice+BCE loss fwd/bwd": "launch/latency sized (33.5 MB)",
    "AdamW": "dram throughput",
    "split-K reduce + weight packing": "latency / instruction issue",
}


def shutdown_process_group(timeout_s=20.0):
    """destroy_process_group() with a deadline: the result line is already printed when this runs, a slow NCCL
    teardown must not turn a finished run into a hang."""
    import threading
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    t = threading.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(timeout_s)
    if t.is_alive():
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def dp_check(world, rank, dev, sd, S, no_graph=False, vb=8):
    """Correctness of the N>1 path on this hardware against the reference's data-parallel semantics (per-replica
    BatchNorm statistics, summed gradients: utils/trainer.py:28-30): fresh TrainSteps, the SAME code path as the timed
    one, a small batch. Collective: every rank must call it. Returns the report (complete on rank 0)."""
    import torch
    import torch.distributed as dist
    from b200seg.train import TrainStep, emulate_replicas, ranks_agree
    from b200seg.synth import synth_batch
    vx, vt = synth_batch(vb, S, S, seed=777 + rank)
    vx, vt = vx.to(dev), vt.to(dev)
    # (a) gradients: host-launched step with the bucketed all-reduce, same initial parameters on every rank
    vts = TrainStep(sd, dev, lr=0.0)
    vts.forward_backward(vx, vt)
    torch.cuda.synchronize()
    g_ddp = vts.flat_g.clone()
    # (b) parameters: the captured step (in-graph buckets). capture() runs two real eager steps first; with lr = 0
    # they leave the parameters alone (Adam moments, step count and BatchNorm buffers advance), so the ONE graphed
    # update that follows starts from identical parameters on every path and differs only by summation order
    v_graphed = (not no_graph) and vts.capture(vx, vt)
    (vts.step_graphed if v_graphed else vts.step)(vx, vt, lr=1e-3)
    torch.cuda.synchronize()
    agree = ranks_agree(vts.flat_p)
    gathered = [torch.empty_like(vts.flat_p) for _ in range(world)] if rank == 0 else None
    dist.gather(vts.flat_p, gathered, dst=0)
    vts.release_graph()
    check = {"ranks_hold_bit_identical_parameters": bool(agree), "batch_per_rank": vb,
             "path": "CUDA graph with in-graph NCCL buckets" if v_graphed and vts.graph_comm else
                     ("CUDA graph + flat all-reduce" if v_graphed else "host-launched bucketed")}
    if rank == 0:
        shards = []
        for r in range(world):
            sx, st = synth_batch(vb, S, S, seed=777 + r)
            shards.append((sx.to(dev), st.to(dev)))
        reps = emulate_replicas(sd, dev, shards, lrs=[None])
        g_ref = reps[0].flat_g
        gd = (g_ddp.double() - g_ref.double()).abs().max().item()
        gmax = g_ref.double().abs().max().item()
        check["gradient_vs_chunked_replicas"] = {"max_abs_diff": gd, "max_abs_grad": gmax,
                                                 "bit_identical": bool(torch.equal(g_ddp, g_ref))}
        del reps
        reps = emulate_replicas(sd, dev, shards, lrs=[None, 0.0, 0.0, 1e-3])    # gradient probe + 2 capture steps + 1
        ref = reps[0].flat_p
        init = TrainStep(sd, dev, lr=0.0, use_dist=False).flat_p
        upd_ref = (ref.double() - init.double())
        upd = (gathered[0].double() - init.double())
        rel = float((upd - upd_ref).norm() / upd_ref.norm().clamp_min(1e-30))
        nd = int((gathered[0] != ref).sum())
        check["update_vs_chunked_replicas"] = {"rel_l2_of_update": rel, "elements_differing": nd,
                                               "elements": int(ref.numel()),
                                               "bit_identical": bool(torch.equal(gathered[0], ref))}
        # BatchNorm running statistics stay per replica; rank 0 keeps replica 0's (what DataParallel checkpoints)
        rm = max(float((vts.P[k].double() - reps[0].P[k].double()).abs().max()) for k in vts.P if "running" in k)
        check["running_stats_max_abs_diff_rank0"] = rm
        # world 2: a + b is the same sum in any order -> bit-identical expected; more ranks: summation order only
        check["ok"] = bool(agree and gd <= 1e-5 * max(gmax, 1e-30) and rel <= 1e-2 and rm <= 1e-5)
        del reps, shards, init
    return check


def run_b200(args):
    import torch
    import torch.distributed as dist
    import b200seg  # noqa: F401
    from b200seg import _lib, ops
    from b200seg.models.model import UNet
    from b200seg.train import TrainStep, emulate_replicas, ranks_agree
    from b200seg.synth import synth_batch
    from tools import bench_legs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")           # keep NCCL's version banner off stdout: one JSON line only
        if args.nccl_max_ctas > 0:       # bound the SMs the overlapped all-reduces may take from the compute kernels
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = args.nccl_max_ctas
            opts.config.min_ctas = min(args.nccl_max_ctas, 4)
            dist.init_process_group("nccl", device_id=dev, pg_options=opts)
        else:
            dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    B, S = args.batch, args.size
    torch.manual_seed(42)
    sd = UNet().state_dict()
    ts = TrainStep(sd, dev, lr=1e-5, bucket_mb=args.bucket_mb)
    x_cpu, t_cpu = synth_batch(B, S, S, seed=1234 + rank)
    x_pin, t_pin = x_cpu.pin_memory(), t_cpu.pin_memory()
    x_dev, t_dev = x_pin.to(dev, non_blocking=True), t_pin.to(dev, non_blocking=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        ts.step(x_dev, t_dev)
    barrier()
    graphed = (not args.no_graph) and ts.capture(x_dev, t_dev)
    run_step = ts.step_graphed if graphed else ts.step
    capture_error = getattr(ts, "capture_error", "")
    for _ in range(2):
        run_step(x_dev, t_dev)
    barrier()

    def timed(fn, steps):
        barrier()
        w0 = time.time()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count() - l0
        w1 = time.time()
        barrier()
        return max_over_ranks(ms), launches, (w0, w1)

    # (1) device-resident inputs
    ms_total, launches, clocks = timed(lambda: run_step(x_dev, t_dev), args.steps)
    if graphed:   # replays bypass the library's host-side launch counter: count the captured kernels
        launches += ts.graph_launches * args.steps

    # (2) end to end: every step's inputs come from pinned host memory (H2D inside the timed region, issued on a copy
    # stream into a double buffer so that the copy of step i+1 overlaps the compute of step i) and every step's loss
    # vector is copied back to pinned host memory; the host reads the loss of step i-1 while step i runs (no per-step
    # device-wide synchronisation, which is how a training loop that logs the loss uses the API).
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(x_dev), torch.empty_like(t_dev)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]     # H2D of buffer b finished
    consumed = [torch.cuda.Event() for _ in range(2)]  # the step reading buffer b finished
    loss_ring = [torch.empty(8, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_done = [torch.cuda.Event() for _ in range(2)]
    state = {"i": 0, "last": None}

    def stage_inputs(b):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])
            bufs[b][0].copy_(x_pin, non_blocking=True)
            bufs[b][1].copy_(t_pin, non_blocking=True)
            ready[b].record(copy_stream)

    main = torch.cuda.current_stream()
    for b in range(2):
        consumed[b].record(main)
    stage_inputs(0)

    def e2e_step():
        i = state["i"]
        b = i & 1
        stage_inputs(b ^ 1)                 # next step's inputs, overlapping this step's kernels
        main.wait_event(ready[b])
        out = run_step(bufs[b][0], bufs[b][1])
        consumed[b].record(main)
        loss_ring[b].copy_(out, non_blocking=True)
        loss_done[b].record(main)
        if i > 0:                           # read the previous step's loss on the host
            loss_done[b ^ 1].synchronize()
            state["last"] = float(loss_ring[b ^ 1][0])
        state["i"] = i + 1
    e2e_step()
    ms_e2e, _, _ = timed(e2e_step, args.steps)

    if sampler:
        time.sleep(0.3)
        sampler.stop()
        clocks = sampler.summary(*clocks)
        try:
            os.remove(sampler.path)
        except OSError:
            pass
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    peaks = load_peaks()
    extra = {}

    # (2b) data parallel: how much of the step is communication that backward does not hide, and do the ranks agree
    if world > 1:
        # the same captured step WITHOUT any collective (a world-1 TrainStep on the same engine and shapes)
        # (own engine: a captured graph bakes in the addresses of its engine's packed-weight buffers)
        nocomm = TrainStep(sd, dev, lr=1e-5, use_dist=False)
        nc_graphed = (not args.no_graph) and nocomm.capture(x_dev, t_dev)
        nc_step = nocomm.step_graphed if nc_graphed else nocomm.step
        for _ in range(2):
            nc_step(x_dev, t_dev)
        k_nc = max(10, args.steps // 2)
        ms_nc, _, _ = timed(lambda: nc_step(x_dev, t_dev), k_nc)
        ms_again, _, _ = timed(lambda: run_step(x_dev, t_dev), k_nc)     # interleaved re-measurement of the real step
        extra["exposed_comm_ms"] = ms_again / k_nc - ms_nc / k_nc
        extra["exposed_comm_note"] = (f"step with in-graph bucketed all-reduce ({ms_again / k_nc:.3f} ms) minus the same "
                                      f"captured step without collectives ({ms_nc / k_nc:.3f} ms), both max over ranks, "
                                      f"{k_nc} steps each; buckets: {len(ts.buckets)}; what stays exposed is the last "
                                      "bucket's all-reduce (issued after encoder1's weight gradient) plus rank skew")
        extra["comm_in_graph"] = bool(graphed and getattr(ts, "graph_comm", False))
        nocomm.release_graph()
        del nocomm, nc_step
        check = dp_check(world, rank, dev, sd, S, no_graph=args.no_graph)
        extra["dp_check"] = check
        torch.cuda.empty_cache()
        # BASELINE configs[2] as written: GLOBAL batch 512 (strong scaling: 256 / 128 / 64 images per GPU at 2 / 4 / 8)
        gb = 512 // world
        if gb != B and not args.no_extras:
            gx, gt = synth_batch(gb, S, S, seed=4321 + rank)
            gx, gt = gx.to(dev), gt.to(dev)
            gts = TrainStep(sd, dev, lr=1e-5)
            g_graphed = (not args.no_graph) and gts.capture(gx, gt)
            g_step = gts.step_graphed if g_graphed else gts.step
            for _ in range(2):
                g_step(gx, gt)
            k_g = 10
            ms_g, _, _ = timed(lambda: g_step(gx, gt), k_g)
            extra["configs2_global_batch_512"] = {"batch_per_gpu": gb, "n_gpus": world, "ms_per_step": ms_g / k_g,
                                                  "images_per_s": 512 * k_g / (ms_g * 1e-3), "steps": k_g,
                                                  "cuda_graph": bool(g_graphed)}
            gts.release_graph()
            del gts, gx, gt
            torch.cuda.empty_cache()
        elif gb == B:
            extra["configs2_global_batch_512"] = "this run (64 images per GPU x 8 GPUs)"

    # (3) roofline pass: per-launch CUDA events around every kernel of one extra host-launched step (not part of the
    # timed region). The GPU idles between these launches, so they run at burst clocks: fractions are quoted against
    # the BURST peak (MEASURED_PEAKS bf16_tflops); the step-level fraction is quoted against the sustained one.
    ts.forward_backward(x_dev, t_dev, reduce=False)     # un-profiled lead-in: the profiled pass starts on a busy GPU
    ops.PROFILE = []
    ts.forward_backward(x_dev, t_dev, reduce=False)
    ops.adamw_step(ts.flat_p, ts.flat_g, ts.flat_m, ts.flat_v, 0.0, 0.9, 0.999, 1e-8, 0.0, 1, 0.0)   # lr 0: timing only
    torch.cuda.synchronize()
    recs = [(n, k, w, s.elapsed_time(e), nb) for (n, k, w, s, e, nb) in ops.PROFILE]
    ops.PROFILE = None
    tc = [r for r in recs if r[1] == "tensor"]
    hb = [r for r in recs if r[1] == "hbm"]
    tc_ms, tc_flops = sum(r[3] for r in tc), sum(r[2] for r in tc)
    hb_ms, hb_bytes = sum(r[3] for r in hb), sum(r[2] for r in hb)
    achieved_tf = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    # dominant kernel: conv2_tc_kernel = every 3x3 conv forward and input-gradient launch (ops.conv_fwd)
    dom = [r for r in tc if r[0].startswith("conv3x3[")]
    dom_ms, dom_flops = sum(r[3] for r in dom), sum(r[2] for r in dom)
    dom_tf = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    fams = {}
    for n, k, w, ms, nb in recs:
        f = fams.setdefault(family_of(n, k, nb), {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        f["launches"] += 1
        f["ms"] += ms
        if k == "tensor":
            f["flops"] += w
            f["bytes"] += nb or 0.0
        else:
            f["bytes"] += w
    families = []
    for name, f in sorted(fams.items(), key=lambda kv: -kv[1]["ms"]):
        hbm_bound = f["flops"] == 0.0 or "HBM-bound" in name
        if hbm_bound:
            ach = f["bytes"] / (f["ms"] * 1e-3) / 1e9
            families.append({"family": name, "bound": "hbm", "launches": f["launches"], "ms_per_step": f["ms"],
                             "achieved": ach, "unit": "GB/s", "frac_of_measured_hbm": ach / peaks["hbm_gbs"],
                             "limited_by": FAMILY_NOTE.get(name, "")})
        else:
            ach = f["flops"] / (f["ms"] * 1e-3) / 1e12
            families.append({"family": name, "bound": "tensor", "launches": f["launches"], "ms_per_step": f["ms"],
                             "achieved": ach, "unit": "TFLOP/s", "frac_of_measured_burst": ach / peaks["tf_burst"],
                             "limited_by": FAMILY_NOTE.get(name, "")})
    traffic, traffic_note = None, "no ncu capture found under profiles/"
    try:   # DRAM bytes of one launch from the committed `ncu --set full` capture (profiles/)
        with open(os.path.join(ROOT, "profiles", "r2_ncu_conv_full.json")) as f:
            for d in json.load(f)["kernels"]:
                if d["kernel"].startswith("void conv2_tc_kernel<256, 2, 0, 3, 0"):
                    def to_bytes(v):
                        x, u = v.split()
                        return float(x) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
                    traffic = to_bytes(d["dram_read"]) + to_bytes(d["dram_write"])
                    traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum of one conv2_tc_kernel<256,2,0,3,0> launch "
                                    "(512->512 @32x32, batch 64; algorithmic 2 x 67.1 MB activations + 4.7 MB weights; the "
                                    "input is partly L2-resident), profiles/r2_ncu_conv_full.json; the 64->64 @256x256 launch "
                                    "moves 537 + 491 MB = its algorithmic bytes")
                    break
    except (OSError, KeyError, ValueError):
        pass
    step_tf = (B * FLOP_PER_IMG_TRAIN_256 * (S / 256) ** 2) / (ms_step * 1e-3) / 1e12
    roofline = {"bound": "tensor",
                "kernel": "conv2_tc_kernel: all 3x3 conv forward + input-gradient launches of one step (tcgen05 implicit GEMM)",
                "achieved": dom_tf, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                "frac": dom_tf / peaks["tf_burst"],
                "peak_source": f"{peaks['source']} bf16_tflops (burst: the launches are timed one by one with CUDA events "
                               "in a host-launched pass, the GPU idles between them)",
                "traffic": traffic, "traffic_note": traffic_note, "launches": len(dom), "ms_per_step_in_kernel": dom_ms,
                "share_of_step": dom_ms / (tc_ms + hb_ms) if tc_ms + hb_ms > 0 else None,
                "all_tcgen05": {"kernels": "conv2_tc_kernel + wgrad_halo_kernel + wgrad_tc_kernel (conv, transposed conv, "
                                           "weight gradients)", "achieved": achieved_tf,
                                "frac": achieved_tf / peaks["tf_burst"], "launches": len(tc), "ms_per_step": tc_ms,
                                "share_of_step": tc_ms / (tc_ms + hb_ms) if tc_ms + hb_ms > 0 else None},
                "step": {"achieved": step_tf, "unit": "TFLOP/s (288.476 GFLOP per image, whole graph-replayed step)",
                         "frac_of_sustained": step_tf / peaks["tf_sustained"], "frac_of_burst": step_tf / peaks["tf_burst"]},
                "hbm_kernels": {"achieved_gbs": hb_bytes / (hb_ms * 1e-3) / 1e9 if hb_ms > 0 else 0.0,
                                "peak_gbs": peaks["hbm_gbs"], "ms_per_step": hb_ms, "launches": len(hb)},
                "sum_of_kernel_ms": tc_ms + hb_ms, "families": families}
    if rank == 0 and args.profile_out:
        agg = {}
        for n, k, w, ms, nb in recs:
            a = agg.setdefault(n, {"kind": k, "launches": 0, "ms": 0.0, "work": 0.0, "bytes": 0.0})
            a["launches"] += 1; a["ms"] += ms; a["work"] += w; a["bytes"] += nb or 0.0
        for n, a in agg.items():
            rate = a["work"] / (a["ms"] * 1e-3) if a["ms"] > 0 else 0.0
            a["achieved"] = rate / 1e12 if a["kind"] == "tensor" else rate / 1e9
            a["unit"] = "TFLOP/s" if a["kind"] == "tensor" else "GB/s"
            if a["kind"] == "tensor" and a["bytes"] > 0:
                a["achieved_gbs"] = a["bytes"] / (a["ms"] * 1e-3) / 1e9
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
        with open(args.profile_out, "w") as f:
            json.dump({"batch": B, "size": S, "ms_per_step": ms_step, "kernels": agg}, f, indent=1)

    # (4) the other BASELINE configurations and the baselines, outside every timed region above
    ts.release_graph()          # also required before the NCCL communicator can be torn down
    ts.engine.plans.clear()
    del ts, bufs, run_step
    torch.cuda.empty_cache()
    gpu_baseline = None
    extra["bucket_mb"], extra["nccl_max_ctas"] = args.bucket_mb, args.nccl_max_ctas
    if not args.no_extras:
        try:
            extra["configs4_vnet_training"] = bench_legs.vnet_leg(dev, world, rank, local, batch=args.vnet_batch)
        except Exception as e:   # a secondary leg must not take the headline down
            extra["configs4_vnet_training"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        if world == 1:
            try:
                extra["configs3_inference"] = bench_legs.inference_leg(dev)
            except Exception as e:
                extra["configs3_inference"] = {"error": f"{type(e).__name__}: {e}"[:300]}
            try:
                gpu_baseline = bench_legs.gpu_baseline_leg(dev, batch=B, size=S)
                if "value" in gpu_baseline:
                    gpu_baseline["b200_path_over_stock_torch"] = value / gpu_baseline["value"]
            except Exception as e:
                gpu_baseline = {"error": f"{type(e).__name__}: {e}"[:300]}
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        times, kind, note = bench_legs.cpu_reference_times(4, S, 3, 1, threads)
        v = 4 / (sum(times) / len(times))
        cpu_baseline = {"value": v, "unit": "images/s", "cores": threads, "kind": kind,
                        "sample": f"3 timed steps (1 warm-up) on 4 images at {S}x{S} (BASELINE configs[0]): {note}"}

    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(B, S, world),
                "clocks": clocks, "e2e": {"value": e2e_value, "unit": "images/s",
                                          "h2d_bytes_per_step": x_pin.numel() * 4 + t_pin.numel() * 4,
                                          "d2h_bytes_per_step": 32},
                "gpu_launches": launches, "cuda_graph": bool(graphed), "roofline": roofline, "cpu_baseline": cpu_baseline,
                "gpu_baseline": gpu_baseline, "extra": extra}
        if capture_error:
            line["cuda_graph_error"] = capture_error
        print(json.dumps(line), flush=True)
    if world > 1:
        shutdown_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100, help="timed steps (default: >= 2 s of device time)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[2]/[3]/[4] legs and the stock-torch baseline")
    ap.add_argument("--bucket-mb", type=float, default=32.0, help="gradient bucket size of the all-reduce")
    ap.add_argument("--nccl-max-ctas", type=int, default=0, help="ncclConfig max_ctas for the gradient all-reduces (0: NCCL default)")
    ap.add_argument("--vnet-batch", type=int, default=16, help="per-GPU batch of the V-Net leg (configs[4])")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of one CUDA graph per step")
    ap.add_argument("--profile-out", default="", help="write the per-kernel roofline table (JSON) here")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch ourselves one rank per GPU (the driver normally does this)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
