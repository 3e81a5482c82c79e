"""Secondary legs of bench.py (each returns a dict that goes under the JSON line's "extra" / "gpu_baseline" keys):

  inference_leg     BASELINE configs[3]: UNet eval forward + fused threshold masks, batch 256 at 1x512x512, 1 GPU
  vnet_leg          BASELINE configs[4]: ImprovedVNet (models/vnet.py) training at 1x512x512, N GPUs (DDP buckets)
  gpu_baseline_leg  SURVEY §8(d) "kernel to beat": the reference's OWN modules (baseline/_ref, unmodified) on stock
                    PyTorch eager / cuDNN on the same B200 — bf16 autocast + channels_last, same batch, same step
  cpu_reference_leg the reference's own modules on the host cores, BASELINE configs[0] exactly (fp32, B=4 @256^2,
                    forward + BCE + Dice + backward, no optimiser step)

Everything that touches the reference goes through tools/ref_env.py (a loader; measurement infrastructure); the product
package is never routed through it, and oracle/ is imported only by the CPU fallback of cpu_reference_times.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FWD_FLOP_256 = 96.184e9        # SURVEY §8d, per image at 1x256x256
TRAIN_FLOP_256 = 288.476e9
VNET_TRAIN_FLOP_512 = 3894e9


def _events():
    import torch
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def inference_leg(dev, batch=256, chunk=64, size=512, steps=3):
    import torch
    import b200seg  # noqa: F401
    from b200seg import _lib
    from b200seg.models.model import UNet
    from b200seg.synth import synth_batch
    torch.manual_seed(42)
    net = UNet().to(dev).eval()
    with torch.no_grad():   # non-degenerate running statistics (a fresh net has mean 0 / var 1)
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.uniform_(0.0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
    x_cpu, _ = synth_batch(chunk, size, size, seed=1234)
    x_pin = x_cpu.pin_memory()
    x_dev = x_pin.to(dev)
    x_in = torch.empty_like(x_dev)
    mask_host = torch.empty((chunk, 1, size, size), dtype=torch.uint8).pin_memory()
    chunks = batch // chunk

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
        e0, e1 = _events()
        e0.record()
        for _ in range(steps):
            run(mode)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res["e2e" if mode else "resident"] = {"ms_per_batch": ms, "frames_per_s": chunks * chunk / (ms * 1e-3),
                                              "launches_per_batch": (_lib.launch_count() - l0) / steps}
    flops = FWD_FLOP_256 * (size / 256) ** 2
    res["algorithmic_tflops"] = chunks * chunk * flops / (res["resident"]["ms_per_batch"] * 1e-3) / 1e12
    res.update({"workload": f"UNet eval forward + uint8 masks, batch {chunks * chunk} at 1x{size}x{size} in chunks of "
                            f"{chunk} (BASELINE configs[3])", "steps": steps,
                "e2e_note": "pinned H2D of every chunk + D2H of its uint8 masks inside the timed region"})
    del net, x_dev, x_in
    torch.cuda.empty_cache()
    return res


def vnet_leg(dev, world, rank, local, batch=16, size=512, steps=5, warmup=3):
    import torch
    import torch.distributed as dist
    import b200seg  # noqa: F401
    from b200seg import _lib
    from b200seg.models.vnet import ImprovedVNet
    from b200seg.models.loss import BCEDiceLoss
    from b200seg.synth import synth_batch
    torch.manual_seed(42)
    net = ImprovedVNet().to(dev).train()
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
    opt = torch.optim.AdamW(net.parameters(), lr=1e-5, fused=True)
    crit = BCEDiceLoss()
    x, t = synth_batch(batch, size, size, seed=1234 + rank)
    x, t = x.to(dev), t.to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), t)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _lib.launch_count()
    e0, e1 = _events()
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms) / steps
    flops = VNET_TRAIN_FLOP_512 * (size / 512) ** 2
    res = {"workload": f"ImprovedVNet (models/vnet.py) training, batch {batch} at 1x{size}x{size} per GPU, dropout 0.05, "
                       f"fused BCE+Dice, torch AdamW (fused), DDP gradient buckets (BASELINE configs[4])",
           "n_gpus": world, "ms_per_step": ms, "images_per_s": world * batch / (ms * 1e-3),
           "algorithmic_tflops_per_gpu": batch * flops / (ms * 1e-3) / 1e12, "loss": float(loss), "steps": steps,
           "libb2s_launches_per_step": (_lib.launch_count() - l0) / steps,
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    del model, net, opt, x, t
    torch.cuda.empty_cache()
    return res


def _reference_unet():
    """(UNet class, DiceLoss class, kind): the reference's own modules when staged, else None."""
    from tools import ref_env
    if ref_env.ref_root() is None:
        return None
    return ref_env.load_ref_module("models/model.py").UNet, ref_env.load_ref_module("models/loss.py").DiceLoss


def gpu_baseline_leg(dev, batch=64, size=256, steps=10, warmup=4):
    """Stock PyTorch on the same GPU: reference UNet + nn.BCEWithLogitsLoss + reference DiceLoss + torch AdamW (fused),
    autocast(bf16) + channels_last (the fastest stock variant measured in round 1)."""
    import torch
    from b200seg.synth import synth_batch
    mods = _reference_unet()
    if mods is None:
        return {"unavailable": "reference modules not staged under baseline/_ref (tools/stage_reference.py)"}
    UNet, DiceLoss = mods
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(42)
    m = UNet().to(dev).train().to(memory_format=torch.channels_last)
    x, t = synth_batch(batch, size, size)
    x, t = x.to(dev).contiguous(memory_format=torch.channels_last), t.to(dev)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-5, fused=True)
    bce, dice = torch.nn.BCEWithLogitsLoss(), DiceLoss()

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            lg = m(x)
            loss = bce(lg, t) + dice(lg, t)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = _events()
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res = {"what": "the reference's own models/model.py UNet + models/loss.py DiceLoss + nn.BCEWithLogitsLoss, UNMODIFIED, on "
                   "stock PyTorch eager (cuDNN/cuBLAS): autocast(bf16) + channels_last + torch AdamW(fused)",
           "kind": "reference", "torch": torch.__version__, "batch": batch, "image": f"1x{size}x{size}", "steps": steps,
           "ms_per_step": ms, "value": batch / (ms * 1e-3), "unit": "images/s",
           "algorithmic_tflops": batch * TRAIN_FLOP_256 * (size / 256) ** 2 / (ms * 1e-3) / 1e12, "loss": float(loss)}
    del m, opt, x, t
    torch.cuda.empty_cache()
    return res


def cpu_reference_times(batch, size, steps, warmup, threads):
    """Seconds per step of the reference's CPU path. Returns (times, kind, note).

    kind "reference": the reference's own nn.Modules (baseline/_ref/models/model.py, models/loss.py) — forward + BCE +
    Dice + backward in fp32, no optimiser step: BASELINE configs[0] / SURVEY §8(d) exactly.
    kind "port" (only when the reference tree is not staged): oracle/unet_torch_ref.py, the same torch CPU ops."""
    import torch
    from b200seg.synth import synth_batch
    torch.set_num_threads(threads)
    x, t = synth_batch(batch, size, size, seed=1234)
    mods = _reference_unet()
    if mods is not None:
        UNet, DiceLoss = mods
        torch.manual_seed(42)
        net = UNet().train()
        bce, dice = torch.nn.BCEWithLogitsLoss(), DiceLoss()

        def step():
            net.zero_grad(set_to_none=True)
            lg = net(x)
            (bce(lg, t) + dice(lg, t)).backward()
        kind = "reference"
        note = ("the reference's own models/model.py UNet + models/loss.py DiceLoss + nn.BCEWithLogitsLoss (unmodified, "
                "baseline/_ref), fp32, train mode, forward + loss + backward, no optimiser step")
    else:
        from oracle import unet_torch_ref as T
        import b200seg  # noqa: F401
        from b200seg.models.model import UNet as DropIn
        torch.manual_seed(42)
        P = T.make_params(DropIn().state_dict())

        def step():
            T.train_step(P, x, t)
        kind = "port"
        note = "oracle/unet_torch_ref.py (the torch CPU ops the reference dispatches to); reference tree not staged"
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return times, kind, note
