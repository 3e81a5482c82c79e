#!/usr/bin/env python
"""Runs the reference's UNMODIFIED ``utils/trainer.py`` (``Trainer.train_one_epoch`` / ``validate`` / ``test``,
reference utils/trainer.py:47-120,122-172,207-260) on a GPU with ``models.model`` / ``models.loss`` / ``models.vnet`` /
``models.mod`` resolving to THIS repo's drop-in modules, over a synthetic DataLoader. Prints one JSON line.

    python tests/run_reference_trainer.py [--amp 0|1] [--data-parallel 0|1] [--model UNet|ImprovedVNet|ResUNet]
                                          [--epochs 3] [--size 64] [--samples 32] [--batch 8] [--mixup 0|1]

Own process (it aliases top-level module names); driven by tests/test_trainer_gpu.py. The reference tree comes from
baseline/_ref/ (tools/stage_reference.py) or /root/reference; nothing of it is modified.
"""
import argparse
import json
import logging
import os
import re
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("B2S_HANG_DUMP"):          # debugging aid: dump every thread's stack if the run is still alive after N s
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ["B2S_HANG_DUMP"]), exit=True)


def install_plot_stubs():
    """matplotlib / skimage / pytz / seaborn are not installed; Trainer.test()'s plotting tail (trainer.py:268-299)
    needs callable stand-ins, the rest only needs the imports to succeed."""
    from tools import ref_env
    ref_env.stub_missing_modules()
    plt = sys.modules["matplotlib.pyplot"]
    if not hasattr(plt, "subplots"):
        import numpy as np

        class _Ax:
            def imshow(self, *a, **k): pass
            def plot(self, *a, **k): pass
            def axis(self, *a, **k): pass

        def subplots(r, c, **k):
            ax = np.empty((r, c), dtype=object)
            for i in range(r):
                for j in range(c):
                    ax[i, j] = _Ax()
            return object(), ax
        plt.subplots = subplots
        plt.tight_layout = lambda *a, **k: None
        plt.savefig = lambda path, *a, **k: open(path, "wb").close()
        plt.close = lambda *a, **k: None
    meas = sys.modules["skimage.measure"]
    if not hasattr(meas, "find_contours"):
        meas.find_contours = lambda *a, **k: []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--amp", type=int, default=1)
    ap.add_argument("--data-parallel", type=int, default=0)
    ap.add_argument("--model", default="UNet")
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--samples", type=int, default=32)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--mixup", type=int, default=0)
    ap.add_argument("--lr", type=float, default=2e-3)
    args = ap.parse_args()

    import numpy as np
    import torch
    import b200seg  # noqa: F401
    from b200seg.models import model as b_model, loss as b_loss, vnet as b_vnet, mod as b_mod
    import b200seg.models as b_models
    from b200seg.models.metrics import SegMetrics
    from b200seg.synth import synth_batch
    from b200seg import _lib
    from tools import ref_env

    install_plot_stubs()
    ref = ref_env.ref_root()
    assert ref is not None, "reference tree not found: run tools/stage_reference.py in the build container"
    # `from models.loss import ...` inside the reference trainer must find THIS repo's modules
    sys.modules["models"] = b_models
    sys.modules["models.model"], sys.modules["models.loss"] = b_model, b_loss
    sys.modules["models.vnet"], sys.modules["models.mod"] = b_vnet, b_mod
    sys.path.insert(0, ref)
    import utils.trainer as ref_trainer                      # the reference's file, unmodified
    assert os.path.abspath(ref_trainer.__file__).startswith(os.path.abspath(ref)), ref_trainer.__file__
    assert ref_trainer.DiceLoss is b_loss.DiceLoss

    dev = torch.device("cuda")
    n_gpu = torch.cuda.device_count()
    use_dp = bool(args.data_parallel) and n_gpu > 1

    class Cfg:
        pass
    cfg = Cfg()
    cfg.device = dev
    cfg.use_data_parallel = use_dp
    cfg.use_amp_autocast = bool(args.amp)
    cfg.lr = args.lr
    cfg.early_stop_patience = 50
    cfg.result_dir = tempfile.mkdtemp(prefix="b2s_trainer_")
    cfg.model_dir = cfg.result_dir
    cfg.model_type = args.model
    cfg.epochs = args.epochs
    cfg.use_mixup, cfg.mixup_prob, cfg.mixup_alpha = bool(args.mixup), 0.5, 0.4
    cfg.bce_ratio, cfg.dice_ratio, cfg.focal_ratio, cfg.boundary_ratio = 1.0, 1.0, 0.5, 0.0

    S = args.size
    x, t = synth_batch(args.samples, S, S, seed=4321)
    ds = torch.utils.data.TensorDataset(x, t)
    mk = lambda shuffle: torch.utils.data.DataLoader(ds, batch_size=args.batch, shuffle=shuffle, drop_last=True)
    loaders = (mk(True), mk(False), mk(False))

    records = []

    class Grab(logging.Handler):
        def emit(self, rec):
            records.append(rec.getMessage())
    logger = logging.getLogger("b2s_trainer")
    logger.setLevel(logging.INFO)
    logger.addHandler(Grab())

    torch.manual_seed(42)
    np.random.seed(0)
    import random
    random.seed(0)
    if args.model == "UNet":
        net = b_model.UNet(in_channels=1, out_channels=1)
    elif args.model == "ImprovedVNet":
        net = b_vnet.ImprovedVNet(in_channels=1, num_classes=1)
    else:
        net = b_mod.ResUNet(in_channels=1, out_channels=1)

    # without DataParallel requested, hide the extra GPUs from the trainer's own device_count() test (trainer.py:28)
    if not use_dp:
        cfg.use_data_parallel = False
    launches0 = _lib.launch_count()
    tr = ref_trainer.Trainer(cfg, loaders, logger, net)
    wrapped = isinstance(tr.model, torch.nn.DataParallel)
    assert wrapped == use_dp

    train_losses, val = [], []
    for epoch in range(args.epochs):
        tr.train_one_epoch(epoch)
        msg = [m for m in records if m.startswith("Train Epoch")][-1]
        train_losses.append(float(re.search(r"Avg Loss: ([0-9.eE+-]+)", msg).group(1)))
        vloss, viou = tr.validate(epoch)
        val.append((float(vloss), float(viou)))
        tr.scheduler.step()
    launches = _lib.launch_count() - launches0

    # validate()'s IoU against the on-device SegMetrics of the same eval-mode predictions
    core = tr.model.module if wrapped else tr.model
    core.eval()
    m = SegMetrics(dev)
    with torch.no_grad():
        for xb, tb in loaders[1]:
            m.update(core(xb.to(dev)), tb.to(dev))
    ours = m.compute()

    # Trainer.test(): TP/FP/FN/TN line (trainer.py:236-250) against the same counters
    records.clear()
    tr.test()
    tmsg = [mm for mm in records if "Test Metrics" in mm][-1]
    tp, fp, fn, tn = (int(re.search(rf"{k}=(\d+)", tmsg).group(1)) for k in ("TP", "FP", "FN", "TN"))
    t_iou = float(re.search(r"IoU=([0-9.]+)", tmsg).group(1))

    # checkpoint round trip the way Trainer.train() saves it (trainer.py:188-190,200-202; main.py:142)
    path = os.path.join(cfg.model_dir, f"{cfg.model_type}_last.pth")
    torch.save(core.state_dict(), path)
    fresh = type(core)()
    missing, unexpected = fresh.load_state_dict(torch.load(path, weights_only=True), strict=True)
    fresh = fresh.to(dev).eval()
    with torch.no_grad():
        xb = x[: args.batch].to(dev)
        same = bool(torch.equal(fresh(xb), core(xb)))

    grads_ok = all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in core.parameters())
    out = {"model": args.model, "amp": bool(args.amp), "data_parallel": wrapped, "gpus": n_gpu,
           "train_losses": train_losses, "val": val, "val_iou_segmetrics": ours["iou"],
           "test_counts": {"tp": tp, "fp": fp, "fn": fn, "tn": tn}, "test_iou": t_iou,
           "segmetrics_counts": {k.lower(): int(ours[k]) for k in ("TP", "FP", "FN", "TN")},
           "segmetrics": {k: v for k, v in ours.items() if isinstance(v, (int, float))},
           "checkpoint_roundtrip_bit_identical": same, "missing": list(missing), "unexpected": list(unexpected),
           "grads_finite": grads_ok, "scaler_enabled": bool(tr.scaler.is_enabled()),
           "scaler_scale": float(tr.scaler.get_scale()) if tr.scaler.is_enabled() else None,
           "libb2s_launches": int(launches), "steps": args.epochs * (args.samples // args.batch)}
    print("TRAINER_JSON " + json.dumps(out))


if __name__ == "__main__":
    main()
