"""The reference's UNMODIFIED trainer running on the B200 path (north_star: "drops into
utils/trainer.py and test.py unchanged"): ``Trainer.train_one_epoch`` (autocast + GradScaler + four losses + stock AdamW,
reference utils/trainer.py:47-120), ``validate`` (:122-172) and ``test`` (:207-260) are executed from the reference's own
file — staged byte-for-byte under the git-ignored baseline/_ref/ by tools/stage_reference.py — over a synthetic
DataLoader, with ``models.*`` resolving to this repo's drop-in modules. Each case is its own process
(tests/run_reference_trainer.py aliases top-level module names)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_reference():
    sys.path.insert(0, ROOT)
    from tools import ref_env
    return ref_env.ref_root() is not None


def _run(*flags, timeout=900):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_trainer.py"), *flags],
                         capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith("TRAINER_JSON ")]
    assert out.returncode == 0 and lines, out.stdout[-2000:] + out.stderr[-6000:]
    res = json.loads(lines[-1][len("TRAINER_JSON "):])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = "_".join(f.strip("-") for f in flags) or "default"
    with open(os.path.join(ROOT, "gpurun_out", f"trainer_{tag}.json"), "w") as f:
        json.dump(res, f, indent=1)
    return res


def _check(res, steps_min_launches=100):
    L = res["train_losses"]
    assert all(l == l and l < 1e3 for l in L), L
    assert L[-1] < L[0], f"training loss did not decrease over the epochs: {L}"
    assert res["grads_finite"]
    # validate(): the IoU the trainer computed on the host from thresholded predictions == on-device SegMetrics
    assert abs(res["val"][-1][1] - res["val_iou_segmetrics"]) < 1e-9, (res["val"], res["val_iou_segmetrics"])
    # test(): the reference's TP/FP/FN/TN == the on-device counters, pixel for pixel
    assert res["test_counts"] == res["segmetrics_counts"], (res["test_counts"], res["segmetrics_counts"])
    assert res["checkpoint_roundtrip_bit_identical"] and not res["missing"] and not res["unexpected"]
    # the run went through libb2s (every forward/backward of the model is > 100 of its launches)
    assert res["libb2s_launches"] > steps_min_launches * res["steps"], res["libb2s_launches"]


@pytest.mark.skipif(not _have_reference(), reason="baseline/_ref not staged (tools/stage_reference.py)")
@pytest.mark.parametrize("amp", [1, 0])
def test_unmodified_trainer_epochs_unet(amp):
    res = _run("--amp", str(amp), "--data-parallel", "0")
    assert res["scaler_enabled"] == bool(amp) and not res["data_parallel"]
    _check(res)


@pytest.mark.skipif(not _have_reference(), reason="baseline/_ref not staged (tools/stage_reference.py)")
def test_unmodified_trainer_mixup_soft_targets():
    """mixup (trainer.py:62-78) produces soft masks in [0,1]: the fused losses and metrics take them"""
    res = _run("--amp", "1", "--mixup", "1", "--epochs", "2")
    assert all(l == l for l in res["train_losses"]) and res["grads_finite"]
    assert res["test_counts"] == res["segmetrics_counts"]


@pytest.mark.skipif(not _have_reference(), reason="baseline/_ref not staged (tools/stage_reference.py)")
def test_unmodified_trainer_data_parallel():
    """nn.DataParallel (trainer.py:28-30) over all visible GPUs: replicas of the drop-in module run concurrently, one
    host thread per GPU, gradients reach the master parameters through torch's Broadcast.backward."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    res = _run("--amp", "1", "--data-parallel", "1")
    assert res["data_parallel"] and res["gpus"] >= 2
    _check(res, steps_min_launches=150)


@pytest.mark.skipif(not _have_reference(), reason="baseline/_ref not staged (tools/stage_reference.py)")
def test_unmodified_trainer_vnet_and_resunet():
    for model in ("ImprovedVNet", "ResUNet"):
        res = _run("--amp", "1", "--model", model, "--epochs", "2", "--samples", "16", "--batch", "4", "--size", "32",
                   "--lr", "1e-3")
        assert all(l == l for l in res["train_losses"]) and res["grads_finite"], res
        assert res["test_counts"] == res["segmetrics_counts"]
        assert res["checkpoint_roundtrip_bit_identical"]
