"""The drop-in boundary, exercised by the reference's OWN code: the unmodified utils/trainer.py of the
reference is imported with `models.model` / `models.loss` / `models.vnet` resolving to this repo's modules, and its
Trainer is constructed around the B200 UNet (AdamW over the parameters, the four zero-argument losses, GradScaler,
scheduler, DataParallel branch not taken on CPU). Running an epoch needs a GPU (trainer.py:60 calls .cuda()) and the
reference tree does not travel to the GPU box, so this test stops at construction plus what the trainer touches on the
model. Skipped where /root/reference is absent. Runs in a subprocess: it aliases top-level module names."""
import os
import subprocess
import sys
import textwrap

import pytest

REFERENCE = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import sys, types, tempfile, logging
    sys.path.insert(0, %(root)r)
    import torch
    import b200seg
    from b200seg.models import model as b_model, loss as b_loss, vnet as b_vnet
    import b200seg.models as b_models
    # plotting / image libraries the trainer imports but construction does not use (not installed here)
    for name in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.measure", "pytz", "seaborn"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    # the reference's `from models.loss import ...` must find THIS repo's modules
    sys.modules["models"] = b_models
    sys.modules["models.model"], sys.modules["models.loss"], sys.modules["models.vnet"] = b_model, b_loss, b_vnet
    sys.path.insert(0, %(ref)r)
    import utils.trainer as ref_trainer                      # the reference's file, unmodified
    assert ref_trainer.__file__.startswith(%(ref)r), ref_trainer.__file__
    assert ref_trainer.DiceLoss is b_loss.DiceLoss and ref_trainer.CompositeLoss is b_loss.CompositeLoss

    class Cfg: pass
    cfg = Cfg()
    cfg.device = torch.device("cpu"); cfg.use_data_parallel = True; cfg.use_amp_autocast = True; cfg.lr = 1e-4
    cfg.early_stop_patience = 5; cfg.result_dir = tempfile.mkdtemp(); cfg.model_dir = cfg.result_dir
    torch.manual_seed(42)
    net = b_model.UNet(in_channels=1, out_channels=1)
    tr = ref_trainer.Trainer(cfg, (None, None, None), logging.getLogger("t"), net)
    assert tr.model is net
    assert isinstance(tr.criterion_dice, b_loss.DiceLoss) and isinstance(tr.criterion_focal, b_loss.FocalTverskyLoss)
    assert isinstance(tr.criterion_boundary, b_loss.BoundaryLoss) and isinstance(tr.criterion, b_loss.CompositeLoss)
    n_opt = sum(p.numel() for g in tr.optimizer.param_groups for p in g["params"])
    assert n_opt == 31042369, n_opt                         # reference test.py's parameter count
    assert all(p.is_leaf and p.dtype == torch.float32 for g in tr.optimizer.param_groups for p in g["params"])
    tr.model.train(); assert net.encoder1[2].training
    tr.model.eval(); assert not net.encoder1[2].training
    sd = tr.model.state_dict()
    assert len(sd) == 136 and "middle.1.3.weight" in sd and "final.1.bias" in sd
    tr.model.load_state_dict(sd)
    tr.scheduler.step()
    # the V-Net variant constructs the way reference test.py:10 does
    v = b_vnet.ImprovedVNet(in_channels=1, num_classes=1)
    assert sum(p.numel() for p in v.parameters() if p.requires_grad) == 160435681
    print("TRAINER_OK")
''')


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "utils")), reason="reference tree not present")
def test_unmodified_reference_trainer_constructs_around_the_drop_in_modules():
    out = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "ref": REFERENCE}], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0 and "TRAINER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
