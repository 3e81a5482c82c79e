"""Generates tests/golden/metrics_golden.pt with the reference's own metric functions (utils/utils.py:225-251, imported
from /root/reference with its plotting / timezone dependencies stubbed) — run in the build container only."""
import os, sys, types
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle.make_golden import REF, OUT  # noqa: E402

for name in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.measure", "pytz", "seaborn", "sklearn", "sklearn.metrics"):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            m.__dict__.setdefault("confusion_matrix", None)
            sys.modules[name] = m
sys.path.insert(0, REF)
from utils.utils import calculate_iou, calculate_acc, calculate_precision_recall_f1  # noqa: E402

g = torch.Generator().manual_seed(3)
cases = {}
for name, soft in (("binary", False), ("soft", True)):
    logits = torch.randn((3, 1, 40, 56), generator=g) * 2
    t = (torch.rand((3, 1, 40, 56), generator=g) > 0.7).float()
    if soft:   # JPEG / bilinear / mixup style targets: values in (0,1) truncate to 0 for astype(int)
        t = torch.where(torch.rand(t.shape, generator=g) > 0.8, torch.rand(t.shape, generator=g), t)
    preds = (torch.sigmoid(logits) > 0.5).numpy()
    tn = t.numpy()
    p, r, f1 = calculate_precision_recall_f1(preds, tn)
    cases[name] = dict(logits=logits, targets=t, acc=float(calculate_acc(preds, tn)), precision=float(p), recall=float(r),
                       f1=float(f1), iou=float(calculate_iou(preds, tn)))
os.makedirs(OUT, exist_ok=True)
torch.save(cases, os.path.join(OUT, "metrics_golden.pt"))
print({k: {m: round(v[m], 5) for m in ("acc", "precision", "recall", "f1", "iou")} for k, v in cases.items()})
