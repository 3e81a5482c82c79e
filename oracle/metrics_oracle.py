"""ORACLE (test infrastructure, not product code) — numpy restatement of the reference's segmentation metrics.

Only tests/ may import this module. Follows utils/utils.py:225-251 (calculate_iou / calculate_acc /
calculate_precision_recall_f1) applied to preds = sigmoid(logits) > 0.5 (utils/trainer.py:101,152) and the float masks.
Pinned against the reference's own functions by oracle/make_golden_metrics.py -> tests/golden/metrics_golden.pt."""
import numpy as np


def metrics(preds_bool, targets_float):
    pred_i, targ_i = preds_bool.astype(int), targets_float.astype(int)        # utils.py:234-235,241-242
    pred_b, targ_b = preds_bool.astype(bool), targets_float.astype(bool)      # utils.py:227-228
    iou = np.logical_and(pred_b, targ_b).sum() / np.logical_or(pred_b, targ_b).sum()
    acc = (pred_i == targ_i).sum() / pred_i.size
    TP = np.logical_and(pred_i == 1, targ_i == 1).sum()
    FP = np.logical_and(pred_i == 1, targ_i == 0).sum()
    FN = np.logical_and(pred_i == 0, targ_i == 1).sum()
    precision = TP / (TP + FP) if TP + FP > 0 else 0.0
    recall = TP / (TP + FN) if TP + FN > 0 else 0.0
    f1 = 2 * precision * recall / (precision + recall) if precision + recall > 0 else 0.0
    return {"acc": float(acc), "precision": float(precision), "recall": float(recall), "f1": float(f1), "iou": float(iou)}
