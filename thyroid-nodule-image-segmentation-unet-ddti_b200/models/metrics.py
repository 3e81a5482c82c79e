"""On-device segmentation metrics: the numbers the reference computes on the host from per-step device->host copies of
the thresholded predictions (utils/trainer.py:101-107,151-158,217-250 with utils/utils.py:225-251), accumulated in
seven int64 counters on the GPU instead. One D2H copy per ``compute()`` (per epoch), none per step."""
import torch

from .. import ops


class SegMetrics:
    """update(logits, targets) adds the confusion counts of ``sigmoid(logits) > 0.5``; compute() returns the reference's
    accuracy, precision, recall, F1 and IoU (same formulas, same int / bool target conversions)."""

    def __init__(self, device="cuda"):
        self.counters = torch.zeros(7, dtype=torch.int64, device=device)

    def reset(self):
        self.counters.zero_()

    @ops.on_device_of_input
    def update(self, logits, targets):
        if not logits.is_cuda:
            raise RuntimeError("b200seg metrics run on CUDA (sm_100a) only; there is no CPU fallback")
        ops.seg_metrics(logits, targets, self.counters)

    def compute(self):
        TP, FP, FN, TN, inter, union, n = (int(v) for v in self.counters.tolist())
        # utils/utils.py:232-251 (train / validate epochs)
        acc = (TP + TN) / n if n else 0.0            # calculate_acc counts pred == target after astype(int)
        precision = TP / (TP + FP) if TP + FP > 0 else 0.0
        recall = TP / (TP + FN) if TP + FN > 0 else 0.0
        f1 = 2 * precision * recall / (precision + recall) if precision + recall > 0 else 0.0
        iou = inter / union if union else float("nan")
        return {"acc": acc, "precision": precision, "recall": recall, "f1": f1, "iou": iou,
                "TP": TP, "FP": FP, "FN": FN, "TN": TN, "n": n}
