"""Pins oracle/unet_oracle.py against golden vectors generated from the UNMODIFIED reference
(oracle/make_golden.py; reference models/model.py, models/loss.py, torch ops). CPU only."""
import hashlib

import torch

from oracle import unet_oracle as O

D = torch.float64


def close(a, b, tol=1e-10):
    a, b = a.double(), b.double()
    return float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


def test_conv3x3_matches_torch(ops_golden):
    g = ops_golden["conv3x3"]
    assert close(O.conv3x3(g["x"], g["w"], g["b"]), g["z"])
    dx, dw, db = O.conv3x3_bwd(g["x"], g["w"], g["dz"])
    assert close(dx, g["dx"]) and close(dw, g["dw"]) and close(db, g["db"])


def test_conv_transpose_matches_torch(ops_golden):
    g = ops_golden["convt"]
    assert close(O.conv_transpose2x2(g["x"], g["w"], g["b"]), g["y"])
    dx, dw, db = O.conv_transpose2x2_bwd(g["x"], g["w"], g["dy"])
    assert close(dx, g["dx"]) and close(dw, g["dw"]) and close(db, g["db"])


def test_batchnorm_matches_torch(ops_golden):
    g = ops_golden["bn"]
    mean, var = O.batchnorm_stats(g["r"])
    scale, shift, invstd = O.batchnorm_affine(mean, var, g["gamma"], g["beta"])
    y = g["r"] * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    assert close(y, g["y"])
    dr, dgamma, dbeta = O.batchnorm_bwd(g["dy"], g["r"], mean, invstd, g["gamma"])
    assert close(dr, g["dr"]) and close(dgamma, g["dgamma"]) and close(dbeta, g["dbeta"])
    n = g["r"].numel() // g["r"].shape[1]
    assert close(0.1 * mean, g["running_mean"])
    assert close(0.9 + 0.1 * var * n / (n - 1), g["running_var"])
    assert int(g["nbt"]) == 1


def test_maxpool_first_max_tiebreak(ops_golden):
    g = ops_golden["pool"]
    p, arg = O.maxpool2x2(g["y"])
    assert torch.equal(p, g["p"])
    assert torch.equal(O.maxpool2x2_bwd(g["dp"], arg, g["y"].shape), g["dy"])
    # all-equal window -> gradient to position 0
    y = torch.ones((1, 1, 2, 2), dtype=D)
    _, a = O.maxpool2x2(y)
    assert int(a) == 0


def test_losses_match_reference(ops_golden):
    g = ops_golden["loss"]
    r = O.seg_loss(g["logits"], g["targets"], w_bce=g["w"][0], w_dice=g["w"][1], w_ft=g["w"][2])
    # Dice tolerance: the reference sums its .float()-cast targets in fp32 (models/loss.py:19,22)
    assert close(r["bce"], g["bce"]) and close(r["dice"], g["dice"], 1e-7) and close(r["ft"], g["ft"])
    assert close(r["total"], g["total"], 1e-7)
    assert close(r["dlogits"], g["dlogits"], 1e-9)


def test_threshold_semantics(ops_golden):
    g = ops_golden["threshold"]
    # fp32 sigmoid(x) > 0.5 flips somewhere in (5.96e-8, 1.19e-7] depending on the exp implementation's last ulp
    # (torch's own CPU sigmoid vs exp kernels disagree there): that interval is the declared guard band.
    outside = ~((g["logits"] > 0) & (g["logits"] <= 1.2e-7))
    assert torch.equal(O.threshold_mask(g["logits"])[outside], g["mask"][outside])
    assert not bool(O.threshold_mask(torch.zeros(1))[0])  # logit == 0 -> False


def test_adamw_matches_torch(ops_golden):
    g = ops_golden["adamw"]
    p = g["p0"].clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for s in range(3):
        O.adamw_step(p, g["grads"][s], m, v, g["lr"], step=s + 1)
    assert close(p, g["p3"], 1e-12)


def test_param_layout_known_answers(unet_golden, ref_params):
    shapes = O.unet_param_shapes()
    assert list(shapes.keys()) == unet_golden["state_dict_keys"]
    assert {k: tuple(v) for k, v in shapes.items()} == unet_golden["state_dict_shapes"]
    n = sum(int(torch.tensor(s).prod()) for k, s in shapes.items()
            if not any(t in k for t in ("running", "num_batches")))
    assert n == unet_golden["param_count"] == 31042369
    assert len(shapes) == 136
    for k, v in ref_params.items():
        assert hashlib.sha1(v.contiguous().numpy().tobytes()).hexdigest() == unet_golden["init_digest"][k]["sha1"], k


def test_unet_forward_backward_matches_reference(unet_golden, ref_params):
    """Whole-net oracle (explicit forward + hand-derived backward, fp64) vs the reference's fp32 autograd run."""
    A = unet_golden["A"]
    P = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in ref_params.items()}
    cache = {}
    logits = O.unet_forward(P, A["x"].double(), train=True, cache=cache)
    assert float((logits - A["logits"].double()).abs().max()) < 2e-5
    L = O.seg_loss(logits, A["t"].double(), w_bce=1.0, w_dice=1.0)
    assert abs(float(L["bce"]) - A["bce"]) < 1e-6 and abs(float(L["dice"]) - A["dice"]) < 1e-6
    G = O.unet_backward(P, cache, L["dlogits"])
    worst = 0.0
    for k, ref in A["grads"].items():
        got = G[k].flatten()[ref["idx"]]
        denom = max(ref["norm"] / (G[k].numel() ** 0.5), 1e-12)
        err = float((got - ref["vals"].double()).abs().max()) / denom
        worst = max(worst, err)
        assert abs(float(G[k].norm()) - ref["norm"]) <= 2e-3 * ref["norm"] + 1e-9, k
        assert err < 5e-2, (k, err)
    # running statistics after one train step
    O.running_stats_update(P, cache)
    for k, v in A["running"].items():
        if v.is_floating_point():
            assert float((P[k] - v.double()).abs().max()) < 1e-5, k
        else:
            assert int(P[k]) == int(v) == 1
    # eval mode with the updated running stats: logits and thresholded mask
    le = O.unet_forward(P, A["x"].double(), train=False)
    assert float((le - A["eval_logits"].double()).abs().max()) < 2e-5
    far = A["eval_logits"].abs() > 1e-4
    assert torch.equal(O.threshold_mask(le.float())[far], A["eval_mask"][far])


def test_unet_known_answer_losses(unet_golden):
    """BASELINE.md / SURVEY.md §8c known answers of the seeded probe."""
    B = unet_golden["B"]
    assert abs(B["bce"] - 0.746483) < 2e-6 and abs(B["dice"] - 0.620441) < 2e-6


def test_bf16_emulation_close_to_exact(unet_golden, ref_params):
    """The q=bf16 oracle (the precision model of the CUDA path) stays within the reference's own bf16 noise."""
    A = unet_golden["A"]
    P = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in ref_params.items()}
    lq = O.unet_forward(P, A["x"].double(), train=True, q=O.bf16_round)
    assert float((lq - A["logits"].double()).abs().mean()) < 2.5e-2


def test_rejects_bad_spatial_size(ref_params):
    import pytest
    with pytest.raises(RuntimeError):
        O.unet_forward(ref_params, torch.zeros(1, 1, 40, 40))


def test_functional_torch_port_matches_reference(unet_golden, ref_params):
    """oracle/unet_torch_ref.py (the timed CPU baseline) reproduces the reference module's fp32 numbers."""
    from oracle import unet_torch_ref as T
    A = unet_golden["A"]
    P = T.make_params(ref_params)
    loss, logits, grads = T.train_step(P, A["x"], A["t"])
    assert float((logits - A["logits"]).abs().max()) < 1e-6
    assert abs(float(loss) - A["loss"]) < 1e-6
    names = [k for k, v in P.items() if v.requires_grad]
    for k, g in zip(names, grads):
        ref = A["grads"][k]
        assert abs(float(g.double().norm()) - ref["norm"]) <= 1e-4 * ref["norm"] + 1e-9, k


def test_metrics_oracle_matches_reference_functions():
    """oracle/metrics_oracle.py against goldens computed with the reference's own utils/utils.py:225-251"""
    import os
    from oracle import metrics_oracle as MO
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_golden.pt"), weights_only=False)
    for name, c in gold.items():
        got = MO.metrics((torch.sigmoid(c["logits"]) > 0.5).numpy(), c["targets"].numpy())
        for k in ("acc", "precision", "recall", "f1", "iou"):
            assert abs(got[k] - c[k]) < 1e-12, (name, k)


def test_workload_generator_matches_the_oracles_copy():
    """bench.py and the tools draw their synthetic frames from the package (b200seg.synth), the parity tests from the
    oracle: both must be the same function of (shape, seed)"""
    import b200seg  # noqa: F401
    from b200seg.synth import synth_batch
    for (B, H, W, seed) in [(2, 32, 32, 1234), (3, 48, 80, 7)]:
        xa, ta = synth_batch(B, H, W, seed=seed)
        xb, tb = O.synth_batch(B, H, W, seed=seed)
        assert torch.equal(xa, xb) and torch.equal(ta, tb)
