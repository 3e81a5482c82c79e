"""Parity AT THE BENCHMARKED SHAPES (256-pixel-wide rows select the row-halo conv / wgrad kernels, which the small
cases of test_unet_gpu.py never launch):

  * golden case B of tests/golden/unet_golden.pt — the UNMODIFIED reference on the BASELINE.md probe (B = 4 @ 256^2,
    seed-42 weights, rand / rand > 0.7 data): known-answer losses, 4 099 logit samples, 82 gradient norms;
  * the same case against the oracle (oracle/unet_oracle.py, q = bf16_round) evaluated in fp64 ON THE GPU — the same
    elementary-algebra restatement, only the device differs (the CPU needs minutes at this size);
  * backward-given-forward-state at 2 x 256^2 and 1 x 256 x 512: every one of the 82 gradients to bf16 rounding noise;
  * CUDA-graph step == host-launched step at 4 x 256^2.
"""
import json
import os

import pytest
import torch

from oracle import unet_oracle as O
from test_unet_gpu import _cache_from_plan, cosine, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def net(ref_params):
    import b200seg  # noqa: F401
    from b200seg.models.model import UNet
    m = UNet()
    m.load_state_dict(ref_params, strict=True)
    return m.to(DEV)


def _p64(ref_params, dev):
    return {k: (v.double().to(dev) if v.is_floating_point() else v.clone().to(dev)) for k, v in ref_params.items()}


def _dump(name, obj):
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", name), "w") as f:
        json.dump(obj, f, indent=1)


def test_golden_case_B_baseline_probe_256(net, unet_golden, ref_params):
    from b200seg.models.loss import BCEDiceLoss
    Bg = unet_golden["B"]
    g = torch.Generator().manual_seed(1234)
    x = torch.rand((4, 1, 256, 256), generator=g)
    t = (torch.rand((4, 1, 256, 256), generator=g) > 0.7).float()
    net.load_state_dict(ref_params, strict=True)
    net.train()
    net.zero_grad(set_to_none=True)
    logits = net(x.to(DEV))
    loss = BCEDiceLoss()(logits, t.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    loss = float(loss)

    # the oracle with the kernels' storage-precision model, fp64 on the device
    P = _p64(ref_params, DEV)
    cache = {}
    lq = O.unet_forward(P, x.double().to(DEV), train=True, q=O.bf16_round, cache=cache)
    Lq = O.seg_loss(lq, t.double().to(DEV))
    Gq = O.unet_backward(P, cache, Lq["dlogits"], q=O.bf16_round)
    del cache

    ref_loss = Bg["bce"] + Bg["dice"]
    lg = logits.detach().double()
    samp = lg.flatten()[Bg["logits_idx"].to(DEV)].cpu()
    d_ref = (samp - Bg["logits_vals"].double()).abs()
    d_or = (lg - lq).abs()
    m = {"loss": loss, "loss_reference_fp32": ref_loss, "loss_reference_bf16_autocast": 1.366608,
         "loss_oracle_bf16": float(Lq["total"]), "logits_std_reference": Bg["logits_std"],
         "logit_samples_vs_reference": {"mean": float(d_ref.mean()), "max": float(d_ref.max()), "n": int(d_ref.numel())},
         "logits_vs_oracle_bf16": {"mean": float(d_or.mean()), "max": float(d_or.max())}, "grads": {}}
    for k, p in net.named_parameters():
        gq = Gq[k].to(p.grad.device)
        m["grads"][k] = {"norm": float(p.grad.double().norm()), "ref_norm": Bg["grad_norms"][k],
                         "rel_l2_vs_oracle": rel_l2(p.grad, gq), "cos_vs_oracle": cosine(p.grad, gq)}
    _dump("golden_case_B_256.json", m)

    # known answers of the reference (BCE 0.746483 + Dice 0.620441 = 1.366924; bf16 autocast 1.366608)
    assert abs(ref_loss - 1.366924) < 2e-6
    assert abs(loss - ref_loss) < 5e-3, m
    assert abs(loss - 1.366608) < 2e-3, m
    assert abs(loss - float(Lq["total"])) < 1e-3, m
    # logits: same precision model -> tight; vs the fp32 reference within 2x the reference's own bf16 noise (App. C)
    assert m["logits_vs_oracle_bf16"]["mean"] < 6e-3, m["logits_vs_oracle_bf16"]
    assert m["logit_samples_vs_reference"]["mean"] < 2.5e-2, m["logit_samples_vs_reference"]
    # gradients on pure noise: the reference's OWN bf16 run is 22-60 % rel. L2 away from its fp32 run everywhere but
    # the head, so norms are held to a factor and the head tightly; direction against the bf16 oracle
    for k, v in m["grads"].items():
        assert 0.5 * v["ref_norm"] - 1e-9 <= v["norm"] <= 2.0 * v["ref_norm"] + 1e-9, (k, v)
        assert v["cos_vs_oracle"] > 0.70, (k, v)
    for k in ("final.1.weight", "final.1.bias", "final.0.5.weight", "final.0.5.bias"):
        v = m["grads"][k]
        assert abs(v["norm"] - v["ref_norm"]) <= 0.05 * v["ref_norm"] + 1e-9, (k, v)
        assert v["rel_l2_vs_oracle"] < 0.10 and v["cos_vs_oracle"] > 0.995, (k, v)


@pytest.mark.parametrize("B,H,W", [(2, 256, 256), (1, 256, 512), (2, 128, 256)])
def test_backward_given_forward_state_at_benchmark_width(net, ref_params, B, H, W):
    """Same statement as test_unet_gpu.py::test_backward_given_forward_state (the backward is linear in dlogits once
    the forward state is fixed) on shapes whose 256^2 / 128^2 levels run conv2_tc_kernel<..,HALO=1> and
    wgrad_halo_kernel; the oracle's explicit backward runs in fp64 on the device."""
    net.load_state_dict(ref_params, strict=True)
    net.train()
    net.zero_grad(set_to_none=True)
    x, t = O.synth_batch(B, H, W, seed=78)
    xg = x.to(DEV)
    logits = net(xg)
    P = _p64(ref_params, DEV)
    lq = O.unet_forward(P, x.double().to(DEV), train=True, q=O.bf16_round)
    assert float((logits.detach().double() - lq).abs().mean()) < 8e-3
    dl = O.seg_loss(logits.detach().double(), t.double().to(DEV))["dlogits"]
    logits.backward(dl.float())
    torch.cuda.synchronize()
    plan = net._engine.plans[(B, H, W, str(xg.device))]
    cache = _cache_from_plan(plan, x, dev=DEV)
    Gq = O.unet_backward(P, cache, dl, q=O.bf16_round)
    res = {k: (rel_l2(p.grad, Gq[k]), cosine(p.grad, Gq[k])) for k, p in net.named_parameters()}
    _dump(f"backward_given_state_{B}x{H}x{W}.json", res)
    bad = {k: v for k, v in res.items() if v[0] > 0.05 or v[1] < 0.998}
    assert not bad, bad


def test_cuda_graph_step_equals_eager_step_256(ref_params):
    """TrainStep.capture / step_graphed against the host-launched step at 4 x 256^2 (the benchmarked kernel variants)."""
    from b200seg.train import TrainStep
    x, t = O.synth_batch(4, 256, 256, seed=7)
    x2, t2 = O.synth_batch(4, 256, 256, seed=8)
    x, t, x2, t2 = (v.to(DEV) for v in (x, t, x2, t2))
    eager = TrainStep({k: v.clone() for k, v in ref_params.items()}, DEV, lr=1e-3)
    graph = TrainStep({k: v.clone() for k, v in ref_params.items()}, DEV, lr=1e-3)
    assert graph.capture(x, t), getattr(graph, "capture_error", "")
    for _ in range(2):
        eager.step(x, t)
    losses = []
    for i, lr in enumerate([1e-3, 5e-4, 2e-3]):
        xb, tb = (x, t) if i % 2 == 0 else (x2, t2)
        le = eager.step(xb, tb, lr=lr).clone()
        lg = graph.step_graphed(xb, tb, lr=lr).clone()
        losses.append((float(le[0]), float(lg[0])))
    torch.cuda.synchronize()
    for a, b in losses:
        assert abs(a - b) <= 1e-6 * max(1.0, abs(a)), losses
    se, sg = eager.state_dict(), graph.state_dict()
    for k in se:
        assert torch.allclose(se[k].float(), sg[k].float(), rtol=1e-6, atol=1e-7), k
    assert graph.graph_launches > 150
