"""V-Net variant: the CPU oracle (oracle/vnet_oracle.py) against golden vectors generated from the unmodified reference
(oracle/make_golden_vnet.py -> tests/golden/vnet_golden.pt), and the drop-in module's parameter layout / seeded
initialisation against the reference's. CPU only."""
import os

import pytest
import torch

from oracle import unet_oracle as O
from oracle import vnet_oracle as V

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vnet_golden.pt")


@pytest.fixture(scope="module")
def vg():
    return torch.load(GOLDEN, weights_only=False)


def close(a, b, tol=1e-9):
    a, b = a.double(), b.double()
    assert a.shape == b.shape
    assert float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max())), float((a - b).abs().max())


def test_strided_conv_matches_torch(vg):
    c = vg["conv_s2"]
    x, w, b = (c[k].clone().requires_grad_(True) for k in ("x", "w", "b"))
    z = V.conv3x3_s2(x, w, b)
    close(z, c["z"])
    z.backward(c["dz"])
    close(x.grad, c["dx"]); close(w.grad, c["dw"]); close(b.grad, c["db"])


def test_se_block_matches_reference(vg):
    c = vg["se"]
    P = {k: v.clone().requires_grad_(True) for k, v in c["params"].items()}
    x = c["x"].clone().requires_grad_(True)
    y = V.se_block(x, P["fc1.weight"], P["fc1.bias"], P["fc2.weight"], P["fc2.bias"])
    close(y, c["y"])
    y.backward(c["dy"])
    close(x.grad, c["dx"])
    for k, g in c["grads"].items():
        close(P[k].grad, g)


@pytest.mark.parametrize("name", ["block_proj", "block_id"])
def test_conv_block_matches_reference(vg, name):
    c = vg[name]
    P = {f"blk.{k}": (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
         for k, v in c["params"].items()}
    x = c["x"].clone().requires_grad_(True)
    y = V.conv_block(P, "blk", x, c["num_convs"], True, O.identity, O.identity, first=False)
    close(y, c["y"], 1e-8)
    y.backward(c["dy"])
    close(x.grad, c["dx"], 1e-7)
    for k, g in c["grads"].items():
        close(P[f"blk.{k}"].grad, g, 1e-7)


@pytest.fixture(scope="module")
def vnet_params():
    import b200seg  # noqa: F401
    from b200seg.models.vnet import ImprovedVNet
    torch.manual_seed(42)
    net = ImprovedVNet(dropout_rate=0.0)
    return net, {k: v.detach().clone() for k, v in net.state_dict().items()}


def test_module_layout_and_seeded_init_match_reference(vg, vnet_params):
    net, sd = vnet_params
    assert sum(p.numel() for p in net.parameters() if p.requires_grad) == vg["param_count"] == 160435681
    assert list(sd.keys()) == vg["state_dict_keys"]
    assert {k: tuple(v.shape) for k, v in sd.items()} == vg["state_dict_shapes"]
    for k, d in vg["init_digest"].items():
        v = sd[k].double()
        assert abs(float(v.sum()) - d["sum"]) <= 1e-9 * max(1.0, d["abs_sum"]), k
        assert abs(float(v.abs().sum()) - d["abs_sum"]) <= 1e-9 * max(1.0, d["abs_sum"]), k


def test_vnet_forward_matches_reference(vg, vnet_params):
    _, sd = vnet_params
    A = vg["A"]
    P = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    stats = {}
    logits = V.vnet_forward(P, A["x"].double(), train=True, stats_out=stats)
    d = (logits - A["logits"].double()).abs()
    assert float(d.max()) < 2e-4, float(d.max())          # the reference ran in fp32
    L = O.seg_loss(logits, A["t"].double())
    assert abs(float(L["bce"]) - A["bce"]) < 2e-5 and abs(float(L["dice"]) - A["dice"]) < 2e-5
    upd = V.running_stats_update(P, stats)
    for k, dg in A["running_digest"].items():
        assert abs(float(upd[k].sum()) - dg["sum"]) <= 1e-4 * max(1.0, dg["abs_sum"]), k
    # eval mode with the updated running statistics
    P2 = dict(P); P2.update(upd)
    le = V.vnet_forward(P2, A["x"].double(), train=False)
    assert float((le - A["eval_logits"].double()).abs().max()) < 2e-4
    band = A["eval_logits"].abs() > 1e-4
    assert bool((O.threshold_mask(le.float())[band] == A["eval_mask"][band]).all())


def test_vnet_rejects_cpu_and_bad_sizes(vnet_params):
    net, _ = vnet_params
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 1, 32, 32))            # CPU tensor: no fallback
    with pytest.raises(AssertionError):
        net(torch.zeros(1, 2, 32, 32))            # wrong channel count (models/vnet.py:118)
