"""models/mod.py variants (UNet with Conv->BN->ReLU blocks, ResUNet): the CPU oracle (oracle/mod_oracle.py) against
goldens generated from the unmodified reference (oracle/make_golden_mod.py), and the drop-in modules' parameter layout /
seeded initialisation. CPU only."""
import os

import pytest
import torch

from oracle import unet_oracle as O
from oracle import mod_oracle as M

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mod_golden.pt")


@pytest.fixture(scope="module")
def mg():
    return torch.load(GOLDEN, weights_only=False)


# golden case -> (drop-in class, constructor kwargs, oracle forward)
CASES = {"ResUNet": ("ResUNet", {}, M.resunet_forward), "UNet": ("UNet", {}, M.unet_forward),
         "AttentionUNet": ("AttentionUNet", {}, M.attention_unet_forward),
         "UNet_odd": ("UNet", {}, M.unet_forward), "AttentionUNet_odd": ("AttentionUNet", {}, M.attention_unet_forward),
         "ResUNet_rgb": ("ResUNet", {"in_channels": 3}, M.resunet_forward),
         "UNet_rgb": ("UNet", {"in_channels": 3}, M.unet_forward)}


def build(name):
    import b200seg  # noqa: F401
    from b200seg.models import mod
    cls, kw, _ = CASES[name]
    torch.manual_seed(42)
    return getattr(mod, cls)(depth=3, **kw)


@pytest.mark.parametrize("name", list(CASES))
def test_layout_init_and_forward_match_reference(mg, name):
    g = mg[name]
    net = build(name)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    assert list(sd.keys()) == g["state_dict_keys"]
    for k, d in g["init_digest"].items():
        assert abs(float(sd[k].double().sum()) - d["sum"]) <= 1e-9 * max(1.0, d["abs_sum"]), k
    P = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    fwd = CASES[name][2]
    stats = {}
    logits = fwd(P, g["x"].double(), 3, train=True, stats_out=stats)
    assert float((logits - g["logits"].double()).abs().max()) < 2e-4
    L = O.seg_loss(logits, g["t"].double())
    assert abs(float(L["bce"]) - g["bce"]) < 2e-5 and abs(float(L["dice"]) - g["dice"]) < 2e-5
    from oracle.vnet_oracle import running_stats_update
    P2 = dict(P); P2.update(running_stats_update(P, stats))
    le = fwd(P2, g["x"].double(), 3, train=False)
    assert float((le - g["eval_logits"].double()).abs().max()) < 2e-4


def test_default_parameter_counts(mg):
    import b200seg  # noqa: F401
    from b200seg.models import mod
    assert sum(p.numel() for p in mod.ResUNet().parameters()) == mg["param_count_default"]["ResUNet"]
    assert sum(p.numel() for p in mod.UNet().parameters()) == mg["param_count_default"]["UNet"]
    assert sum(p.numel() for p in mod.AttentionUNet().parameters()) == mg["param_count_default"]["AttentionUNet"]


def test_bilinear_resize_matches_torch():
    """the oracle's restatement of F.interpolate(mode='bilinear', align_corners=False), values and gradient"""
    g = torch.Generator().manual_seed(2)
    for (hi, wi, ho, wo) in ((8, 10, 9, 11), (4, 5, 9, 11), (7, 7, 5, 3), (6, 6, 6, 6)):
        x = torch.randn((2, 3, hi, wi), generator=g, dtype=torch.float64, requires_grad=True)
        dy = torch.randn((2, 3, ho, wo), generator=g, dtype=torch.float64)
        y = M.bilinear_resize(x, ho, wo)
        y.backward(dy)
        xr = x.detach().clone().requires_grad_(True)
        yr = torch.nn.functional.interpolate(xr, size=(ho, wo), mode="bilinear", align_corners=False)
        yr.backward(dy)
        assert float((y - yr).abs().max()) < 1e-12 and float((x.grad - xr.grad).abs().max()) < 1e-12


def test_first_max_pool_gradient():
    x = torch.tensor([[[[1., 1.], [1., 0.]]]], dtype=torch.float64, requires_grad=True)
    M.maxpool_first(x).sum().backward()
    assert x.grad.flatten().tolist() == [1.0, 0.0, 0.0, 0.0]
