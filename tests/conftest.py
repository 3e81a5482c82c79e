import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ops_golden():
    import torch
    return torch.load(os.path.join(GOLDEN, "ops_golden.pt"), weights_only=False)


@pytest.fixture(scope="session")
def unet_golden():
    import torch
    return torch.load(os.path.join(GOLDEN, "unet_golden.pt"), weights_only=False)


@pytest.fixture(scope="session")
def ref_params():
    """Reference-initialised UNet parameters: torch.manual_seed(42) through the drop-in module's containers
    (bit-identical to the reference's init; test_oracle_cpu checks the digests against the golden file)."""
    import torch
    import b200seg  # noqa: F401
    from b200seg.models.model import UNet
    torch.manual_seed(42)
    m = UNet()
    return {k: v.detach().clone() for k, v in m.state_dict().items()}
