"""Data-parallel step on real NCCL ranks (needs >= 2 GPUs; the 1-GPU round-end box skips it — run with
`gpurun --gpus 2 -- python -m pytest tests/test_ddp_gpu.py -m gpu`; its report is committed under profiles/)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("size", [64, 256])
def test_nccl_ranks_match_chunked_single_gpu_replicas(size):
    n = min(_ngpu(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "run_ddp_check.py"), "--size", str(size)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith("DDP_JSON ")]
    assert out.returncode == 0 and lines, out.stdout[-3000:] + out.stderr[-6000:]
    rep = json.loads(lines[-1][len("DDP_JSON "):])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"ddp_check_n{n}_s{size}.json"), "w") as f:
        json.dump(rep, f, indent=1)
    c = rep["dp_check"]
    assert c["ranks_hold_bit_identical_parameters"], c
    assert c["ok"], c
    if n == 2:      # a + b in either order is the same fp32 sum
        assert c["gradient_vs_chunked_replicas"]["bit_identical"], c
    g = rep["graph_vs_eager"]
    assert g["captured"] and g["graph_ranks_agree"] and g["eager_ranks_agree"], g
    assert g["max_rel_param_diff_over_ranks"] < 1e-5, g
    for a, b in g["losses"]:
        assert abs(a - b) <= 1e-5 * max(1.0, abs(a)), g


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_data_parallel_equals_chunked_single_gpu():
    """nn.DataParallel over two replicas of the drop-in UNet (one host thread per GPU, reference utils/trainer.py:28-30)
    == the same two chunks run one after the other on one GPU: logits, every parameter gradient (torch sums the replica
    gradients in Broadcast.backward) and the master's BatchNorm buffers (replica 0's update)."""
    import torch
    import b200seg  # noqa: F401
    from b200seg.models.model import UNet
    from b200seg.synth import synth_batch
    torch.manual_seed(42)
    net = UNet().to("cuda:0").train()
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x, t = synth_batch(8, 64, 64, seed=31)
    x, t = x.to("cuda:0"), t.to("cuda:0")
    crit = torch.nn.BCEWithLogitsLoss()
    dp = torch.nn.DataParallel(net, device_ids=[0, 1])
    net.zero_grad(set_to_none=True)
    logits_dp = dp(x)
    crit(logits_dp, t).backward()
    g_dp = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    sd_dp = {k: v.detach().clone() for k, v in net.state_dict().items()}
    # chunked on one GPU: dlogits of the full-batch loss, then forward + backward per chunk (gradients accumulate)
    net.load_state_dict(sd0)
    net.zero_grad(set_to_none=True)
    lg = logits_dp.detach().clone().requires_grad_(True)
    crit(lg, t).backward()
    outs = []
    for c in range(2):
        sl = slice(4 * c, 4 * c + 4)
        if c == 1:      # replica 1's BatchNorm update is discarded by DataParallel: keep replica 0's buffers
            keep = {k: v.detach().clone() for k, v in net.state_dict().items() if "running" in k or "num_batches" in k}
        o = net(x[sl].contiguous())
        o.backward(lg.grad[sl].contiguous())
        outs.append(o.detach())
    assert torch.equal(torch.cat(outs), logits_dp.detach()), "replica logits differ from the chunked run"
    for k, p in net.named_parameters():
        assert p.grad is not None
        d = float((p.grad - g_dp[k]).abs().max())
        assert d <= 1e-6 * float(g_dp[k].abs().max()) + 1e-12, (k, d)
    for k, v in keep.items():
        assert torch.equal(v, sd_dp[k]), k
    # eval mode (Trainer.validate / test under DataParallel): samples are independent, so the replicas' logits are the
    # single-GPU logits of the same frames
    net.load_state_dict(sd_dp)
    net.eval()
    with torch.no_grad():
        e_dp = dp(x)
        e_one = torch.cat([net(x[:4].contiguous()), net(x[4:].contiguous())])
        e_full = net(x)
    assert torch.equal(e_dp, e_one), float((e_dp - e_one).abs().max())
    assert float((e_dp - e_full).abs().max()) < 1e-5


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_modules_on_a_second_gpu_without_set_device():
    """libb2s launches on the current device; the entry points make the input's device current themselves
    (a model on cuda:1 called while cuda:0 is current must not launch on cuda:0)."""
    import torch
    import b200seg  # noqa: F401
    from b200seg.models.model import UNet
    from b200seg.models.mod import ResUNet
    from b200seg.models.loss import BCEDiceLoss
    from b200seg.models.metrics import SegMetrics
    from b200seg.synth import synth_batch
    assert torch.cuda.current_device() == 0
    x, t = synth_batch(2, 32, 32, seed=3)
    for make in (UNet, lambda: ResUNet(depth=3)):
        outs = []
        for dev in ("cuda:0", "cuda:1"):
            torch.manual_seed(42)
            net = make().to(dev).train()
            lg = net(x.to(dev))
            loss = BCEDiceLoss()(lg, t.to(dev))
            loss.backward()
            m = SegMetrics(dev)
            m.update(lg.detach(), t.to(dev))
            torch.cuda.synchronize(dev)
            g = next(p.grad for p in net.parameters() if p.grad is not None)
            outs.append((lg.detach().cpu(), float(loss), g.cpu(), m.compute()["iou"]))
        assert torch.equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]
        assert torch.equal(outs[0][2], outs[1][2]) and outs[0][3] == outs[1][3]
    assert torch.cuda.current_device() == 0
