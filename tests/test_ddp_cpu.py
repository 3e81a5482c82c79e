"""Host-side logic of the data-parallel step on CPU: flat bucket layout, gradient-ready order, bucketed all-reduce
hook (gloo, world_size 2), LR schedule, C-ABI symbol export. No kernels are launched (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import b200seg  # noqa: F401
from b200seg import _lib
from b200seg.models.model import UNet
from b200seg.train import TrainStep, cosine_warm_restarts_lr, grad_ready_order

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b2s.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    decls = re.findall(r"\b(b2s_\w+)\s*\(([^)]*)\)\s*;", hdr)
    assert len(decls) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name, args in decls:
        assert hasattr(lib, name), f"libb2s.so does not export {name}"
        nargs = 0 if args.strip() in ("", "void") else len(args.split(","))
        assert name in _lib.SIGNATURES, f"no ctypes binding for {name}"
        assert len(_lib.SIGNATURES[name][1]) == nargs, f"ctypes arity mismatch for {name}"
    assert set(_lib.SIGNATURES) == {n for n, _ in decls}
    # host-only entry points are callable without a GPU
    L = _lib.lib()
    assert L.b2s_version() >= 100 and L.b2s_ew_rows() % 148 == 0
    assert 2 <= L.b2s_conv_stats_rows(64, 256, 256, 64, 0) <= 2 * 160
    s = ctypes.c_int(0)
    assert L.b2s_conv_wgrad_workspace(64, 256, 256, 64, 64, 3, 0, 0, ctypes.byref(s)) > 0 and s.value >= 1
    assert L.b2s_conv_wgrad_workspace(1, 16, 16, 48, 64, 3, 0, 0, ctypes.byref(s)) < 0     # Cin % 64 != 0
    assert b"unsupported" in L.b2s_last_error()


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libb2s.so")
    with pytest.raises(_lib.B2SError):
        _lib.lib()


def test_cpu_tensors_are_rejected():
    m = UNet()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 32, 32))
    from b200seg.models.loss import DiceLoss
    with pytest.raises(RuntimeError):
        DiceLoss()(torch.zeros(1, 1, 8, 8), torch.zeros(1, 1, 8, 8))


def test_grad_ready_order_covers_all_parameters():
    m = UNet()
    names = [n for n, _ in m.named_parameters()]
    order = grad_ready_order()
    assert sorted(order) == sorted(names) and len(order) == 82
    assert order[0] == "final.1.weight" and order[-1] == "encoder1.0.weight"


def test_bucket_layout():
    m = UNet()
    ts = TrainStep(m.state_dict(), "cpu", use_dist=False, bucket_mb=16.0)
    # buckets tile the flat buffer exactly once, in order
    assert ts.buckets[0][0] == 0 and ts.buckets[-1][1] == ts.flat_g.numel()
    for (s0, e0), (s1, e1) in zip(ts.buckets, ts.buckets[1:]):
        assert e0 == s1 and e0 > s0
    # parameter views alias the flat buffers and carry the initial values
    sd = m.state_dict()
    for k in ts.order:
        assert torch.equal(ts.P[k], sd[k])
        assert ts.P[k].data_ptr() >= ts.flat_p.data_ptr() and ts.G[k].data_ptr() % 16 == 0
    assert sum(ts.P[k].numel() for k in ts.order) == 31042369
    # each bucket's trigger is the last-ready parameter inside it
    for b, (s0, e0) in enumerate(ts.buckets):
        last = ts.bucket_last[b]
        off = (ts.G[last].data_ptr() - ts.flat_g.data_ptr()) // 4
        assert s0 <= off < e0


def test_cosine_warm_restarts_matches_torch():
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1e-3)
    sch = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=20, T_mult=2, eta_min=0)
    for epoch in range(70):
        assert abs(opt.param_groups[0]["lr"] - cosine_warm_restarts_lr(1e-3, epoch)) < 1e-12
        opt.step()
        sch.step()


class _FakeEngine:
    """Stands in for UNetEngine on CPU: 'computes' rank-dependent gradients in gradient-ready order."""

    def __init__(self, rank):
        self.rank = rank
        self.calls = []

    def invalidate_packed(self):
        pass

    def forward(self, P, x, train):
        return None, None

    def loss(self, pl, t, **kw):
        return torch.zeros(8)

    def loss_backward(self, pl, t, **kw):
        return None

    def backward(self, P, pl, dl, G, on_grad_ready=None):
        for i, k in enumerate(grad_ready_order()):
            G[k].fill_(float(self.rank + 1) * (1 + i % 7))
            self.calls.append(k)
            on_grad_ready(k)


def _ddp_worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        sd = UNet().state_dict()
        ts = TrainStep(sd, "cpu", bucket_mb=8.0, engine=_FakeEngine(rank))
        assert ts.world == world and len(ts.buckets) > 3
        ts.forward_backward(torch.zeros(1), torch.zeros(1))
        # every gradient element now holds the SUM over ranks: (1 + 2) * (1 + i % 7)
        for i, k in enumerate(ts.order):
            expect = 3.0 * (1 + i % 7)
            assert float(ts.G[k].min()) == expect == float(ts.G[k].max()), k
        assert ts.engine.calls == ts.order
        open(os.path.join(outdir, f"ok{rank}"), "w").write("1")
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo(tmp_path):
    world = 2
    port = 29000 + os.getpid() % 2000
    mp.spawn(_ddp_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_host_side_planners_over_many_shapes():
    """the launch planners behind the C ABI are pure host code: exercise them without a GPU over the UNet / V-Net layer
    shapes and odd sizes (grid <= SM count per wave, partial-statistics rows, split-K workspace consistency)"""
    L = _lib.lib()
    s = ctypes.c_int(0)
    PAIR, HALO, LEGACY = 1 << 10, 1 << 11, 1 << 12
    for N in (1, 2, 3, 16, 64):
        for (H, W) in ((16, 16), (32, 32), (48, 80), (64, 64), (128, 128), (256, 256), (32, 128), (512, 512)):
            for C in (64, 128, 256, 512, 1024):
                if N * H * W * C > 64 * 512 * 512 * 128:
                    continue
                rows = L.b2s_conv_stats_rows(N, H, W, C, 0)
                assert 2 <= rows <= 2 * 148 and rows % 2 == 0, (N, H, W, C, rows)
                for forced in (PAIR, LEGACY) + ((HALO,) if W % 128 == 0 and H % 2 == 0 else ()):
                    r = L.b2s_conv_stats_rows(N, H, W, C, forced)
                    assert 2 <= r <= 2 * 148, (N, H, W, C, forced, r)
                if not (W % 128 == 0 and H % 2 == 0):
                    assert L.b2s_conv_stats_rows(N, H, W, C, HALO) < 0          # halo kernel refuses narrow images
                for cin in (64, 128, 256):
                    nbytes = L.b2s_conv_wgrad_workspace(N, H, W, cin, C, 3, 0, 0, ctypes.byref(s))
                    assert nbytes == s.value * 9 * cin * C * 4 and 1 <= s.value <= 148, (N, H, W, cin, C, s.value)
                    nb1 = L.b2s_conv_wgrad_workspace(N, H, W, cin, C, 1, 0, 0, ctypes.byref(s))
                    assert nb1 == s.value * cin * C * 4 and s.value >= 1
                    if cin % 128 == 0:
                        nb4 = L.b2s_conv_wgrad_workspace(N, H, W, cin, C, 4, 0, 0, ctypes.byref(s))
                        assert nb4 == s.value * 4 * cin * C * 4 and s.value >= 1
                    forced_splits = L.b2s_conv_wgrad_workspace(N, H, W, cin, C, 3, LEGACY, 3, ctypes.byref(s))
                    assert forced_splits > 0 and 1 <= s.value <= 3
    assert L.b2s_metrics_blocks(1) == 1 and L.b2s_loss_chunks(65536) >= 1
    # SE pooling chunks: every pixel covered, between 4 and 64 blocks per sample at the V-Net levels (enough blocks at
    # 32^2, no more than 64 partial rows for the FC kernels to sum at 512^2), never decreasing with the image size
    prev = 0
    for side in (1, 2, 7, 8, 16, 32, 64, 100, 128, 256, 512, 1024):
        hw = side * side
        c = L.b2s_se_chunks(hw)
        assert c >= 1 and c * 4096 >= hw and c >= prev, (side, c)
        if 32 <= side <= 512 and side & (side - 1) == 0:
            assert 4 <= c <= 64, (side, c)
        prev = c
    assert L.b2s_c1_rows(64, 256, 256) % 148 == 0


def test_library_is_sm100a_only_and_uses_the_blackwell_units():
    """the shipped libb2s.so holds sm_100a cubins only, and its SASS carries the instructions the design rests on:
    tcgen05 MMA / commit / TMEM loads (UTCHMMA, UTCBAR, LDTM), TMA loads and stores (UTMALDG, UTMASTG), the async
    global->shared copies of the prefetch ring (LDGSTS) and packed fp32 FMAs (FFMA2). Needs cuobjdump (CUDA toolkit)."""
    import re
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elfs = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"\.(sm_\w+)\.cubin", elfs))
    assert archs == {"sm_100a"}, archs
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    for mnemonic in ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "LDGSTS", "FFMA2"):
        assert re.search(r"\b" + mnemonic + r"\b", sass), f"{mnemonic} missing from the SASS of libb2s.so"
    assert "HMMA." not in sass.replace("UTCHMMA", ""), "legacy mma.sync tensor-core instructions in libb2s.so"


def test_parameter_collection_on_data_parallel_replicas():
    """nn.DataParallel (reference utils/trainer.py:28-30) hands forward() a REPLICA whose `_parameters` are empty: torch's
    replicate() sets the broadcast copies as plain attributes and lists them in `_former_parameters`. The drop-in UNet
    must find all 82 tensors there (named_parameters() of a replica yields nothing). The replica tree is
    built here the way torch.nn.parallel.replicate builds it, on CPU (the real thing needs two GPUs:
    tests/test_ddp_gpu.py::test_data_parallel_equals_chunked_single_gpu)."""
    from collections import OrderedDict
    import b200seg  # noqa: F401
    from b200seg.models.model import UNet, _named_params
    torch.manual_seed(0)
    net = UNet()
    modules = list(net.modules())
    index = {m: i for i, m in enumerate(modules)}
    copies = [m._replicate_for_data_parallel() for m in modules]
    for r in copies:
        r._former_parameters = OrderedDict()
    for m, r in zip(modules, copies):
        for key, child in m._modules.items():
            r._modules[key] = None if child is None else copies[index[child]]
        for key, p in m._parameters.items():
            c = p * 1.0                       # a non-leaf tensor with a grad_fn, like Broadcast's outputs
            setattr(r, key, c)
            r._former_parameters[key] = c
        for key, b in m._buffers.items():
            r._buffers[key] = b
    replica = copies[0]
    assert len(list(replica.named_parameters())) == 0           # what broke round 1's module under DataParallel
    master = dict(net.named_parameters())
    got = _named_params(replica)
    assert list(got.keys()) == list(master.keys()) and len(got) == 82
    for k, v in got.items():
        assert v.shape == master[k].shape and not v.is_leaf and v.requires_grad, k
    assert list(_named_params(net).keys()) == list(master.keys())
    assert all(_named_params(net)[k] is master[k] for k in master)
    assert list(replica._tensor_dict().keys())[:82] == list(master.keys())
    assert len(replica._tensor_dict()) == 136
