"""Per-kernel parity of libb2s (through the C ABI via ctypes) against the oracle (CPU, fp64, on the same
bf16-rounded operands) and, at larger sizes, against plain torch fp32 ops on the GPU. Marked gpu.

Tolerances: tensor-core kernels accumulate in fp32 and store bf16 -> |err| <= 2^-8 * |ref| + 1e-3 * max|ref|
(one bf16 rounding of the output + accumulation-order noise); fp32-output kernels (wgrad, reductions, loss)
<= 1e-4 relative to max|ref| unless stated.
"""
import pytest
import torch

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    import b200seg  # noqa: F401
    from b200seg import ops as _ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _ops


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale)


def bf(t):
    return t.to(torch.bfloat16).float()


def report(name, got, ref, rel=2 ** -8, abs_frac=1e-3):
    got, ref = got.double().cpu(), ref.double().cpu()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    err = (got - ref).abs()
    tol = rel * ref.abs() + abs_frac * float(ref.abs().max()) + 1e-30
    bad = err > tol
    nbad = int(bad.sum())
    if nbad:
        idx = bad.nonzero()[:8]
        lines = [f"{name}: {nbad}/{err.numel()} elements out of tolerance; max err {float(err.max()):.4g}, "
                 f"max|ref| {float(ref.abs().max()):.4g}"]
        for i in idx:
            i = tuple(int(v) for v in i)
            lines.append(f"   at {i}: got {float(got[i]):.6g} ref {float(ref[i]):.6g}")
        pytest.fail("\n".join(lines))


def act_from_nchw(ops, x, ctot=None, c0=0):
    """NCHW float -> Act (optionally as a channel slice of a wider zero-filled buffer)."""
    N, C, H, W = x.shape
    ctot = ctot or C
    buf = torch.zeros((N, H, W, ctot), dtype=torch.bfloat16, device=DEV)
    buf[..., c0:c0 + C] = x.permute(0, 2, 3, 1).to(torch.bfloat16).to(DEV)
    return ops.Act(buf, c0, C)


# ------------------------------------------------------------------------------------------------------------
# tensor-core kernels
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,H,W,Cin,Cout,tile_n", [
    (1, 1, 128, 64, 64, 0),       # one tile, one K chunk: the minimal tcgen05 GEMM
    (1, 1, 128, 128, 64, 0),      # two K chunks
    (2, 16, 16, 128, 256, 64),
    (2, 16, 16, 128, 256, 128),
    (2, 16, 16, 128, 256, 256),
    (3, 8, 8, 64, 128, 0),        # tile spans two images, last tile partially out of range
])
def test_conv1x1_gemm(ops, N, H, W, Cin, Cout, tile_n):
    x = bf(rnd((N, Cin, H, W), 1))
    w = bf(rnd((Cout, Cin, 1, 1), 2, 0.1))
    b = rnd((Cout,), 3)
    xa = act_from_nchw(ops, x)
    wf, _ = ops.pack_conv_weight(w.to(DEV), want_dgrad=False)
    y = ops.Act.empty(N, H, W, Cout, DEV)
    ops.conv_fwd(xa, wf, b.to(DEV), y, ksize=1, tile_n=tile_n)
    torch.cuda.synchronize()
    ref = O.conv1x1(x.double(), w.double(), b.double())
    report("conv1x1", y.to_nchw_float(), ref)


@pytest.mark.parametrize("N,H,W,Cin,Cout,tile_n,relu", [
    (1, 16, 16, 64, 64, 0, False),
    (2, 16, 16, 64, 64, 0, True),
    (1, 32, 32, 128, 128, 0, True),
    (2, 8, 8, 64, 128, 64, True),      # 64-pixel images: two images per tile
    (1, 24, 40, 64, 64, 0, True),      # non power-of-two extent: partial tiles, masked stats
    (1, 16, 256, 64, 64, 0, True),     # W > 128: two tiles per row
    (2, 16, 16, 256, 512, 256, True),
    (2, 16, 16, 256, 512, 128, True),
])
def test_conv3x3_fwd_bias_relu_stats(ops, N, H, W, Cin, Cout, tile_n, relu):
    x = bf(rnd((N, Cin, H, W), 11))
    w = bf(rnd((Cout, Cin, 3, 3), 12, 0.05))
    b = rnd((Cout,), 13, 0.5)
    xa = act_from_nchw(ops, x)
    wf, _ = ops.pack_conv_weight(w.to(DEV), want_dgrad=False)
    y = ops.Act.empty(N, H, W, Cout, DEV)
    rows = ops.conv_stats_rows(N, H, W, Cout, tile_n)
    stats = torch.full((rows, 2, Cout), float('nan'), dtype=torch.float32, device=DEV)
    ops.conv_fwd(xa, wf, b.to(DEV), y, ksize=3, relu=relu, stats=stats, tile_n=tile_n)
    torch.cuda.synchronize()
    ref = O.conv3x3(x.double(), w.double(), b.double())
    if relu:
        ref = ref.clamp_min(0)
    got = y.to_nchw_float()
    report("conv3x3", got, ref)
    # statistics are those of the bf16 tensor actually stored
    s = stats.double().sum(dim=0).cpu()
    g64 = got.double().cpu()
    report("stats.sum", s[0], g64.sum(dim=(0, 2, 3)), rel=1e-5, abs_frac=1e-5)
    report("stats.sumsq", s[1], (g64 * g64).sum(dim=(0, 2, 3)), rel=1e-5, abs_frac=1e-5)


PAIR, HALO, LEGACY = 1 << 10, 1 << 11, 1 << 12   # kernel-variant bits of tile_n (csrc/conv_common.cuh)


@pytest.mark.parametrize("N,H,W,Cin,Cout,tile_n,relu", [
    (2, 16, 16, 64, 64, PAIR, True),            # tile-pair kernel, two sets of two accumulators
    (3, 8, 8, 64, 128, PAIR + 64, True),        # odd tile count: the last pair has one out-of-range tile
    (2, 16, 16, 256, 512, PAIR + 256, True),    # 256-wide tiles: one accumulator set (no epilogue overlap)
    (2, 16, 16, 256, 512, PAIR + 128, True),
    (1, 24, 40, 64, 64, PAIR, True),            # non power-of-two extent
    (5, 32, 32, 128, 128, PAIR, False),         # several items per CTA row-group? no: 40 tiles; phases still wrap
    (1, 2, 128, 64, 64, HALO, True),            # halo kernel: one item (two rows, one strip)
    (2, 4, 256, 64, 64, HALO, True),            # two strips per row
    (1, 6, 128, 128, 128, HALO, True),          # two K chunks, 128-wide tile
    (1, 4, 256, 128, 64, HALO, False),
    (3, 8, 128, 64, 256, HALO, True),           # Cout 256 = two column tiles
    (2, 16, 16, 64, 64, LEGACY, True),
])
def test_conv3x3_variants(ops, N, H, W, Cin, Cout, tile_n, relu):
    """every kernel variant behind b2s_conv_fwd, forced through the variant bits of tile_n"""
    x = bf(rnd((N, Cin, H, W), 111))
    w = bf(rnd((Cout, Cin, 3, 3), 112, 0.05))
    b = rnd((Cout,), 113, 0.5)
    xa = act_from_nchw(ops, x, ctot=Cin + 64, c0=64)
    wf, _ = ops.pack_conv_weight(w.to(DEV), want_dgrad=False)
    ybuf = torch.full((N, H, W, Cout + 64), 5.0, dtype=torch.bfloat16, device=DEV)
    y = ops.Act(ybuf, 0, Cout)
    rows = ops.conv_stats_rows(N, H, W, Cout, tile_n)
    stats = torch.full((rows, 2, Cout), float('nan'), dtype=torch.float32, device=DEV)
    ops.conv_fwd(xa, wf, b.to(DEV), y, ksize=3, relu=relu, stats=stats, tile_n=tile_n)
    torch.cuda.synchronize()
    ref = O.conv3x3(x.double(), w.double(), b.double())
    if relu:
        ref = ref.clamp_min(0)
    got = y.to_nchw_float()
    report("conv3x3", got, ref)
    assert bool((ybuf[..., Cout:] == 5.0).all()), "kernel wrote outside its channel slice"
    s = stats.double().sum(dim=0).cpu()
    g64 = got.double().cpu()
    report("stats.sum", s[0], g64.sum(dim=(0, 2, 3)), rel=1e-5, abs_frac=1e-5)
    report("stats.sumsq", s[1], (g64 * g64).sum(dim=(0, 2, 3)), rel=1e-5, abs_frac=1e-5)


def test_conv3x3_halo_many_items(ops):
    """halo kernel with more work items than SMs (ring phases wrap, both accumulator sets reused) vs torch fp32"""
    N, H, W, Cin, Cout = 6, 64, 128, 64, 64
    x = bf(rnd((N, Cin, H, W), 121)).to(DEV)
    w = bf(rnd((Cout, Cin, 3, 3), 122, 0.05)).to(DEV)
    b = rnd((Cout,), 123).to(DEV)
    wf, wd = ops.pack_conv_weight(w)
    y = ops.Act.empty(N, H, W, Cout, DEV)
    ops.conv_fwd(ops.Act.from_nchw(x), wf, b, y, ksize=3, relu=True, tile_n=HALO)
    ref = torch.nn.functional.conv2d(x, w, b, padding=1).clamp_min(0)
    report("halo conv3x3 large", y.to_nchw_float(), ref)
    y2 = ops.Act.empty(N, H, W, Cout, DEV)
    ops.conv_fwd(ops.Act.from_nchw(x), wf, b, y2, ksize=3, relu=True, tile_n=PAIR)
    report("pair conv3x3 large", y2.to_nchw_float(), ref)


@pytest.mark.parametrize("tile_n", [PAIR, PAIR + 64, LEGACY])
def test_conv1x1_and_convt_variants(ops, tile_n):
    N, H, W, Cin, Cout = 2, 16, 16, 128, 128
    x = bf(rnd((N, Cin, H, W), 131))
    w = bf(rnd((Cout, Cin, 1, 1), 132, 0.1))
    wf, _ = ops.pack_conv_weight(w.to(DEV), want_dgrad=False)
    y = ops.Act.empty(N, H, W, Cout, DEV)
    ops.conv_fwd(act_from_nchw(ops, x), wf, None, y, ksize=1, tile_n=tile_n)
    torch.cuda.synchronize()
    report("conv1x1", y.to_nchw_float(), O.conv1x1(x.double(), w.double(), None))
    wt = bf(rnd((Cin, 64, 2, 2), 133, 0.05))
    bt = rnd((64,), 134, 0.5)
    dy = bf(rnd((N, 64, 2 * H, 2 * W), 135))
    tf, td = ops.pack_convt_weight(wt.to(DEV))
    yt = ops.Act.empty(N, 2 * H, 2 * W, 64, DEV)
    ops.convt_fwd(act_from_nchw(ops, x), tf, bt.to(DEV), yt, tile_n=(tile_n & ~1023) + 64)
    dx = ops.Act.empty(N, H, W, Cin, DEV)
    ops.convt_dgrad(act_from_nchw(ops, dy), td, dx, tile_n=tile_n)
    torch.cuda.synchronize()
    report("convT fwd", yt.to_nchw_float(), O.conv_transpose2x2(x.double(), wt.double(), bt.double()))
    rdx, _, _ = O.conv_transpose2x2_bwd(x.double(), wt.double(), dy.double())
    report("convT dgrad", dx.to_nchw_float(), rdx)


def test_conv3x3_channel_slices(ops):
    """input read from, and output written into, channel slices of wider (concat) buffers"""
    N, H, W, Cin, Cout = 1, 16, 16, 64, 64
    x = bf(rnd((N, Cin, H, W), 21))
    w = bf(rnd((Cout, Cin, 3, 3), 22, 0.05))
    xa = act_from_nchw(ops, x, ctot=192, c0=64)
    wf, _ = ops.pack_conv_weight(w.to(DEV), want_dgrad=False)
    ybuf = torch.full((N, H, W, 128), 7.0, dtype=torch.bfloat16, device=DEV)
    y = ops.Act(ybuf, 64, 64)
    ops.conv_fwd(xa, wf, None, y, ksize=3)
    torch.cuda.synchronize()
    report("conv3x3 slice", y.to_nchw_float(), O.conv3x3(x.double(), w.double()))
    assert bool((ybuf[..., :64] == 7.0).all()), "kernel wrote outside its channel slice"


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 16, 16, 64, 64), (1, 32, 32, 128, 64), (2, 8, 8, 128, 256)])
def test_conv3x3_dgrad(ops, N, H, W, Cin, Cout):
    x = bf(rnd((N, Cin, H, W), 31))
    w = bf(rnd((Cout, Cin, 3, 3), 32, 0.05))
    dz = bf(rnd((N, Cout, H, W), 33))
    _, wd = ops.pack_conv_weight(w.to(DEV))
    dza = act_from_nchw(ops, dz)
    dx = ops.Act.empty(N, H, W, Cin, DEV)
    ops.conv_fwd(dza, wd, None, dx, ksize=3)
    torch.cuda.synchronize()
    ref, _, _ = O.conv3x3_bwd(x.double(), w.double(), dz.double())
    report("dgrad", dx.to_nchw_float(), ref)


@pytest.mark.parametrize("N,H,W,Cin,Cout,tile_n,splits", [
    (1, 16, 16, 64, 64, 0, 1),      # Cin = 64: two taps per M tile, odd row-block count
    (2, 16, 16, 64, 64, 0, 0),
    (1, 32, 32, 128, 64, 0, 3),
    (2, 16, 16, 128, 128, 0, 0),
    (2, 8, 8, 256, 256, 256, 0),
    (1, 24, 40, 64, 128, 0, 2),     # partial pixel chunks
    (1, 4, 128, 64, 64, 0, 0),      # row-halo kernel (W % 128 == 0): tap pairs, one CTA group
    (2, 6, 256, 64, 64, 0, 5),      # two strips per row, uneven split of the row tiles
    (1, 4, 128, 128, 64, 0, 0),     # Cin 128: one tap per tile, two CTA groups (5 + 4 taps)
    (1, 4, 128, 64, 128, 0, 3),     # Cout 128: two CTA groups (3 + 2 tiles)
    (2, 6, 128, 128, 128, 0, 0),    # three CTA groups, one tap row each
    (1, 4, 128, 256, 64, 0, 0),     # Cin 256: 18 (tap, half) tiles in three CTA groups
    (1, 4, 128, 64, 64, 1 << 12, 0),  # same shape through the legacy kernel
])
def test_conv3x3_wgrad(ops, N, H, W, Cin, Cout, tile_n, splits):
    x = bf(rnd((N, Cin, H, W), 41))
    dz = bf(rnd((N, Cout, H, W), 42))
    xa, dza = act_from_nchw(ops, x), act_from_nchw(ops, dz)
    nbytes, s = ops.wgrad_workspace(N, H, W, Cin, Cout, 9, tile_n, splits)
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=DEV)
    dw = torch.empty((Cout, Cin, 3, 3), dtype=torch.float32, device=DEV)
    ops.conv3x3_wgrad(xa, dza, ws, dw, tile_n=tile_n, splits=splits)
    torch.cuda.synchronize()
    _, ref, _ = O.conv3x3_bwd(x.double(), torch.zeros((Cout, Cin, 3, 3), dtype=torch.float64), dz.double())
    report("wgrad", dw, ref, rel=1e-4, abs_frac=1e-4)


@pytest.mark.parametrize("N,Hi,Wi,Cin,Cout", [(2, 8, 8, 128, 64), (1, 16, 16, 256, 128), (2, 4, 4, 128, 64),
                                              (1, 8, 16, 128, 64), (3, 3, 5, 128, 64), (2, 6, 6, 256, 64)])
def test_conv_transpose_fwd_dgrad_wgrad(ops, N, Hi, Wi, Cin, Cout):
    x = bf(rnd((N, Cin, Hi, Wi), 51))
    w = bf(rnd((Cin, Cout, 2, 2), 52, 0.05))
    b = rnd((Cout,), 53, 0.5)
    dy = bf(rnd((N, Cout, 2 * Hi, 2 * Wi), 54))
    wf, wd = ops.pack_convt_weight(w.to(DEV))
    xa = act_from_nchw(ops, x)
    # forward into the first half of a concat buffer
    ybuf = torch.full((N, 2 * Hi, 2 * Wi, 2 * Cout), 3.0, dtype=torch.bfloat16, device=DEV)
    y = ops.Act(ybuf, 0, Cout)
    ops.convt_fwd(xa, wf, b.to(DEV), y)
    torch.cuda.synchronize()
    report("convT fwd", y.to_nchw_float(), O.conv_transpose2x2(x.double(), w.double(), b.double()))
    assert bool((ybuf[..., Cout:] == 3.0).all())
    # backward
    dya = act_from_nchw(ops, dy, ctot=2 * Cout, c0=0)
    dx = ops.Act.empty(N, Hi, Wi, Cin, DEV)
    ops.convt_dgrad(dya, wd, dx)
    nbytes, _ = ops.wgrad_workspace(N, Hi, Wi, Cin, Cout, 4)
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=DEV)
    dw = torch.empty((Cin, Cout, 2, 2), dtype=torch.float32, device=DEV)
    ops.convt_wgrad(xa, dya, ws, dw)
    torch.cuda.synchronize()
    rdx, rdw, _ = O.conv_transpose2x2_bwd(x.double(), w.double(), dy.double())
    report("convT dgrad", dx.to_nchw_float(), rdx)
    report("convT wgrad", dw, rdw, rel=1e-4, abs_frac=1e-4)


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 16, 16, 64, 64), (1, 32, 32, 128, 64), (2, 8, 8, 128, 256),
                                            (2, 4, 128, 64, 64), (1, 6, 256, 64, 128), (3, 24, 40, 64, 128)])
def test_conv3x3_dgrad_with_fused_bn_reduction(ops, N, H, W, Cin, Cout):
    """b2s_conv_dgrad_bnred: same dx as the plain input-gradient launch (bit for bit) plus the BatchNorm-backward
    partials {sum dx, sum dx * r} over the stored bf16 values; tile-pair and row-halo (W % 128 == 0) kernels, a ragged
    shape with out-of-range tile rows, r as a channel slice of a wider buffer."""
    w = bf(rnd((Cout, Cin, 3, 3), 32, 0.05))
    dz = bf(rnd((N, Cout, H, W), 33))
    r = bf(rnd((N, Cin, H, W), 34).clamp_min(0))
    _, wd = ops.pack_conv_weight(w.to(DEV))
    dza = act_from_nchw(ops, dz)
    ra = act_from_nchw(ops, r, ctot=Cin + 64, c0=64)
    dx0, dx1 = ops.Act.empty(N, H, W, Cin, DEV), ops.Act.empty(N, H, W, Cin, DEV)
    ops.conv_fwd(dza, wd, None, dx0, ksize=3)
    partial = torch.full((2 * 160 * 2 * Cin,), float("nan"), dtype=torch.float32, device=DEV)
    rows = ops.conv_dgrad_bnred(dza, wd, dx1, ra, partial, ksize=3)
    torch.cuda.synchronize()
    if rows == 0:
        pytest.skip("shape takes the one-tile kernel (no fused reduction)")
    assert torch.equal(dx0.buf, dx1.buf)
    sums = partial[: rows * 2 * Cin].view(rows, 2, Cin).double().sum(0).cpu()
    dx = dx1.to_nchw_float().double().cpu()
    ref_s, ref_q = dx.sum(dim=(0, 2, 3)), (dx * r.double()).sum(dim=(0, 2, 3))
    scale = float((dx.abs() * r.double()).sum(dim=(0, 2, 3)).max()) + 1e-30
    assert float((sums[0] - ref_s).abs().max()) <= 1e-5 * float(dx.abs().sum(dim=(0, 2, 3)).max())
    assert float((sums[1] - ref_q).abs().max()) <= 1e-5 * scale


@pytest.mark.parametrize("N,Hi,Wi,Cin,Cout", [(2, 8, 8, 128, 64), (1, 16, 16, 256, 128), (3, 6, 10, 128, 64)])
def test_convt_dgrad_with_fused_bn_reduction(ops, N, Hi, Wi, Cin, Cout):
    w = bf(rnd((Cin, Cout, 2, 2), 52, 0.05))
    dy = bf(rnd((N, Cout, 2 * Hi, 2 * Wi), 54))
    r = bf(rnd((N, Cin, Hi, Wi), 55).clamp_min(0))
    _, wd = ops.pack_convt_weight(w.to(DEV))
    dya = act_from_nchw(ops, dy, ctot=2 * Cout, c0=0)
    ra = act_from_nchw(ops, r)
    dx0, dx1 = ops.Act.empty(N, Hi, Wi, Cin, DEV), ops.Act.empty(N, Hi, Wi, Cin, DEV)
    ops.convt_dgrad(dya, wd, dx0)
    partial = torch.full((2 * 160 * 2 * Cin,), float("nan"), dtype=torch.float32, device=DEV)
    rows = ops.convt_dgrad_bnred(dya, wd, dx1, ra, partial)
    torch.cuda.synchronize()
    if rows == 0:
        pytest.skip("shape takes the one-tile kernel (no fused reduction)")
    assert torch.equal(dx0.buf, dx1.buf)
    sums = partial[: rows * 2 * Cin].view(rows, 2, Cin).double().sum(0).cpu()
    dx = dx1.to_nchw_float().double().cpu()
    assert float((sums[0] - dx.sum(dim=(0, 2, 3))).abs().max()) <= 1e-5 * float(dx.abs().sum(dim=(0, 2, 3)).max())
    ref_q = (dx * r.double()).sum(dim=(0, 2, 3))
    assert float((sums[1] - ref_q).abs().max()) <= 1e-5 * (float((dx.abs() * r.double()).sum(dim=(0, 2, 3)).max()) + 1e-30)


def test_conv3x3_large_vs_torch(ops):
    """B x 64 x 128 x 128 against torch fp32 conv on the GPU (size the CPU oracle would not finish in seconds)."""
    N, H, W, Cin, Cout = 4, 128, 128, 64, 128
    x = bf(rnd((N, Cin, H, W), 61)).to(DEV)
    w = bf(rnd((Cout, Cin, 3, 3), 62, 0.05)).to(DEV)
    b = rnd((Cout,), 63).to(DEV)
    xa = ops.Act.from_nchw(x)
    wf, wd = ops.pack_conv_weight(w)
    y = ops.Act.empty(N, H, W, Cout, DEV)
    ops.conv_fwd(xa, wf, b, y, ksize=3, relu=True)
    ref = torch.nn.functional.conv2d(x, w, b, padding=1).clamp_min(0)
    report("conv3x3 large", y.to_nchw_float(), ref)
    dz = bf(rnd((N, Cout, H, W), 64)).to(DEV)
    dza = ops.Act.from_nchw(dz)
    nbytes, _ = ops.wgrad_workspace(N, H, W, Cin, Cout, 9)
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=DEV)
    dw = torch.empty((Cout, Cin, 3, 3), dtype=torch.float32, device=DEV)
    ops.conv3x3_wgrad(xa, dza, ws, dw)
    rdw = torch.nn.grad.conv2d_weight(x, w.shape, dz, padding=1)
    report("wgrad large", dw, rdw, rel=1e-3, abs_frac=1e-4)
    dx = ops.Act.empty(N, H, W, Cin, DEV)
    ops.conv_fwd(dza, wd, None, dx, ksize=3)
    rdx = torch.nn.grad.conv2d_input(x.shape, w, dz, padding=1)
    report("dgrad large", dx.to_nchw_float(), rdx)


def test_pack_weights(ops):
    w = rnd((128, 64, 3, 3), 71).to(DEV)
    wf, wd = ops.pack_conv_weight(w)
    wb = w.to(torch.bfloat16)
    assert torch.equal(wf, wb.permute(2, 3, 0, 1).reshape(9, 128, 64))
    assert torch.equal(wd, wb.flip(2, 3).permute(2, 3, 1, 0).reshape(9, 64, 128))
    wt = rnd((128, 64, 2, 2), 72).to(DEV)
    tf, td = ops.pack_convt_weight(wt)
    tb = wt.to(torch.bfloat16)
    assert torch.equal(tf, tb.permute(2, 3, 1, 0).reshape(4 * 64, 128))
    assert torch.equal(td, tb.permute(2, 3, 0, 1).reshape(4 * 128, 64))


def test_pack_all_matches_per_tensor_packing(ops):
    """one-launch packing of a set of conv / transposed-conv weights == the per-tensor kernels, bit for bit"""
    ws = [("a", rnd((128, 64, 3, 3), 73).to(DEV), False), ("b", rnd((64, 192, 1, 1), 74).to(DEV), False),
          ("t", rnd((128, 64, 2, 2), 75).to(DEV), True), ("c", rnd((40, 72, 3, 3), 76).to(DEV), False)]
    plan = ops.PackPlan(ws, want_dgrad=True)
    plan.run()
    torch.cuda.synchronize()
    for name, w, is_t in ws:
        wf, wd = plan.packed[name]
        rf, rd = ops.pack_convt_weight(w) if is_t else ops.pack_conv_weight(w, want_dgrad=True)
        assert torch.equal(wf, rf) and torch.equal(wd, rd), name
    plan2 = ops.PackPlan(ws[:2], want_dgrad=False)
    plan2.run()
    assert plan2.packed["a"][1] is None and torch.equal(plan2.packed["a"][0], plan.packed["a"][0])


# ------------------------------------------------------------------------------------------------------------
# bandwidth kernels
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,H,W", [(2, 16, 16), (1, 48, 80), (3, 32, 32)])
def test_first_conv_fwd_and_wgrad(ops, N, H, W):
    x = torch.rand((N, 1, H, W), generator=torch.Generator().manual_seed(81))
    w = rnd((64, 1, 3, 3), 82, 0.3)
    b = rnd((64,), 83, 0.3)
    r = ops.Act.empty(N, H, W, 64, DEV)
    rows = ops.c1_rows(N, H, W)
    stats = torch.zeros((rows, 2, 64), dtype=torch.float32, device=DEV)
    ops.conv3x3_c1_fwd(x.to(DEV), w.to(DEV), b.to(DEV), r, relu=True, stats=stats)
    torch.cuda.synchronize()
    ref = O.conv3x3(x.double(), w.double(), b.double()).clamp_min(0)
    got = r.to_nchw_float()
    report("c1 fwd", got, ref, rel=2 ** -8, abs_frac=1e-5)
    g64 = got.double().cpu()
    report("c1 stats", stats.double().sum(0).cpu()[0], g64.sum(dim=(0, 2, 3)), rel=1e-5, abs_frac=1e-5)
    report("c1 stats sq", stats.double().sum(0).cpu()[1], (g64 * g64).sum(dim=(0, 2, 3)), rel=1e-5, abs_frac=1e-5)
    dz = bf(rnd((N, 64, H, W), 84))
    dza = act_from_nchw(ops, dz)
    partial = torch.empty(rows * 64 * 9, dtype=torch.float32, device=DEV)
    scratch = torch.empty(128 * 64 * 9, dtype=torch.float32, device=DEV)
    dw = torch.empty((64, 1, 3, 3), dtype=torch.float32, device=DEV)
    ops.conv3x3_c1_wgrad(x.to(DEV), dza, partial, scratch, dw)
    torch.cuda.synchronize()
    _, rdw, _ = O.conv3x3_bwd(x.double(), w.double(), dz.double())
    report("c1 wgrad", dw, rdw, rel=1e-5, abs_frac=1e-5)


@pytest.mark.parametrize("C,pool", [(64, False), (64, True), (256, True), (1024, False)])
def test_bn_finalize_apply_pool(ops, C, pool):
    N, H, W = 2, 8, 8
    r = bf(rnd((N, C, H, W), 91).clamp_min(0))
    # many exact ties after ReLU: pooling must pick the first maximum
    gamma, beta = rnd((C,), 92), rnd((C,), 93)
    ra = act_from_nchw(ops, r)
    rows = 5
    g64 = r.double()
    part = torch.zeros((rows, 2, C), dtype=torch.float32, device=DEV)
    part[0, 0] = g64.sum(dim=(0, 2, 3)).float().to(DEV)
    part[0, 1] = (g64 * g64).sum(dim=(0, 2, 3)).float().to(DEV)
    f32 = dict(dtype=torch.float32, device=DEV)
    scale, shift, mean, invstd = (torch.empty(C, **f32) for _ in range(4))
    rm, rv = torch.zeros(C, **f32), torch.ones(C, **f32)
    nbt = torch.zeros((), dtype=torch.long, device=DEV)
    scratch = torch.empty(128 * 2 * C, **f32)
    ops.bn_finalize(part, rows, C, N * H * W, gamma.to(DEV), beta.to(DEV), rm, rv, nbt, 0.1, 1e-5, scale, shift, mean,
                    invstd, scratch)
    m, v = O.batchnorm_stats(g64)
    sc, sh, istd = O.batchnorm_affine(m, v, gamma.double(), beta.double())
    report("bn mean", mean, m, rel=1e-5, abs_frac=1e-5)
    report("bn invstd", invstd, istd, rel=1e-4, abs_frac=1e-5)
    report("bn scale", scale, sc, rel=1e-4, abs_frac=1e-5)
    report("bn shift", shift, sh, rel=1e-4, abs_frac=1e-4)
    n = N * H * W
    report("running_mean", rm, 0.1 * m, rel=1e-5, abs_frac=1e-5)
    report("running_var", rv, 0.9 + 0.1 * v * n / (n - 1), rel=1e-4, abs_frac=1e-5)
    assert int(nbt) == 1
    ybuf = torch.zeros((N, H, W, 2 * C), dtype=torch.bfloat16, device=DEV)
    y = ops.Act(ybuf, C, C)
    pooled = ops.Act.empty(N, H // 2, W // 2, C, DEV) if pool else None
    ops.bn_apply(ra, scale, shift, y, pooled)
    torch.cuda.synchronize()
    yref = bf((r * scale.cpu().view(1, -1, 1, 1) + shift.cpu().view(1, -1, 1, 1)))
    report("bn apply", y.to_nchw_float(), yref, rel=2 ** -7, abs_frac=1e-6)
    if pool:
        pref, _ = O.maxpool2x2(y.to_nchw_float().cpu())
        assert torch.equal(pooled.to_nchw_float().cpu(), pref)


@pytest.mark.parametrize("C,pool", [(64, False), (64, True), (512, True), (128, False)])
def test_bn_relu_pool_backward(ops, C, pool):
    N, H, W = 2, 8, 8
    r = bf(rnd((N, C, H, W), 101).clamp_min(0))
    if pool:  # force ties: quantise
        r = bf((r * 2).round() / 2)
    dy = bf(rnd((N, C, H, W), 102))
    gamma, beta = rnd((C,), 103), rnd((C,), 104)
    g64 = r.double()
    m, v = O.batchnorm_stats(g64)
    sc, sh, istd = O.batchnorm_affine(m, v, gamma.double(), beta.double())
    f32 = dict(dtype=torch.float32, device=DEV)
    scale, shift, mean, invstd = sc.float().to(DEV), sh.float().to(DEV), m.float().to(DEV), istd.float().to(DEV)
    dy_total = dy.double()
    dpool_a = None
    if pool:
        dpool = bf(rnd((N, C, H // 2, W // 2), 105))
        dpool_a = act_from_nchw(ops, dpool)
        y = bf((r * scale.cpu().view(1, -1, 1, 1) + shift.cpu().view(1, -1, 1, 1)))
        _, arg = O.maxpool2x2(y)
        dy_total = dy_total + O.maxpool2x2_bwd(dpool.double(), arg, y.shape)
    ra, dya = act_from_nchw(ops, r), act_from_nchw(ops, dy, ctot=2 * C, c0=C)
    dz = ops.Act.empty(N, H, W, C, DEV)
    partial = torch.empty(ops.ew_rows() * 2 * C, **f32)
    scratch = torch.empty(128 * 2 * C, **f32)
    coef = torch.empty(3 * C, **f32)
    dgamma, dbeta, dbias = (torch.empty(C, **f32) for _ in range(3))
    ops.bn_bwd(dya, dpool_a, ra, scale, shift, mean, invstd, gamma.to(DEV), N * H * W, dz, partial, scratch, coef,
               dgamma, dbeta, dbias)
    torch.cuda.synchronize()
    dr, rdg, rdb = O.batchnorm_bwd(dy_total, g64, mean.cpu().double(), invstd.cpu().double(), gamma.double())
    rdz = torch.where(g64 > 0, dr, torch.zeros_like(dr))
    report("dgamma", dgamma, rdg, rel=1e-4, abs_frac=1e-4)
    report("dbeta", dbeta, rdb, rel=1e-4, abs_frac=1e-4)
    report("dz", dz.to_nchw_float(), rdz, rel=2 ** -7, abs_frac=2e-3)
    report("dbias", dbias, dz.to_nchw_float().double().sum(dim=(0, 2, 3)), rel=1e-4, abs_frac=1e-4)


def _close_on_gpu(name, got, ref, rel, abs_frac):
    """report() without the trip to the host (tens of millions of elements)"""
    got, ref = got.double(), ref.double()
    err = (got - ref).abs()
    tol = rel * ref.abs() + abs_frac * float(ref.abs().max()) + 1e-30
    nbad = int((err > tol).sum())
    assert nbad == 0, f"{name}: {nbad}/{err.numel()} out of tolerance, max err {float(err.max()):.4g}"


@pytest.mark.parametrize("pool", [False, True])
def test_bn_streaming_kernels_many_items_per_thread(ops, pool):
    """BN apply / backward at a size where every thread walks its prefetch ring (ew_common.cuh PrefetchRing) several
    times around (the 8x8 cases above give a thread at most one work item); reference = plain torch ops on the GPU."""
    N, H, W, C = 8, 256, 256, 64
    g = torch.Generator(device=DEV).manual_seed(7)
    r = torch.randn((N, H, W, C), generator=g, device=DEV).clamp_min(0).to(torch.bfloat16)
    dyb = torch.zeros((N, H, W, 2 * C), dtype=torch.bfloat16, device=DEV)       # dy is a channel slice (concat buffer)
    dyb[..., C:] = torch.randn((N, H, W, C), generator=g, device=DEV).to(torch.bfloat16)
    gamma = torch.randn(C, generator=g, device=DEV) * 0.5 + 1.0
    beta = torch.randn(C, generator=g, device=DEV) * 0.1
    rd = r.double()
    mean = rd.mean(dim=(0, 1, 2))
    var = rd.var(dim=(0, 1, 2), unbiased=False)
    invstd = 1.0 / torch.sqrt(var + 1e-5)
    scale, shift = (gamma.double() * invstd).float(), (beta.double() - mean * gamma.double() * invstd).float()
    mean32, invstd32 = mean.float(), invstd.float()
    ra, dya = ops.Act(r), ops.Act(dyb, C, C)
    # forward: y (written into a channel slice) and the pooled copy
    ybuf = torch.zeros((N, H, W, 2 * C), dtype=torch.bfloat16, device=DEV)
    y = ops.Act(ybuf, C, C)
    pooled = ops.Act.empty(N, H // 2, W // 2, C, DEV) if pool else None
    ops.bn_apply(ra, scale, shift, y, pooled)
    yref = torch.addcmul(shift, r.float(), scale).to(torch.bfloat16)
    _close_on_gpu("bn apply", y.view(), yref, rel=2 ** -7, abs_frac=1e-6)
    assert bool((ybuf[..., :C] == 0).all())
    dy_total = dyb[..., C:].double()
    dpool = None
    if pool:
        ypl = y.view().float().permute(0, 3, 1, 2)
        pref, idx = torch.nn.functional.max_pool2d(ypl, 2, return_indices=True)
        assert torch.equal(pooled.view().float().permute(0, 3, 1, 2), pref)
        dpool = ops.Act(torch.randn((N, H // 2, W // 2, C), generator=g, device=DEV).to(torch.bfloat16))
        # torch's max_pool2d backward routes to the first maximum in window order, like the kernel
        routed = torch.nn.functional.max_unpool2d(dpool.view().double().permute(0, 3, 1, 2), idx, 2, output_size=(H, W))
        dy_total = dy_total + routed.permute(0, 2, 3, 1)
    # backward
    f32 = dict(dtype=torch.float32, device=DEV)
    dz = ops.Act.empty(N, H, W, C, DEV)
    partial = torch.empty(ops.ew_rows() * 2 * C, **f32)
    scratch = torch.empty(128 * 2 * C, **f32)
    coef = torch.empty(3 * C, **f32)
    dgamma, dbeta, dbias = (torch.empty(C, **f32) for _ in range(3))
    ops.bn_bwd(dya, dpool, ra, scale, shift, mean32, invstd32, gamma, N * H * W, dz, partial, scratch, coef, dgamma, dbeta,
               dbias)
    torch.cuda.synchronize()
    xh = (rd - mean) * invstd
    rdb, rdg = dy_total.sum(dim=(0, 1, 2)), (dy_total * xh).sum(dim=(0, 1, 2))
    cnt = float(N * H * W)
    dr = gamma.double() * invstd * (dy_total - rdb / cnt - xh * rdg / cnt)
    rdz = torch.where(rd > 0, dr, torch.zeros_like(dr))
    _close_on_gpu("dgamma", dgamma, rdg, rel=1e-4, abs_frac=1e-4)
    _close_on_gpu("dbeta", dbeta, rdb, rel=1e-4, abs_frac=1e-4)
    _close_on_gpu("dz", dz.view(), rdz, rel=2 ** -7, abs_frac=2e-3)
    _close_on_gpu("dbias", dbias, dz.view().double().sum(dim=(0, 1, 2)), rel=1e-4, abs_frac=1e-4)


def test_copy_channels_many_items_per_thread(ops):
    """slice copy at a size where every thread runs its unrolled loop several times, into / out of channel slices"""
    N, H, W = 4, 256, 256
    g = torch.Generator(device=DEV).manual_seed(9)
    srcb = torch.randn((N, H, W, 192), generator=g, device=DEV).to(torch.bfloat16)
    dstb = torch.zeros((N, H, W, 128), dtype=torch.bfloat16, device=DEV)
    ops.copy_channels(ops.Act(srcb, 64, 64), ops.Act(dstb, 64, 64))
    assert torch.equal(dstb[..., 64:], srcb[..., 64:128]) and bool((dstb[..., :64] == 0).all())


@pytest.mark.parametrize("O_", [1, 3])
def test_head_fwd_bwd_and_mask(ops, O_):
    N, H, W, C = 2, 16, 16, 64
    r = bf(rnd((N, C, H, W), 111).clamp_min(0))
    scale, shift = rnd((C,), 112), rnd((C,), 113)
    w, b = rnd((O_, C, 1, 1), 114, 0.2), rnd((O_,), 115)
    ra = act_from_nchw(ops, r)
    logits = torch.empty((N, O_, H, W), dtype=torch.float32, device=DEV)
    mask = torch.empty((N, O_, H, W), dtype=torch.uint8, device=DEV)
    ops.head_fwd(ra, scale.to(DEV), shift.to(DEV), w.to(DEV), b.to(DEV), logits, mask)
    torch.cuda.synchronize()
    y = r.double() * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    ref = O.conv1x1(y, w.double(), b.double())
    report("head logits", logits, ref, rel=1e-5, abs_frac=1e-5)
    lg = logits.cpu()
    far = lg.abs() > 1e-6
    assert torch.equal(mask.cpu().bool()[far], O.threshold_mask(lg)[far])
    dl = rnd((N, O_, H, W), 116)
    dy = ops.Act.empty(N, H, W, C, DEV)
    partial = torch.empty(ops.ew_rows() * (O_ * C + O_), dtype=torch.float32, device=DEV)
    scratch = torch.empty(128 * (O_ * C + O_), dtype=torch.float32, device=DEV)
    dwdb = torch.empty(O_ * C + O_, dtype=torch.float32, device=DEV)
    ops.head_bwd(dl.to(DEV), ra, scale.to(DEV), shift.to(DEV), w.to(DEV), dy, partial, scratch, dwdb)
    torch.cuda.synchronize()
    yq = bf(y.float()).double()
    rdx, rdw, rdb = O.conv1x1_bwd(yq, w.double(), dl.double())
    report("head dy", dy.to_nchw_float(), rdx, rel=2 ** -7, abs_frac=1e-4)
    report("head dw", dwdb[: O_ * C], rdw.flatten(), rel=1e-4, abs_frac=1e-4)
    report("head db", dwdb[O_ * C:], rdb, rel=1e-4, abs_frac=1e-4)


@pytest.mark.parametrize("B,HW,soft", [(3, 16 * 16, True), (2, 200 * 200, False), (4, 77, True)])
def test_seg_loss_fwd_bwd(ops, B, HW, soft):
    logits = rnd((B, 1, HW, 1), 121, 3.0)
    t = torch.rand((B, 1, HW, 1), generator=torch.Generator().manual_seed(122))
    t = t if soft else (t > 0.7).float()
    cfg = dict(w_bce=1.0, w_dice=1.0, w_ft=0.5)
    per = HW
    partial = torch.empty(B * ops.loss_chunks(per) * 4, dtype=torch.float32, device=DEV)
    sums = torch.empty(B * 4, dtype=torch.float32, device=DEV)
    out = torch.empty(8, dtype=torch.float32, device=DEV)
    ops.seg_loss_fwd(logits.to(DEV), t.to(DEV), partial, sums, out, **cfg)
    go = torch.tensor([1.7], dtype=torch.float32, device=DEV)
    dl = torch.empty((B, 1, HW, 1), dtype=torch.float32, device=DEV)
    ops.seg_loss_bwd(logits.to(DEV), t.to(DEV), sums, out[4:7], go, dl, **cfg)
    torch.cuda.synchronize()
    ref = O.seg_loss(logits.double(), t.double(), **cfg)
    o = out.cpu().double()
    for i, k in enumerate(["total", "bce", "dice", "ft"]):
        assert abs(float(o[i]) - float(ref[k])) < 2e-6 * max(1.0, abs(float(ref[k]))), (k, float(o[i]), float(ref[k]))
    report("dlogits", dl, 1.7 * ref["dlogits"], rel=1e-4, abs_frac=1e-5)
    report("sums", sums.view(B, 4), ref["sums"], rel=1e-5, abs_frac=1e-6)


@pytest.mark.parametrize("n,offset", [(10007, 0), (10007, 1), (1200003, 0)])
def test_adamw_matches_oracle(ops, n, offset):
    """offset 1: buffers that are not 16-byte aligned (one-by-one path); 1.2 M: every thread loops, plus a 3-element tail"""
    p, g = rnd((n,), 131), rnd((n,), 132)
    view = lambda t: torch.cat([torch.zeros(offset), t]).to(DEV)[offset:]
    pd, m, v = view(p), view(torch.zeros(n)), view(torch.zeros(n))
    pr, mr, vr = p.double(), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    for s in range(1, 4):
        ops.adamw_step(pd, view(g * s), m, v, 1e-3, 0.9, 0.999, 1e-8, 0.01, s, 1.0)
        O.adamw_step(pr, (g * s).double(), mr, vr, 1e-3, step=s)
    torch.cuda.synchronize()
    report("adamw", pd, pr, rel=1e-6, abs_frac=1e-7)


def test_copy_channels_and_reduce_rows(ops):
    src = ops.Act.from_nchw(rnd((2, 64, 8, 8), 141).to(DEV))
    dbuf = torch.zeros((2, 8, 8, 192), dtype=torch.bfloat16, device=DEV)
    dst = ops.Act(dbuf, 64, 64)
    ops.copy_channels(src, dst)
    assert torch.equal(dbuf[..., 64:128], src.buf) and bool((dbuf[..., :64] == 0).all())
    for rows in (1000, 5000):   # direct, and two-level (rows > 1024)
        part = rnd((rows, 96), 142).to(DEV)
        out = torch.empty(96, device=DEV)
        scratch = torch.empty(128 * 96, device=DEV)
        ops.reduce_rows(part, rows, 96, scratch, out)
        report("reduce_rows", out, part.double().sum(0), rel=1e-6, abs_frac=1e-6)


def test_argument_errors(ops):
    from b200seg._lib import B2SError
    x = ops.Act.empty(1, 16, 16, 48, DEV)   # Cin not a multiple of 64
    y = ops.Act.empty(1, 16, 16, 64, DEV)
    w = torch.empty((9, 64, 48), dtype=torch.bfloat16, device=DEV)
    with pytest.raises(B2SError):
        ops.conv_fwd(x, w, None, y)
    with pytest.raises(B2SError):
        ops.bn_apply(ops.Act.empty(1, 15, 16, 64, DEV), torch.empty(64, device=DEV), torch.empty(64, device=DEV),
                     ops.Act.empty(1, 15, 16, 64, DEV), ops.Act.empty(1, 7, 8, 64, DEV))
