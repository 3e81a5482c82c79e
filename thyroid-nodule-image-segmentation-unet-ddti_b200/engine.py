"""Host-side plan of the UNet hot path: owns the NHWC bf16 activation buffers, the packed bf16 weights and the
launch order of the libb2s kernels for forward (train / eval), loss and backward.

Graph mirrored: reference models/model.py:53-73 (forward), autograd of the same for backward. Data layout in HBM
(DESIGN.md §3): per level l (H/2^l x W/2^l, C_l = 64*2^l)
    cat[l]   [N,H_l,W_l,2*C_l]  = [ up-conv output | encoder skip ]   (torch.cat is never materialised)
    r/y      post-ReLU conv outputs (saved) and BatchNorm outputs (operands of the next conv / wgrad)
    pooled[l] max-pooled skip, input of level l+1
"""
import torch

from . import ops
from .ops import Act

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

ENC = ["encoder1", "encoder2", "encoder3", "encoder4"]
DEC = {3: "decoder3.0", 2: "decoder2.0", 1: "decoder1.0"}          # block consuming cat[l]
CONVT_INTO = {3: "middle.2", 2: "decoder3.1", 1: "decoder2.1", 0: "decoder1.1"}  # up-conv writing cat[l]


def conv_stages():
    """(prefix, idx, Cin, Cout, level) of the 18 conv3x3+ReLU+BN stages in forward order."""
    st = []
    cin = None
    for l, name in enumerate(ENC):
        c = 64 << l
        st.append((name, 0, cin if l else None, c, l))
        st.append((name, 3, c, c, l))
        cin = c
    st.append(("middle.1", 0, 512, 1024, 4))
    st.append(("middle.1", 3, 1024, 1024, 4))
    for l in (3, 2, 1):
        c = 64 << l
        st.append((DEC[l], 0, 2 * c, c, l))
        st.append((DEC[l], 3, c, c, l))
    st.append(("final.0", 0, 128, 64, 0))
    st.append(("final.0", 3, 64, 64, 0))
    return st


class _Stage:
    """Per conv stage: saved tensors and BN vectors."""
    __slots__ = ("name", "idx", "cin", "cout", "level", "x", "r", "y", "scale", "shift", "mean", "invstd", "wf", "wd")


class UNetPlan:
    """Buffers for one (N,H,W) on one device."""

    def __init__(self, N, H, W, device, in_channels=1, out_channels=1, train_buffers=True):
        if H % 16 or W % 16:
            raise RuntimeError(f"Sizes of tensors must match: H={H}, W={W} must be multiples of 16 "
                               "(reference models/model.py:64 torch.cat)")
        if not 1 <= in_channels <= 64:
            raise NotImplementedError("the B200 path implements 1 <= in_channels <= 64")
        # in_channels > 1: the image is converted once to NHWC bf16 zero-padded to 64 channels and encoder1.0 takes the
        # tensor-core conv with its weight zero-padded to 64 inputs (in_channels = 1 keeps the fp32 Cin = 1 kernels)
        self.Cin = in_channels
        self.x_act = Act.empty(N, H, W, 64, device) if in_channels > 1 else None
        self.N, self.H, self.W, self.device = N, H, W, device
        self.O = out_channels
        f32 = dict(dtype=torch.float32, device=device)
        dims = [(H >> l, W >> l, 64 << l) for l in range(5)]
        self.dims = dims
        A = lambda l, C: Act.empty(N, dims[l][0], dims[l][1], C, device)
        self.cat = [A(l, 2 * dims[l][2]) for l in range(4)]
        self.pooled = [A(l + 1, dims[l][2]) for l in range(4)]
        self.stages = {}
        for (name, idx, cin, cout, l) in conv_stages():
            s = _Stage()
            if cin is None and in_channels > 1:
                cin = 64
            s.name, s.idx, s.cin, s.cout, s.level = name, idx, cin, cout, l
            s.r = A(l, cout)
            s.y = None
            for k in ("scale", "shift", "mean", "invstd"):
                setattr(s, k, torch.empty(cout, **f32))
            s.wf = s.wd = None
            self.stages[(name, idx)] = s
        # BN outputs: first stage of each block -> dense y; second stage -> concat slice (encoders) or dense
        for (name, idx, cin, cout, l) in conv_stages():
            s = self.stages[(name, idx)]
            if idx == 0:
                s.y = A(l, cout)
            elif name in ENC:
                s.y = self.cat[l].slice(cout, cout)
            elif name != "final.0":
                s.y = A(l, cout)
        self.logits = torch.empty((N, out_channels, H, W), **f32)
        self.mask = torch.empty((N, out_channels, H, W), dtype=torch.uint8, device=device)
        # partial / scratch buffers
        self.stats_partial = torch.empty(max(2 * 160 * 2 * 2048, ops.c1_rows(N, H, W) * 2 * 64), **f32)
        self.ew_partial = torch.empty(ops.ew_rows() * 2 * 1024, **f32)
        self.c1_partial = torch.empty(ops.c1_rows(N, H, W) * 64 * 9, **f32)
        self.scratch = torch.empty(128 * 2 * 2048, **f32)
        self.coef = torch.empty(3 * 1024, **f32)
        self.tmp_vec = torch.empty(2 * 1024, **f32)
        # loss
        per = out_channels * H * W
        self.loss_partial = torch.empty(N * ops.loss_chunks(per) * 4, **f32)
        self.loss_sums = torch.empty(N * 4, **f32)
        self.loss_out = torch.empty(8, **f32)
        self.generation = 0
        self.train_ready = False
        if train_buffers:
            self._alloc_train()

    def _alloc_train(self):
        if self.train_ready:
            return
        N, dims, device = self.N, self.dims, self.device
        A = lambda l, C: Act.empty(N, dims[l][0], dims[l][1], C, device)
        self.dlogits = torch.empty_like(self.logits)
        self.dcat = [A(l, 2 * dims[l][2]) for l in range(4)]
        self.ga = [A(l, dims[l][2]) for l in range(5)]
        self.gb = [A(l, dims[l][2]) for l in range(5)]
        self.dpool = [A(l + 1, dims[l][2]) for l in range(4)]
        ws_bytes = 0
        for (name, idx, cin, cout, l) in conv_stages():
            if cin is None and self.Cin > 1:
                cin = 64
            if cin is None:
                continue
            nb, _ = ops.wgrad_workspace(N, dims[l][0], dims[l][1], cin, cout, 9)
            ws_bytes = max(ws_bytes, nb)
        for l in range(4):
            cin, cout = 2 * dims[l][2], dims[l][2]
            nb, _ = ops.wgrad_workspace(N, dims[l + 1][0], dims[l + 1][1], cin, cout, 4)
            ws_bytes = max(ws_bytes, nb)
        self.wgrad_ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=device)
        # {sum dy, sum dy*r} partials of the fused BatchNorm-backward reduction (<= 2 rows per SM, <= 1024 channels)
        self.red_partial = torch.empty(2 * 160 * 2 * 1024, dtype=torch.float32, device=device)
        self.train_ready = True


class UNetEngine:
    """Runs the reference UNet graph on libb2s kernels for a dict of parameters/buffers (reference state_dict keys)."""

    def __init__(self, out_channels=1):
        self.O = out_channels
        self.plans = {}
        self._pack_state = {}
        # Optional (B2S_FUSE_BN_REDUCE=1, default off): BatchNorm-backward sums emitted by the dgrad / transposed-conv dgrad
        # launch that produces dy (13 of 18 stages) instead of a separate pass over dy and r. Measured on a B200 at batch
        # 64 @256^2: the reduce passes shrink by 0.92 ms per step, but the tensor-core epilogues that take them over grow
        # by 2.4 ms (32 extra 128-byte row loads per warp and sub-tile on the critical path of kernels with one
        # accumulator set or one or two items per CTA): 25.2 vs 23.7 ms per step. Kept for the A/B and its kernel tests.
        import os
        self.fuse_bn_reduce = os.environ.get("B2S_FUSE_BN_REDUCE", "0") == "1"

    # ---- plumbing -----------------------------------------------------------------------------------------
    def plan(self, N, H, W, device, train, in_channels=1):
        key = (N, H, W, str(device)) if in_channels == 1 else (N, H, W, str(device), in_channels)
        p = self.plans.get(key)
        if p is None:
            p = UNetPlan(N, H, W, device, in_channels=in_channels, out_channels=self.O, train_buffers=train)
            self.plans[key] = p
        if train:
            p._alloc_train()
        return p

    def packed_plan(self, P, need_dgrad=True):
        """Packs now if needed and returns the PackPlan whose buffers the next forwards of P will read."""
        self._pack_weights(P, need_dgrad=need_dgrad)
        names = [k for k in P if k.endswith(".weight") and P[k].dim() == 4 and k != "encoder1.0.weight"
                 and k != "final.1.weight"]
        return self._pack_state[(str(P[names[0]].device), bool(need_dgrad))]["plan"]

    def invalidate_packed(self):
        """Call after parameters were updated outside torch (e.g. by b2s_adamw_step, which does not bump tensor
        version counters): forces the next forward to re-pack the bf16 operands."""
        for st in self._pack_state.values():
            st["versions"] = None

    def _pack_weights(self, P, need_dgrad, cache=True):
        """fp32 parameters -> bf16 GEMM operands, all tensors in one launch; skipped while the parameters' version
        counters and addresses are unchanged. State is kept per device: nn.DataParallel (utils/trainer.py:30) runs
        replicas of one module — sharing this engine object — concurrently, one thread per GPU."""
        names = [k for k in P if k.endswith(".weight") and P[k].dim() == 4 and k != "encoder1.0.weight"
                 and k != "final.1.weight"]
        # one state per (device, with / without the dgrad operands): an eval forward between two training steps must
        # not replace the buffers a captured training graph or TrainStep's per-bucket re-pack points at
        dev = (str(P[names[0]].device), bool(need_dgrad))
        st = self._pack_state.setdefault(dev, {"versions": None, "layout": None, "plan": None})
        # (identity, address, version): a freed tensor's address can be handed to a new tensor at version 0, so the
        # owning tensor object is part of the key (the plan keeps the tensors it packed alive, ids cannot be reused)
        versions = tuple((k, id(P[k]), P[k]._version, P[k].data_ptr(), need_dgrad) for k in names)
        if cache and versions == st["versions"]:
            return st["plan"].packed
        layout = tuple((k, id(P[k]), P[k].data_ptr(), need_dgrad) for k in names)
        if not cache or st["layout"] != layout:
            items = []
            for k in names:
                is_convt = k.split(".")[0] in ("middle", "decoder3", "decoder2", "decoder1") and P[k].shape[2] == 2
                items.append((k, P[k].detach(), is_convt))
            st["plan"] = ops.PackPlan(items, want_dgrad=need_dgrad)
            st["layout"] = layout
            st["owners"] = [P[k] for k in names]
        st["plan"].run()
        st["versions"] = versions
        return st["plan"].packed

    # ---- forward ----------------------------------------------------------------------------------------------
    def forward(self, P, x, train, want_mask=False, need_backward=None, cache_packed=True):
        """x [N,1,H,W] fp32 CUDA. Returns (logits fp32 [N,O,H,W], plan). P: name -> tensor (params and BN buffers)."""
        if not x.is_cuda:
            raise ops._lib.B2SError("UNetEngine.forward needs a CUDA tensor: the B200 path has no CPU fallback")
        cin = P["encoder1.0.weight"].shape[1]
        if x.dim() != 4 or x.shape[1] != cin:
            raise RuntimeError(f"expected input [N,{cin},H,W], got {tuple(x.shape)}")
        need_backward = train if need_backward is None else need_backward
        N, _, H, W = x.shape
        pl = self.plan(N, H, W, x.device, need_backward, in_channels=cin)
        pl.generation += 1
        x = x.contiguous().float()
        pl.x = x
        # nn.DataParallel replicas get fresh broadcast copies of the parameters every forward: never reuse a pack
        pl.packed = self._pack_weights(P, need_dgrad=need_backward, cache=cache_packed)
        if cin > 1:
            ops.image_to_nhwc(x, 64, out=pl.x_act)
            wpad = torch.nn.functional.pad(P["encoder1.0.weight"].detach(), (0, 0, 0, 0, 0, 64 - cin))
            pl.packed = dict(pl.packed)
            pl.packed["encoder1.0.weight"] = (ops.pack_conv_weight(wpad, want_dgrad=False)[0], None)
        count = lambda l: float(N * pl.dims[l][0] * pl.dims[l][1])

        def stage(name, idx, xin, pooled=None):
            s = pl.stages[(name, idx)]
            s.x = xin
            bn = f"{name}.{idx + 2}"
            if not train and xin is not None and s.y is not None:
                # inference: running-statistics BatchNorm applied in the conv epilogue, no separate BN pass
                ops.bn_eval_affine(P[f"{bn}.weight"], P[f"{bn}.bias"], P[f"{bn}.running_mean"],
                                   P[f"{bn}.running_var"], BN_EPS, s.scale, s.shift)
                wf, _ = pl.packed[f"{name}.{idx}.weight"]
                ops.conv_fwd_affine(xin, wf, P[f"{name}.{idx}.bias"], s.scale, s.shift, s.y, ksize=3, relu=True)
                if pooled is not None:
                    ops.maxpool2x2(s.y, pooled)
                return s
            if xin is None and not train and s.y is not None and s.y.c0 == 0 and s.y.C == s.y.cstride:
                ops.bn_eval_affine(P[f"{bn}.weight"], P[f"{bn}.bias"], P[f"{bn}.running_mean"],
                                   P[f"{bn}.running_var"], BN_EPS, s.scale, s.shift)
                ops.conv3x3_c1_fwd_affine(x, P[f"{name}.{idx}.weight"], P[f"{name}.{idx}.bias"], s.scale, s.shift, s.y)
                return s
            if xin is None:  # first conv on the fp32 image
                ops.conv3x3_c1_fwd(x, P[f"{name}.{idx}.weight"], P[f"{name}.{idx}.bias"], s.r, relu=True,
                                   stats=pl.stats_partial if train else None)
                rows = ops.c1_rows(N, H, W)
            else:
                wf, _ = pl.packed[f"{name}.{idx}.weight"]
                ops.conv_fwd(xin, wf, P[f"{name}.{idx}.bias"], s.r, ksize=3, relu=True,
                             stats=pl.stats_partial if train else None)
                rows = ops.conv_stats_rows(N, s.r.H, s.r.W, s.cout)
            if train:
                ops.bn_finalize(pl.stats_partial, rows, s.cout, count(s.level), P[f"{bn}.weight"], P[f"{bn}.bias"],
                                P[f"{bn}.running_mean"], P[f"{bn}.running_var"], P[f"{bn}.num_batches_tracked"],
                                BN_MOMENTUM, BN_EPS, s.scale, s.shift, s.mean, s.invstd, pl.scratch)
            else:
                ops.bn_eval_affine(P[f"{bn}.weight"], P[f"{bn}.bias"], P[f"{bn}.running_mean"],
                                   P[f"{bn}.running_var"], BN_EPS, s.scale, s.shift)
            if s.y is not None:
                ops.bn_apply(s.r, s.scale, s.shift, s.y, pooled)
            return s

        def block(name, xin, pooled=None):
            s0 = stage(name, 0, xin)
            return stage(name, 3, s0.y, pooled)

        cur = pl.x_act           # None for in_channels = 1: encoder1.0 then runs the fp32 Cin = 1 kernels on x
        for l, name in enumerate(ENC):
            block(name, cur, pl.pooled[l])
            cur = pl.pooled[l]
        s = block("middle.1", cur)
        for l in (3, 2, 1, 0):
            ct = CONVT_INTO[l]
            wf, _ = pl.packed[f"{ct}.weight"]
            C = pl.dims[l][2]
            ops.convt_fwd(s.y, wf, P[f"{ct}.bias"], pl.cat[l].slice(0, C))
            s = block(DEC[l] if l else "final.0", pl.cat[l])
        ops.head_fwd(s.r, s.scale, s.shift, P["final.1.weight"], P["final.1.bias"], pl.logits,
                     pl.mask if want_mask else None)
        return pl.logits, pl

    # ---- loss -------------------------------------------------------------------------------------------------
    def loss(self, pl, targets, w_bce=1.0, w_dice=1.0, w_ft=0.0, **kw):
        """Fused Dice+BCE(+FocalTversky) on pl.logits; returns the 8-float result vector (total, bce, dice, ft, ...)."""
        ops.seg_loss_fwd(pl.logits, targets, pl.loss_partial, pl.loss_sums, pl.loss_out, w_bce=w_bce, w_dice=w_dice,
                         w_ft=w_ft, **kw)
        return pl.loss_out

    def loss_backward(self, pl, targets, grad_out=None, ft_tot=None, w_bce=1.0, w_dice=1.0, w_ft=0.0, **kw):
        ops.seg_loss_bwd(pl.logits, targets, pl.loss_sums, ft_tot, grad_out, pl.dlogits, w_bce=w_bce, w_dice=w_dice,
                         w_ft=w_ft, **kw)
        return pl.dlogits

    # ---- backward ---------------------------------------------------------------------------------------------
    def backward(self, P, pl, dlogits, G, on_grad_ready=None):
        """Writes the gradient of every parameter into G[name] (fp32 tensors, parameter shapes).
        on_grad_ready(name) is called as soon as G[name] has been enqueued (DDP bucket hook)."""
        N = pl.N
        dlogits = dlogits.contiguous().float()
        ready = on_grad_ready or (lambda name: None)
        count = lambda l: float(N * pl.dims[l][0] * pl.dims[l][1])

        def stage_bwd(name, idx, dy, dx_out, dpool=None, dx_stats=False, pre=None, red_for=None):
            """BN+ReLU backward -> dz; conv wgrad; conv dgrad into dx_out (None for the image conv).
            pre: (partial, rows) when the launch that produced dy already emitted the BatchNorm-backward sums.
            red_for: the stage whose BatchNorm consumes dx_out directly: this stage's dgrad then emits those sums
            (fused epilogue) and the function returns them for that stage's `pre`."""
            s = pl.stages[(name, idx)]
            bn = f"{name}.{idx + 2}"
            l = s.level
            dz = pl.gb[l]
            ops.bn_bwd(dy, dpool, s.r, s.scale, s.shift, s.mean, s.invstd, P[f"{bn}.weight"], count(l), dz,
                       pl.ew_partial, pl.scratch, pl.coef, G[f"{bn}.weight"], G[f"{bn}.bias"],
                       G[f"{name}.{idx}.bias"], pre=pre)
            ready(f"{bn}.weight"); ready(f"{bn}.bias"); ready(f"{name}.{idx}.bias")
            wname = f"{name}.{idx}.weight"
            if s.x is None:
                ops.conv3x3_c1_wgrad(pl.x, dz, pl.c1_partial, pl.scratch, G[wname])
                ready(wname)
                return None
            if dx_out is None:      # encoder1.0 of a multi-channel image: padded weight gradient, no input gradient
                tmp = torch.empty((s.cout, 64, 3, 3), dtype=torch.float32, device=dz.buf.device)
                ops.conv3x3_wgrad(s.x, dz, pl.wgrad_ws, tmp)
                G[wname].copy_(tmp[:, :pl.Cin])
                ready(wname)
                return None
            ops.conv3x3_wgrad(s.x, dz, pl.wgrad_ws, G[wname])
            _, wd = pl.packed[wname]
            out_pre = None
            if red_for is not None and self.fuse_bn_reduce:
                rows = ops.conv_dgrad_bnred(dz, wd, dx_out, pl.stages[red_for].r, pl.red_partial, ksize=3)
                out_pre = (pl.red_partial, rows) if rows else None
            if out_pre is None:
                ops.conv_fwd(dz, wd, None, dx_out, ksize=3, relu=False, stats=pl.stats_partial if dx_stats else None)
            # after the input-gradient launch: a bucket hook may re-pack this layer's bf16 operands on a side stream
            ready(wname)
            return out_pre

        def block_bwd(name, dy1, dx_out, dpool=None, dx_stats=False, pre=None):
            """pre: fused BatchNorm-backward sums for the block's SECOND stage (dy1 came out of a transposed-conv dgrad)"""
            l = pl.stages[(name, 0)].level
            pre0 = stage_bwd(name, 3, dy1, pl.ga[l], dpool=dpool, pre=pre, red_for=(name, 0))
            stage_bwd(name, 0, pl.ga[l], dx_out, dx_stats=dx_stats, pre=pre0)

        # head (final.1) -> dy of final.0's second BN
        s = pl.stages[("final.0", 3)]
        dwdb = pl.tmp_vec[: self.O * 64 + self.O]
        ops.head_bwd(dlogits, s.r, s.scale, s.shift, P["final.1.weight"], pl.ga[0], pl.ew_partial, pl.scratch, dwdb)
        G["final.1.weight"].view(-1).copy_(dwdb[: self.O * 64])
        G["final.1.bias"].copy_(dwdb[self.O * 64:])
        ready("final.1.weight"); ready("final.1.bias")

        dy = pl.ga[0]
        pre = None
        for l in (0, 1, 2, 3):
            name = "final.0" if l == 0 else DEC[l]
            C = pl.dims[l][2]
            block_bwd(name, dy, pl.dcat[l], dx_stats=True, pre=pre)
            # transposed conv writing cat[l][:, :C]: bias grad = column sums of dcat[l][..., :C] (dgrad epilogue)
            ct = CONVT_INTO[l]
            rows = ops.conv_stats_rows(N, pl.dims[l][0], pl.dims[l][1], 2 * C)
            ops.reduce_rows(pl.stats_partial, rows, 2 * 2 * C, pl.scratch, pl.tmp_vec)
            G[f"{ct}.bias"].copy_(pl.tmp_vec[:C])
            ready(f"{ct}.bias")
            up_in = pl.stages[(DEC[l + 1] if l < 3 else "middle.1", 3)].y      # input of the transposed conv
            dY = pl.dcat[l].slice(0, C)
            ops.convt_wgrad(up_in, dY, pl.wgrad_ws, G[f"{ct}.weight"])
            _, wd = pl.packed[f"{ct}.weight"]
            # ga[l+1] is the gradient of the BatchNorm output of the next block's second stage: fused reduction
            nxt = (DEC[l + 1] if l < 3 else "middle.1", 3)
            rows = ops.convt_dgrad_bnred(dY, wd, pl.ga[l + 1], pl.stages[nxt].r, pl.red_partial) \
                if self.fuse_bn_reduce else 0
            pre = (pl.red_partial, rows) if rows else None
            if pre is None:
                ops.convt_dgrad(dY, wd, pl.ga[l + 1])
            ready(f"{ct}.weight")
            dy = pl.ga[l + 1]
        block_bwd("middle.1", dy, pl.dpool[3], pre=pre)
        for l in (3, 2, 1, 0):
            C = pl.dims[l][2]
            skip_grad = pl.dcat[l].slice(C, C)
            dx_out = pl.dpool[l - 1] if l > 0 else None
            block_bwd(ENC[l], skip_grad, dx_out, dpool=pl.dpool[l])
        return G
