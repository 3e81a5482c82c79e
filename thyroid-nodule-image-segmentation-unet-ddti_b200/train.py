"""Data-parallel training step of the UNet hot path without autograd: forward -> fused Dice+BCE -> backward ->
bucketed gradient all-reduce (NCCL over NVLink, overlapped with the remaining backward) -> fused AdamW over flat
fp32 buckets. Replaces the body of the reference's Trainer.train_one_epoch loop between optimizer.zero_grad()
and scaler.update() (utils/trainer.py:81-93) and its nn.DataParallel wrapper (utils/trainer.py:28-30) with one
process per GPU. BatchNorm statistics stay per replica, as under DataParallel.
"""
import math

import torch
import torch.distributed as dist

from . import ops
from .engine import CONVT_INTO, DEC, ENC, UNetEngine


def grad_ready_order():
    """Parameter names in the order UNetEngine.backward finishes their gradients (reverse of forward use)."""
    order = ["final.1.weight", "final.1.bias"]

    def block(name):
        for idx in (3, 0):
            order.extend([f"{name}.{idx + 2}.weight", f"{name}.{idx + 2}.bias", f"{name}.{idx}.bias",
                          f"{name}.{idx}.weight"])

    for l in (0, 1, 2, 3):
        block("final.0" if l == 0 else DEC[l])
        order.extend([f"{CONVT_INTO[l]}.bias", f"{CONVT_INTO[l]}.weight"])
    block("middle.1")
    for l in (3, 2, 1, 0):
        block(ENC[l])
    return order


def cosine_warm_restarts_lr(base_lr, epoch, T_0=20, T_mult=2, eta_min=0.0):
    """torch CosineAnnealingWarmRestarts closed form (utils/trainer.py:42: T_0=20, T_mult=2, eta_min=0)."""
    T_i, t = T_0, epoch
    while t >= T_i:
        t -= T_i
        T_i *= T_mult
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * t / T_i)) / 2


class TrainStep:
    """Owns flat fp32 parameter / gradient / Adam-moment buckets laid out in gradient-ready order."""

    def __init__(self, state_dict, device, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01,
                 bucket_mb=32.0, w_bce=1.0, w_dice=1.0, w_ft=0.0, process_group=None, use_dist=None, engine=None):
        self.device = torch.device(device)
        self.engine = engine if engine is not None else UNetEngine(out_channels=state_dict["final.1.bias"].numel())
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.w_bce, self.w_dice, self.w_ft = w_bce, w_dice, w_ft
        self.step_count = 0
        self.pg = process_group
        self.use_dist = dist.is_available() and dist.is_initialized() if use_dist is None else use_dist
        self.world = dist.get_world_size(self.pg) if self.use_dist else 1
        order = grad_ready_order()
        assert set(order) == {k for k, v in state_dict.items() if v.is_floating_point() and "running" not in k}
        sizes = [state_dict[k].numel() for k in order]
        # 16-byte aligned offsets inside one flat buffer
        offs, total = [], 0
        for n in sizes:
            offs.append(total)
            total += (n + 3) // 4 * 4
        f32 = dict(dtype=torch.float32, device=self.device)
        self.flat_p = torch.zeros(total, **f32)
        self.flat_g = torch.zeros(total, **f32)
        self.flat_m = torch.zeros(total, **f32)
        self.flat_v = torch.zeros(total, **f32)
        self.P, self.G = {}, {}
        for k, o, n in zip(order, offs, sizes):
            shape = state_dict[k].shape
            self.P[k] = self.flat_p[o:o + n].view(shape)
            self.G[k] = self.flat_g[o:o + n].view(shape)
            self.P[k].copy_(state_dict[k])
        for k, v in state_dict.items():
            if k not in self.P:
                self.P[k] = v.detach().clone().to(self.device)
        # buckets: contiguous ranges of the flat gradient, closed when >= bucket_mb
        limit = int(bucket_mb * (1 << 20) / 4)
        self.buckets, self.bucket_of, self.bucket_last = [], {}, {}
        start = 0
        for i, (k, o, n) in enumerate(zip(order, offs, sizes)):
            end = o + (n + 3) // 4 * 4
            bi = len(self.buckets)
            self.bucket_of[k] = bi
            self.bucket_last[bi] = k              # overwritten until the bucket closes: its last-ready parameter
            if end - start >= limit or i == len(order) - 1:
                self.buckets.append((start, end))
                start = end
        self.order = order
        self._works, self._reduce, self._opt, self._side = [], True, None, None
        self._pack_src, self._bucket_packs = None, {}

    # ---- parameter access -------------------------------------------------------------------------------------
    def state_dict(self):
        return {k: v.detach().clone() for k, v in self.P.items()}

    def _on_grad_ready(self, name):
        """Bucket hook of UNetEngine.backward: when the last gradient of a bucket has been enqueued, its all-reduce is
        issued (NCCL's stream, ordered after the compute stream) and — if this step carries a fused optimiser — AdamW
        for the bucket's parameter range is enqueued on a side stream behind the all-reduce. Both overlap the rest of
        backward. Safe because a parameter's last reader in a step is its own layer's backward, which precedes this
        hook (the tensor-core kernels read the packed bf16 copies, refreshed at the start of the next forward)."""
        b = self.bucket_of[name]
        if self.bucket_last[b] != name:
            return
        s, e = self.buckets[b]
        work = None
        if self.world > 1 and self._reduce:
            work = dist.all_reduce(self.flat_g[s:e], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self._works.append(work)
        if self._opt is not None:
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            if work is None:
                self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                if work is not None:
                    work.wait()                    # the side stream waits for the collective; the host does not block
                self._opt(s, e)
                # ... and the bf16 GEMM operands of the bucket's conv weights are re-packed from the updated fp32
                # masters right away (also overlapped), so the next forward starts without a packing pass
                pk = self._bucket_pack(b)
                if pk is not None:
                    pk.run()

    def _bucket_pack(self, b):
        """Sub-plan of the engine's weight pack covering bucket b (same output buffers), built on first use."""
        plan = self.engine.packed_plan(self.P, need_dgrad=True)
        if self._pack_src is not plan:              # the engine re-allocated its operands: rebuild the sub-plans
            self._pack_src, self._bucket_packs = plan, {}
        if b not in self._bucket_packs:
            names = [k for k in self.order if self.bucket_of[k] == b and k in plan.packed]
            items = [(k, self.P[k].detach(), k.split(".")[0] in ("middle", "decoder3", "decoder2", "decoder1")
                      and self.P[k].shape[2] == 2) for k in names]
            self._bucket_packs[b] = ops.PackPlan(items, want_dgrad=True, outputs=plan.packed) if items else None
        return self._bucket_packs[b]

    # ---- one optimisation step ------------------------------------------------------------------------------------
    def forward_backward(self, x, t, reduce=True, opt=None):
        """x [B,1,H,W] fp32, t [B,1,H,W] fp32 on the device. Returns the 8-float loss vector (device tensor).
        reduce=False leaves the gradients un-reduced (the caller all-reduces flat_g itself).
        opt: None, or a callable (start, end) applying the optimiser to that range of the flat buffers; it is launched
        per bucket, overlapped with the remaining backward (see _on_grad_ready)."""
        eng = self.engine
        _, pl = eng.forward(self.P, x, train=True)
        out = eng.loss(pl, t, w_bce=self.w_bce, w_dice=self.w_dice, w_ft=self.w_ft)
        ft_tot, w_ft = None, self.w_ft
        if self.w_ft != 0.0 and self.world > 1:
            # FocalTversky (models/loss.py:34-46) is a function of BATCH-GLOBAL sums: all-reduce {TP, sum p, sum t}
            # before the gradient. Its gradient is a sum over ranks, not a mean, so it is pre-multiplied by the world
            # size that the optimiser's 1/world gradient scale divides out again. out[3] stays the rank-local value.
            ft_tot = out[4:7].clone()
            dist.all_reduce(ft_tot, op=dist.ReduceOp.SUM, group=self.pg)
            w_ft = self.w_ft * self.world
        dl = eng.loss_backward(pl, t, ft_tot=ft_tot, w_bce=self.w_bce, w_dice=self.w_dice, w_ft=w_ft)
        self._works, self._reduce, self._opt = [], reduce, opt
        hook = self._on_grad_ready if (opt is not None or (reduce and self.world > 1)) else None
        eng.backward(self.P, pl, dl, self.G, on_grad_ready=hook)
        for w in self._works:
            w.wait()
        if opt is not None and self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
        self._opt = None
        return out

    def optimizer_step(self, lr=None):
        self.step_count += 1
        ops.adamw_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, lr if lr is not None else self.lr,
                       self.betas[0], self.betas[1], self.eps, self.wd, self.step_count, 1.0 / self.world)
        self.engine.invalidate_packed()      # the kernel updated flat_p behind torch's version counters

    def step(self, x, t, lr=None):
        """Host-launched step: forward, loss, backward with the bucketed all-reduce and per-bucket AdamW overlapped."""
        self.step_count += 1
        lr_ = lr if lr is not None else self.lr
        b1, b2, step, scale = self.betas[0], self.betas[1], self.step_count, 1.0 / self.world

        def opt(s, e):
            ops.adamw_step(self.flat_p[s:e], self.flat_g[s:e], self.flat_m[s:e], self.flat_v[s:e], lr_, b1, b2, self.eps,
                           self.wd, step, scale)
        out = self.forward_backward(x, t, opt=opt)   # the bucket hooks also refreshed the packed bf16 operands
        return out

    # ---- the same step as ONE CUDA graph -----------------------------------------------------------------------
    # ~230 kernel launches per step cost more host time than the GPU needs to run them at batch 64; the graph removes
    # the host from the loop. Everything that changes between steps lives in device memory: the batch (copied into
    # static input buffers), BatchNorm counters (updated in-kernel), and the AdamW scalars (lr, bias corrections),
    # an 8-float device vector refreshed before each replay by an asynchronous copy from a small ring of pinned host
    # vectors (a slot is reused only after the copy that read it has completed, so the host may run ahead).
    def _write_hyper(self, lr):
        self.step_count += 1
        # betas rounded to fp32 first: the same double-precision bias corrections b2s_adamw_step derives from its
        # float arguments, so the graphed and the host-launched step stay bit-identical
        b1, b2 = (torch.tensor(b, dtype=torch.float32).item() for b in self.betas)
        i = self.step_count % len(self._hyper_ring)
        h, ev = self._hyper_ring[i]
        ev.synchronize()
        h[0], h[1], h[2], h[3], h[4] = lr if lr is not None else self.lr, b1, b2, self.eps, self.wd
        h[5], h[6], h[7] = 1.0 - b1 ** self.step_count, math.sqrt(1.0 - b2 ** self.step_count), 1.0 / self.world
        self._hyper_dev.copy_(h, non_blocking=True)
        ev.record()

    def _graph_body(self):
        # The whole optimisation step is ONE graph on every world size. Data parallel: each bucket's ncclAllReduce is
        # captured on NCCL's own stream (forked from the compute stream by the event recorded after the bucket's last
        # weight-gradient kernel, joined before AdamW), so inside the replay the collectives overlap the remaining
        # backward exactly as in the host-launched step; only the last, small bucket and the join are exposed.
        # AdamW runs per bucket on a side stream behind the bucket's all-reduce (world 1: behind its last gradient).
        if self.world == 1 or self.graph_comm:
            def opt(s, e):
                ops.adamw_step_dev(self.flat_p[s:e], self.flat_g[s:e], self.flat_m[s:e], self.flat_v[s:e], self._hyper_dev)
            self.forward_backward(self._gx, self._gt, reduce=self.graph_comm, opt=opt)
        else:
            self.forward_backward(self._gx, self._gt, reduce=False)

    def capture(self, x, t, graph_comm=True):
        """Captures forward + loss + backward + bucketed all-reduce + AdamW for inputs shaped like x, t.
        Runs two eager steps first (lazy initialisation inside the library, NCCL communicator set-up); they are real
        optimisation steps. graph_comm=False keeps NCCL out of the capture (one all-reduce over the flat gradient and
        AdamW launched after the replay); it is also the automatic second attempt when capturing the collectives fails
        on any rank. Returns True when the graph is in use, False when capture failed (eager path stays in effect)."""
        self._gx, self._gt = torch.empty_like(x), torch.empty_like(t)
        self._hyper_dev = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._hyper_ring = [(torch.zeros(8, dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(8)]
        self._graph = None
        for _ in range(2):
            self.step(x, t)
        torch.cuda.synchronize()
        self.graph_comm = bool(graph_comm) and self.world > 1
        ok = self._agree_on_graph(self._try_capture(x, t))
        if not ok and self.graph_comm:
            first = getattr(self, "capture_error", "another rank failed")
            self.graph_comm = False
            ok = self._agree_on_graph(self._try_capture(x, t))
            self.capture_error = f"in-graph NCCL capture failed ({first}); collectives launched after the replay"
        return ok

    def release_graph(self):
        """Drops the captured graph. Call before torch.distributed.destroy_process_group(): NCCL cannot tear down a
        communicator while a live CUDA graph still holds its captured collectives (the destroy call blocks)."""
        self._graph = None
        torch.cuda.synchronize(self.device)

    def _try_capture(self, x, t):
        from . import _lib
        g = torch.cuda.CUDAGraph()
        step_before = self.step_count
        try:
            self._gx.copy_(x); self._gt.copy_(t)
            self._write_hyper(None)
            l0 = _lib.launch_count()
            # thread_local: NCCL's watchdog / heartbeat threads may query events while this thread captures
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._graph_body()
            self.graph_launches = _lib.launch_count() - l0
            self._graph = g
        except Exception as e:   # capture is an optimisation: report and keep the eager path
            self.capture_error = f"{type(e).__name__}: {e}"
            self._graph = None
            torch.cuda.synchronize()
        # the capture itself did not execute the step (and step_count was advanced for it): undo the count
        self.step_count = step_before
        return self._graph is not None

    def _agree_on_graph(self, ok):
        """Under data parallelism every rank must take the same path (graphed or eager, collectives in or out)."""
        if self.world > 1:
            flag = torch.tensor([1.0 if ok else 0.0], device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.pg)
            ok = bool(flag.item() > 0.5)
            if not ok:
                self._graph = None
        return ok

    def step_graphed(self, x, t, lr=None):
        """Replays the captured step on a new batch; returns the 8-float loss vector (device tensor, static buffer)."""
        if getattr(self, "_graph", None) is None:
            return self.step(x, t, lr)
        self._gx.copy_(x, non_blocking=True)
        self._gt.copy_(t, non_blocking=True)
        self._write_hyper(lr)
        self._graph.replay()
        if self.world > 1 and not self.graph_comm:
            self.engine.invalidate_packed()
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
            ops.adamw_step_dev(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self._hyper_dev)
        key = (x.shape[0], x.shape[2], x.shape[3], str(x.device))
        return self.engine.plans[key].loss_out


# ---- verification of the data-parallel semantics ---------------------------------------------------------------------
def emulate_replicas(state_dict, device, shards, lrs, engine=None, **kw):
    """The reference's data-parallel semantics (nn.DataParallel, utils/trainer.py:28-30) evaluated on ONE GPU: every
    replica r runs forward + loss + backward on shards[r] with its OWN BatchNorm batch statistics and running buffers,
    the parameter gradients are summed over the replicas in rank order, and every replica applies the same AdamW update
    to the mean gradient, one step per entry of `lrs`. Returns the list of per-replica TrainStep objects (replica 0 holds the buffers that would be
    checkpointed). Used by bench.py's post-run check and tests/run_ddp_check.py to verify that N NCCL ranks x b images
    reproduce this chunked single-GPU run."""
    n = len(shards)
    engine = engine if engine is not None else UNetEngine(out_channels=state_dict["final.1.bias"].numel())
    reps = [TrainStep({k: v.clone() for k, v in state_dict.items()}, device, lr=lrs[0] or 0.0, use_dist=False, engine=engine,
                      **kw) for _ in range(n)]
    for rep in reps:
        rep.world = n                      # AdamW consumes g / n, exactly what the ranks do
    for lr in lrs:
        for rep, (x, t) in zip(reps, shards):
            rep.forward_backward(x, t, reduce=False)
        total = reps[0].flat_g.clone()
        for rep in reps[1:]:
            total += rep.flat_g
        for rep in reps:
            rep.flat_g.copy_(total)
            if lr is not None:             # None: forward + backward only (gradient probe), no optimiser step
                rep.optimizer_step(lr)
    return reps


def ranks_agree(flat, group=None):
    """True when every rank holds bit-identical `flat` (compared through their int32 views: NaN-safe, sign-of-zero safe)."""
    world = dist.get_world_size(group)
    mine = flat.view(torch.int32)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    return all(bool(torch.equal(gathered[0], g)) for g in gathered[1:])
