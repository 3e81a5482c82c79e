"""Drop-in replacement for the reference's ``models/model.py``: ``UNet(in_channels=1, out_channels=1)``.

Same constructor signature, attribute names (encoder1..4, middle, decoder3..1, final) and therefore the same
``state_dict`` keys/shapes and — because the parameter containers are created through the same torch layer types
in the same order — bit-identical default initialisation under a given ``torch.manual_seed``
(reference models/model.py:6-51). ``forward`` (reference models/model.py:53-73) does not run the torch layers: it
hands the parameters to the B200 engine (libb2s kernels). CUDA tensors only; no CPU fallback.
"""
import torch
import torch.nn as nn

from ..engine import UNetEngine


def _stage_layers(cin, cout):
    return [nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.ReLU(inplace=True), nn.BatchNorm2d(cout)]


def _named_params(module):
    """name -> parameter tensor of `module`, also on nn.DataParallel replicas (utils/trainer.py:28-30): torch's
    replicate() leaves a replica's `_parameters` EMPTY and keeps the broadcast copies (non-leaf tensors whose gradient
    flows back to the master parameters through Broadcast.backward) as plain attributes listed in
    `_former_parameters`; `named_parameters()` of a replica therefore yields nothing."""
    out = {}
    for mname, m in module.named_modules():
        src = m._parameters
        if not src and getattr(m, "_is_replica", False):
            src = getattr(m, "_former_parameters", {})
        for k, v in src.items():
            if v is not None:
                out[f"{mname}.{k}" if mname else k] = v
    return out


class _UNetFunction(torch.autograd.Function):
    """Whole-network autograd node: forward and backward are single passes through the engine."""

    @staticmethod
    def forward(ctx, x, module, names, need_bwd, *params):
        P = dict(zip(names, params))
        P.update(dict(module.named_buffers()))
        with torch.cuda.device(x.device):
            logits, plan = module._engine.forward(P, x, train=module.training, need_backward=need_bwd,
                                                  cache_packed=not getattr(module, "_is_replica", False))
        ctx.P, ctx.plan, ctx.generation, ctx.names = P, plan, plan.generation, names
        ctx.engine = module._engine
        ctx.train_stats = module.training
        ctx.x_needs_grad = x.requires_grad
        return logits.clone()

    @staticmethod
    def backward(ctx, dlogits):
        plan = ctx.plan
        if plan.generation != ctx.generation:
            raise RuntimeError("b200seg UNet: the saved activations of this forward were overwritten by a later forward "
                               "of the same shape; call backward before the next forward")
        if not ctx.train_stats:
            raise RuntimeError("b200seg UNet: backward through eval-mode BatchNorm is not implemented "
                               "(the reference only back-propagates in train mode, utils/trainer.py:55,91)")
        P = ctx.P
        G = {n: torch.empty_like(P[n], dtype=torch.float32) for n in ctx.names}
        with torch.cuda.device(dlogits.device):
            ctx.engine.backward(P, plan, dlogits, G)
        return (None, None, None, None) + tuple(G[n] for n in ctx.names)


class UNet(nn.Module):
    def __init__(self, in_channels=1, out_channels=1):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.encoder1 = self.conv_block(in_channels, 64)
        self.encoder2 = self.conv_block(64, 128)
        self.encoder3 = self.conv_block(128, 256)
        self.encoder4 = self.conv_block(256, 512)
        self.middle = nn.Sequential(nn.MaxPool2d(kernel_size=2, stride=2), self.conv_block(512, 1024),
                                    nn.ConvTranspose2d(1024, 512, kernel_size=2, stride=2))
        self.decoder3 = self.upconv_block(1024, 256)
        self.decoder2 = self.upconv_block(512, 128)
        self.decoder1 = self.upconv_block(256, 64)
        self.final = nn.Sequential(self.conv_block(128, 64), nn.Conv2d(64, out_channels, kernel_size=1))
        self._engine = UNetEngine(out_channels=out_channels)

    # public helpers kept for API parity (reference models/model.py:33,45)
    def conv_block(self, in_channels, out_channels):
        return nn.Sequential(*(_stage_layers(in_channels, out_channels) + _stage_layers(out_channels, out_channels)))

    def upconv_block(self, in_channels, out_channels):
        return nn.Sequential(self.conv_block(in_channels, in_channels // 2),
                             nn.ConvTranspose2d(in_channels // 2, out_channels, kernel_size=2, stride=2))

    def _tensor_dict(self):
        d = _named_params(self)
        d.update(dict(self.named_buffers()))
        return d

    def invalidate_packed(self):
        """Forces the next forward to re-pack the bf16 GEMM operands. The pack is skipped while every weight's
        (tensor identity, address, version counter) is unchanged; writes that bypass the version counter
        (`p.data.copy_()`, an external kernel) must be followed by this call."""
        self._engine.invalidate_packed()

    def forward(self, x):
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise RuntimeError(f"UNet expected input [N,{self.in_channels},H,W], got {tuple(x.shape)}")
        if x.shape[2] % 16 or x.shape[3] % 16:
            raise RuntimeError("Sizes of tensors must match except in dimension 1: H and W must be multiples of 16")
        if not x.is_cuda:
            raise RuntimeError("b200seg UNet runs on CUDA (sm_100a) only; there is no CPU fallback")
        if x.requires_grad:
            raise RuntimeError("b200seg UNet does not compute the gradient with respect to the input image (the "
                               "reference's trainer never asks for it); detach the input or use the reference module")
        named = _named_params(self)
        names, params = tuple(named.keys()), tuple(named.values())
        if params[0].device != x.device:
            raise RuntimeError("UNet parameters and input are on different devices; call model.to(device)")
        if torch.is_autocast_enabled():
            x = x.float()
        # grad mode is off inside Function.forward, so decide here whether backward buffers are needed
        need_bwd = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _UNetFunction.apply(x, self, names, need_bwd, *params)

    @torch.no_grad()
    def predict_mask(self, x):
        """Inference entry (utils/trainer.py:216-217): returns (logits, uint8 mask) with mask = sigmoid(logits) > 0.5."""
        P = self._tensor_dict()
        with torch.cuda.device(x.device):
            logits, plan = self._engine.forward(P, x, train=False, want_mask=True, need_backward=False,
                                                cache_packed=not getattr(self, "_is_replica", False))
        return logits, plan.mask
