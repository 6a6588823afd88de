"""B200-native 3-D PatchGAN discriminator: drop-in for the reference's ``BaselineDiscriminator``
(/root/reference/src/networks/discriminator/baseline.py:21-88), the network ``AdversarialTrainer._iteration``
(/root/reference/src/engines/trainer.py:215-256) steps next to the VQ-VAE.

Same constructor arguments, same module tree -- ``main.<i>`` holds Conv3d / BatchNorm3d / LeakyReLU at the reference's
indices, so ``state_dict`` keys, buffer names (``running_mean``, ``running_var``, ``num_batches_tracked``) and the
initialisation random stream (``weights_init``: N(0, 0.02) conv weights, N(1, 0.02) BN weights) are identical -- but the
``nn`` modules are parameter containers only.  The stack runs channels-last over the C ABI in
``include/synthanatomy_b200.h``: the k4 strided / unit-stride convolutions with the VQ-VAE's conv kernels, BatchNorm3d +
LeakyReLU as one statistics pass and one normalise-activate pass, and a hand-scheduled backward that returns the
gradient of the input as well (the generator's adversarial loss differentiates through the discriminator).

Precision follows the caller like the VQ-VAE module: fp32 activations normally, bf16 activations with fp32 accumulation
and fp32 statistics under ``torch.autocast`` / ``compute_dtype=torch.bfloat16``.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from ... import ops
from ...ops import ConvSpec


def weights_init(m: nn.Module) -> None:
    """baseline.py:12-18"""
    classname = m.__class__.__name__
    if classname.find("Conv") != -1:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif classname.find("BatchNorm") != -1:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)


class _Container(nn.Sequential):
    def forward(self, x):  # pragma: no cover - containers are never executed
        raise RuntimeError("synthanatomy_b200: parameter container, not executable (no eager fallback)")


class _Block:
    """one Conv3d [-> BatchNorm3d] [-> LeakyReLU] group of the Sequential"""

    def __init__(self, conv: nn.Conv3d, bn: Optional[nn.BatchNorm3d], act: Optional[nn.LeakyReLU]):
        k = conv.kernel_size[0]
        assert conv.kernel_size == (k, k, k) and conv.dilation == (1, 1, 1)
        self.conv, self.bn, self.act = conv, bn, act
        self.spec = ConvSpec("conv", conv.in_channels, conv.out_channels, k, conv.stride[0], conv.padding[0])
        self.slope = float(act.negative_slope) if act is not None else 1.0

    def params(self) -> List[torch.Tensor]:
        out = [self.conv.weight]
        if self.conv.bias is not None:
            out.append(self.conv.bias)
        if self.bn is not None:
            out += [self.bn.weight, self.bn.bias]
        return out


class _DiscFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net: "B200Discriminator", x: torch.Tensor, *params: torch.Tensor):
        dt = net._dtype()
        training = net.training
        h = ops.ncdhw_to_ndhwc(x.contiguous(), dt)
        saved, pi = [], 0
        for blk in net._blocks:
            w = params[pi]; pi += 1
            b = None
            if blk.conv.bias is not None:
                b = params[pi]; pi += 1
            z = ops.conv_forward(blk.spec, h, ops.pack_weight(w, False, dt), b, None, False)
            if blk.bn is not None:
                gamma, beta = params[pi], params[pi + 1]; pi += 2
                bn = blk.bn
                if training or not bn.track_running_stats:
                    mom = bn.momentum
                    if bn.track_running_stats:
                        bn.num_batches_tracked += 1
                        if mom is None:
                            mom = 1.0 / float(bn.num_batches_tracked)
                    mean, rstd = ops.bn_stats(z, bn.eps, mom if mom is not None else 0.0,
                                              bn.running_mean if bn.track_running_stats else None,
                                              bn.running_var if bn.track_running_stats else None)
                else:
                    mean, rstd = ops.bn_eval_stats(bn.running_mean, bn.running_var, bn.eps)
                y = ops.bn_lrelu_fwd(z, mean, rstd, gamma.detach(), beta.detach(), blk.slope)
                saved.append((h, z, y, mean, rstd))
            elif blk.act is not None:
                y = ops.lrelu_fwd_(z, blk.slope)
                saved.append((h, None, y, None, None))
            else:
                y = z
                saved.append((h, None, None, None, None))
            h = y
        ctx.net, ctx.saved, ctx.params, ctx.was_training, ctx.dt = net, saved, params, training, dt
        ctx.in_dhw = tuple(x.shape[2:])
        return ops.ndhwc_to_ncdhw(h, torch.float32)

    @staticmethod
    def backward(ctx, gout: torch.Tensor):
        net, params = ctx.net, ctx.params
        dt = ctx.dt          # the forward's activation type (autocast may no longer be active when backward runs)
        if not ctx.was_training and any(b.bn is not None for b in net._blocks):
            raise RuntimeError("B200Discriminator: backward through eval-mode BatchNorm is not implemented")
        g = ops.ncdhw_to_ndhwc(gout.contiguous(), dt)
        grads: List[Optional[torch.Tensor]] = [None] * len(params)
        offs, pi = [], 0
        for blk in net._blocks:
            offs.append(pi)
            pi += len(blk.params())
        need_dx = ctx.needs_input_grad[1]
        for i in range(len(net._blocks) - 1, -1, -1):
            blk = net._blocks[i]
            h, z, y, mean, rstd = ctx.saved[i]
            o = offs[i]
            w = params[o]
            has_bias = blk.conv.bias is not None
            if blk.bn is not None:
                gamma = params[o + 1 + int(has_bias)]
                g, dgamma, dbeta = ops.bn_lrelu_bwd(g, z, y, mean, rstd, gamma.detach(), blk.slope, True)
                grads[o + 1 + int(has_bias)], grads[o + 2 + int(has_bias)] = dgamma, dbeta
            elif blk.act is not None:
                g = ops.lrelu_bwd_(g if i != len(net._blocks) - 1 else g.clone(), y, blk.slope)
            if ctx.needs_input_grad[2 + o]:
                grads[o] = ops.conv_wgrad(blk.spec, h, g, w)
            if has_bias and ctx.needs_input_grad[2 + o + 1]:
                grads[o + 1] = ops.bias_grad(g)
            if i > 0 or need_dx:
                g = ops.conv_dgrad(blk.spec, g, ops.pack_weight(w, True, dt), tuple(h.shape[1:4]))
        dx = ops.ndhwc_to_ncdhw(g, torch.float32) if need_dx else None
        return (None, dx, *grads)


class B200Discriminator(nn.Module):
    def __init__(self, input_nc: int = 1, ndf: int = 64, n_layers: int = 3, compute_dtype: Optional[torch.dtype] = None):
        super().__init__()
        self.compute_dtype = compute_dtype
        kw, padw = 4, 1
        # BatchNorm3d carries the affine shift, so the convs in front of it have no bias (baseline.py:33-38)
        sequence: List[nn.Module] = [nn.Conv3d(input_nc, ndf, kernel_size=kw, stride=2, padding=padw), nn.LeakyReLU(0.2, True)]
        nf_mult = 1
        for n in range(1, n_layers):
            nf_mult_prev, nf_mult = nf_mult, min(2 ** n, 8)
            sequence += [nn.Conv3d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=2, padding=padw, bias=False),
                         nn.BatchNorm3d(ndf * nf_mult), nn.LeakyReLU(0.2, True)]
        nf_mult_prev, nf_mult = nf_mult, min(2 ** n_layers, 8)
        sequence += [nn.Conv3d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=1, padding=padw, bias=False),
                     nn.BatchNorm3d(ndf * nf_mult), nn.LeakyReLU(0.2, True)]
        sequence += [nn.Conv3d(ndf * nf_mult, 1, kernel_size=kw, stride=1, padding=padw)]
        self.main = _Container(*sequence)
        self.apply(weights_init)
        self._blocks: List[_Block] = []
        mods = list(self.main)
        i = 0
        while i < len(mods):
            conv = mods[i]; i += 1
            bn = act = None
            if i < len(mods) and isinstance(mods[i], nn.BatchNorm3d):
                bn = mods[i]; i += 1
            if i < len(mods) and isinstance(mods[i], nn.LeakyReLU):
                act = mods[i]; i += 1
            self._blocks.append(_Block(conv, bn, act))

    def _dtype(self) -> torch.dtype:
        if self.compute_dtype is not None:
            return self.compute_dtype
        if torch.is_autocast_enabled():
            return torch.bfloat16
        return torch.float32

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if not input.is_cuda:
            raise RuntimeError("synthanatomy_b200: CUDA tensors only -- there is no CPU fallback")
        params: List[torch.Tensor] = []
        for blk in self._blocks:
            params += blk.params()
        return _DiscFunction.apply(self, input.float(), *params)
