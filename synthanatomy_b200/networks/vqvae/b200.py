"""B200-native VQ-VAE: drop-in for the reference's ``BaselineVQVAE``
(/root/reference/src/networks/vqvae/baseline.py:163-362).

Same constructor keywords, same ``state_dict`` keys / shapes (so README checkpoints load), same ``VQVAEBase``
API -- but no ``nn.Conv3d.forward`` is ever executed: the ``nn`` modules below are parameter containers
only.  The encoder / decoder stacks run as channels-last (NDHWC) programs over the C ABI in
``include/synthanatomy_b200.h`` with a hand-scheduled backward (ReLU masks and residual adds fused into the
dgrad epilogues), the quantiser is one fused kernel.

Precision: ``compute_dtype=None`` (default) follows the caller like the reference does -- fp32 activations
normally ("parity" path, CUDA-core fp32 FMA), bf16 activations + fp32 accumulation on tcgen05 tensor cores
when called under ``torch.autocast`` (the reference trains with ``--amp=True``, README.md:52).
``compute_dtype=ops.BF16X3`` is the tensor-core parity mode: fp32 activations, every product on the bf16 tensor
cores as hi.hi + lo.hi + hi.lo of split operands (csrc/sa_x3.cu), 1e-4 against the reference in fp32.  The quantiser
is always fp32 (baseline.py:38,43).
"""
from __future__ import annotations

import weakref
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.distributed as dist
import torch.nn as nn

from ... import ops
from ...ops import ConvSpec
from .vqvae import VQVAEBase


# ------------------------------------------------------------------------------------------------
# parameter containers (identical module tree => identical state_dict keys and identical default init stream)
# ------------------------------------------------------------------------------------------------
class ResidualLayer(nn.Sequential):
    """Container mirroring baseline.py:150-156 (indices 0 and 3 hold the 3x3x3 and 1x1x1 convs)."""

    def __init__(self, n_channels: int, n_res_channels: int, p_dropout: float):
        super().__init__(
            nn.Conv3d(n_channels, n_res_channels, kernel_size=3, padding=1),
            nn.ReLU(True),
            nn.Dropout3d(p_dropout),
            nn.Conv3d(n_res_channels, n_channels, kernel_size=1),
        )

    def forward(self, x):  # pragma: no cover - containers are never executed
        raise RuntimeError("synthanatomy_b200: parameter container, not executable (no eager fallback)")


class _Container(nn.Sequential):
    def forward(self, x):  # pragma: no cover
        raise RuntimeError("synthanatomy_b200: parameter container, not executable (no eager fallback)")


# ------------------------------------------------------------------------------------------------
# stack programme
# ------------------------------------------------------------------------------------------------
class _ConvOp:
    def __init__(self, module: nn.Module, kind: str, relu: bool):
        self.module = module
        self.relu = relu
        k = module.kernel_size[0]
        assert module.kernel_size == (k, k, k) and module.dilation == (1, 1, 1), "cubic kernels, dilation 1 only"
        s, p = module.stride[0], module.padding[0]
        if kind == "deconv":
            assert module.output_padding == (0, 0, 0), "output_padding 0 only"
        self.spec = ConvSpec(kind, module.in_channels, module.out_channels, k, s, p)

    def params(self):
        return [self.module.weight, self.module.bias]

    def single_channel_gemm(self, dt: torch.dtype) -> bool:
        """tensor-core paths only (bf16, or fp32 tensors in the bf16x3 mode): the 1-channel ends of the network (first
        Conv3d 1->C, last ConvTranspose3d C->1) run as im2col / col2im + a 1x1x1 tensor-core GEMM over the k^3 = 64 taps."""
        sp = self.spec
        tensor_core = dt == torch.bfloat16 or (dt == torch.float32 and ops.x3_enabled())
        if not tensor_core or sp.k ** 3 != 64:
            return False
        if sp.kind == "conv":
            return sp.cin == 1 and sp.cout % 16 == 0
        return sp.cout == 1 and sp.cin % 64 == 0 and not self.relu


class _ResOp:
    def __init__(self, layer: ResidualLayer):
        self.c3 = _ConvOp(layer[0], "conv", True)
        self.c1 = _ConvOp(layer[3], "conv", False)
        assert layer[2].p == 0.0, "dropout p > 0 is not implemented (README config uses 0; no fallback)"

    def params(self):
        return self.c3.params() + self.c1.params()


def _stack_forward(ops_list, x: torch.Tensor, params: Sequence[torch.Tensor], save: bool):
    """x: NDHWC activation.  Returns (y, saved) where saved[i] holds what layer i's backward needs."""
    dt = x.dtype
    saved = []
    pi = 0
    for op in ops_list:
        if isinstance(op, _ConvOp):
            w, b = params[pi], params[pi + 1]
            pi += 2
            sp = op.spec
            if op.single_channel_gemm(dt):
                taps = sp.k ** 3
                if sp.kind == "conv":      # cols[pos][t] . w[n][t]
                    cols = ops.im2col_c1(x, sp.k, sp.s, sp.p, sp.out_dhw(x.shape[1:4]))
                    wp = ops.pack_weight(w.view(sp.cout, taps, 1), False, dt)
                    y = ops.conv_forward(ConvSpec("conv", taps, sp.cout, 1, 1, 0), cols, wp, b, None, op.relu)
                    saved.append((x, cols) if save else None)
                else:                      # r[pos][t] = x[pos] . wt[:, t];  y = col2im(r) + b
                    wp = ops.pack_weight(w.view(sp.cin, taps, 1), True, dt)
                    r = ops.conv_forward(ConvSpec("conv", sp.cin, taps, 1, 1, 0), x, wp, None, None, False)
                    y = ops.col2im_c1(r, sp.k, sp.s, sp.p, sp.out_dhw(x.shape[1:4]), b)
                    saved.append((x,) if save else None)
            else:
                wp = ops.pack_weight(w, transpose=(sp.kind == "deconv"), dtype=dt)
                y = ops.conv_forward(sp, x, wp, b, None, op.relu)
                saved.append((x,) if save else None)
            x = y
        else:
            w3, b3, w1, b1 = params[pi:pi + 4]
            pi += 4
            wp3 = ops.pack_weight(w3, False, dt)
            wp1 = ops.pack_weight(w1, False, dt)
            h = ops.conv_forward(op.c3.spec, x, wp3, b3, None, True)
            if ops.conv1x1_bwd_fused_supported(op.c1.spec, h):
                y = ops.conv1x1_fwd_fused(op.c1.spec, h, wp1, b1, x, True)   # relu(x + conv1x1(h)), streaming kernel
            else:
                y = ops.conv_forward(op.c1.spec, h, wp1, b1, x, True)   # relu(x + conv1x1(h))
            saved.append((x, h) if save else None)
            x = y
    return x, saved


def _stack_backward(ops_list, saved, params, g, in_is_relu: bool, need_dx: bool, allow_relu_tail: bool = False):
    """g: gradient w.r.t. the stack output (NDHWC).  Returns (dx | None, grads).  If the last op ends in a ReLU, g must
    already carry that ReLU's mask (allow_relu_tail: the consumer's data-gradient epilogue applied it)."""
    grads: List[Optional[torch.Tensor]] = [None] * len(params)
    n = len(ops_list)
    # offsets of each op's params
    offs, pi = [], 0
    for op in ops_list:
        offs.append(pi)
        pi += 2 if isinstance(op, _ConvOp) else 4
    # whether the INPUT of op i is post-ReLU (then the dgrad epilogue applies the ReLU mask of that tensor)
    relu_in = [in_is_relu]
    for op in ops_list[:-1]:
        relu_in.append(True if isinstance(op, _ResOp) else op.relu)

    last = ops_list[-1]
    if not allow_relu_tail and (isinstance(last, _ResOp) or last.relu):
        # the stacks of this model never end in a ReLU (encoder: 3x3x3 projection, decoder: last transposed conv)
        raise RuntimeError("synthanatomy_b200: stack ending in ReLU is not supported")

    for i in range(n - 1, -1, -1):
        op = ops_list[i]
        o = offs[i]
        first = i == 0
        want_dx = need_dx or not first
        if isinstance(op, _ConvOp) and op.single_channel_gemm(g.dtype):
            sp = op.spec
            taps = sp.k ** 3
            w = params[o]
            grads[o + 1] = ops.bias_grad(g)
            if sp.kind == "conv":
                x, cols = saved[i]
                if want_dx:
                    raise RuntimeError("synthanatomy_b200: data gradient of a 1-channel input conv is not implemented")
                gemm = ConvSpec("conv", taps, sp.cout, 1, 1, 0)
                grads[o] = ops.conv_wgrad(gemm, cols, g, w.view(sp.cout, taps, 1)).view_as(w)
                g = None
            else:
                (x,) = saved[i]
                cols = ops.im2col_c1(g, sp.k, sp.s, sp.p, x.shape[1:4])          # cols[i][t] = g[i*s - p + t]
                gemm = ConvSpec("conv", taps, sp.cin, 1, 1, 0)                   # "input" cols, "output" x-shaped
                grads[o] = ops.conv_wgrad(gemm, cols, x, w.view(sp.cin, taps, 1)).view_as(w)
                if want_dx:
                    wp = ops.pack_weight(w.view(sp.cin, taps, 1), False, g.dtype)
                    g = ops.conv_forward(gemm, cols, wp, None, None, False, x if relu_in[i] else None)
                else:
                    g = None
        elif isinstance(op, _ConvOp):
            (x,) = saved[i]
            w = params[o]
            grads[o] = ops.conv_wgrad(op.spec, x, g, w)
            grads[o + 1] = ops.bias_grad(g)
            if want_dx:
                wp_t = ops.pack_weight(w, transpose=(op.spec.kind == "conv"), dtype=g.dtype)
                g = ops.conv_dgrad(op.spec, g, wp_t, x.shape[1:4], None, x if relu_in[i] else None)
            else:
                g = None
        else:
            x, h = saved[i]
            w3, w1 = params[o], params[o + 2]
            # y = relu(x + conv1(h) + b1), h = relu(conv3(x) + b3); g already carries the (y > 0) mask
            wp1_t = ops.pack_weight(w1, True, g.dtype)
            if ops.conv1x1_bwd_fused_supported(op.c1.spec, g):
                # one pass over g and h: dh = dgrad(g) * (h > 0), dW1, db1 and db3 = the column sums of dh
                dh, grads[o + 2], grads[o + 3], grads[o + 1] = ops.conv1x1_bwd_fused(op.c1.spec, g, h, wp1_t, w1,
                                                                                      with_dbh=True)
            else:
                grads[o + 2] = ops.conv_wgrad(op.c1.spec, h, g, w1)
                grads[o + 3] = ops.bias_grad(g)
                dh = ops.conv_dgrad(op.c1.spec, g, wp1_t, h.shape[1:4], None, h)        # * (h > 0)
                grads[o + 1] = ops.bias_grad(dh)
            grads[o] = ops.conv_wgrad(op.c3.spec, x, dh, w3)
            if want_dx:
                wp3_t = ops.pack_weight(w3, True, g.dtype)
                g = ops.conv_dgrad(op.c3.spec, dh, wp3_t, x.shape[1:4], g, x if relu_in[i] else None)  # (+ g) * (x > 0)
            else:
                g = None
    return g, grads


# The backward of a stack drops its saved activations as soon as it has run (the decoder's 40 GB are gone before the
# encoder's backward starts).  A caller that differentiates the same graph more than once -- the adaptive adversarial
# weight, /root/reference/src/engines/trainer.py:264-289: two autograd.grad(..., retain_graph=True) before backward() --
# wraps the iteration in `retain_activations()`.
_RETAIN_ACTIVATIONS = [False]
_PACK_PLANS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()    # module -> {id(programme): ops.PackPlan}


class retain_activations:
    def __enter__(self):
        self.prev = _RETAIN_ACTIVATIONS[0]
        _RETAIN_ACTIVATIONS[0] = True
        return self

    def __exit__(self, *exc):
        _RETAIN_ACTIVATIONS[0] = self.prev
        return False


class _InFn(torch.autograd.Function):
    """NCDHW fp32 (what the reference's callers hand over) -> channels-last NDHWC in the activation dtype"""

    @staticmethod
    def forward(ctx, x, dt):
        if not x.is_cuda:
            raise RuntimeError("synthanatomy_b200: CUDA tensors only -- there is no CPU fallback")
        return ops.ncdhw_to_ndhwc(x.detach().float().contiguous(), dt)

    @staticmethod
    def backward(ctx, g):
        return ops.ndhwc_to_ncdhw(g.contiguous(), torch.float32), None


class _OutFn(torch.autograd.Function):
    """NDHWC activation -> NCDHW fp32"""

    @staticmethod
    def forward(ctx, y, dt):
        ctx.dt = dt
        return ops.ndhwc_to_ncdhw(y.detach(), torch.float32)

    @staticmethod
    def backward(ctx, gout):
        return ops.ncdhw_to_ndhwc(gout.float().contiguous(), ctx.dt), None


class _OpFn(torch.autograd.Function):
    """One op of a stack programme (a conv / transposed conv, or a whole ResidualLayer) on NDHWC activations.  One
    autograd node per op: a layer's parameter gradients exist as soon as ITS backward has run, so DistributedDataParallel
    reduces the buckets of the upper layers while the lower layers are still in their backward pass, and a layer's saved
    activations are released right after use."""

    @staticmethod
    def forward(ctx, x, op, relu_in, x3, *params):
        need_grad = any(ctx.needs_input_grad)   # (grad mode is off inside Function.forward)
        pdet = [p.detach().float().contiguous() for p in params]
        with ops.x3_mode(x3):
            y, saved = _stack_forward([op], x.detach(), pdet, need_grad)
        if need_grad:
            ctx.op, ctx.saved, ctx.pdet, ctx.relu_in, ctx.x3 = op, saved, pdet, relu_in, x3
            ctx.plan = ops.current_pack_plan()       # the transposed packs of the backward pass come from the same plan
        return y

    @staticmethod
    def backward(ctx, g):
        if ctx.saved is None:
            raise RuntimeError("B200VQVAE: second backward through a layer whose activations were released; wrap the "
                               "iteration in synthanatomy_b200.networks.vqvae.b200.retain_activations()")
        with ops.x3_mode(ctx.x3), ops.pack_plan(ctx.plan, begin=False):
            dx, grads = _stack_backward([ctx.op], ctx.saved, ctx.pdet, g.contiguous(), ctx.relu_in,
                                        ctx.needs_input_grad[0], allow_relu_tail=True)
        if not _RETAIN_ACTIVATIONS[0]:
            ctx.saved = None
        return (dx, None, None, None, *grads)


class _QuantizeFn(torch.autograd.Function):
    """Quantizer_impl.forward, baseline.py:38-87, as one fused kernel (+ EMA kernel in training)."""

    @staticmethod
    def forward(ctx, x, impl, decay, commitment_cost, training):
        if not x.is_cuda:
            raise RuntimeError("synthanatomy_b200: CUDA tensors only -- there is no CPU fallback")
        b, c = x.shape[0], x.shape[1]
        sp = tuple(x.shape[2:])
        xf = x.detach().float().contiguous()
        flat = ops.ncdhw_to_ndhwc(xf, torch.float32).view(-1, c)                      # baseline.py:46
        K = impl.n_embed
        stats = torch.zeros(K + K * c + 1, device=x.device, dtype=torch.float32)
        cb = impl.weight.detach()
        idx, qst = ops.vq_forward(flat, cb, stats, straight_through=True)             # :49-63, :85
        impl.last_perplexity = ops.vq_perplexity(stats[:K], flat.shape[0])            # :110-120 (local histogram)
        if training:                                                                   # :66
            if dist.is_available() and dist.is_initialized():                          # :70-72, one packed call
                dist.all_reduce(stats[: K + K * c], op=dist.ReduceOp.SUM)
            ops.vq_ema_update(impl.N, impl.embed_avg, impl.weight.data, stats, decay, impl.eps)   # :75-80
        numel = flat.numel()
        latent_loss = stats[-1] * (float(commitment_cost) / numel)                     # :82
        qst_ncdhw = ops.ndhwc_to_ncdhw(qst.view(b, *sp, c), torch.float32)
        embed_idx = idx.view(b, *sp)                                                   # :60
        ctx.save_for_backward(xf, qst_ncdhw)
        ctx.coef = 2.0 * float(commitment_cost) / numel
        ctx.mark_non_differentiable(embed_idx)
        return qst_ncdhw, latent_loss, embed_idx

    @staticmethod
    def backward(ctx, g_q, g_loss, _g_idx):
        xf, q = ctx.saved_tensors
        g_q = g_q.float().contiguous() if g_q is not None else None
        g_loss = g_loss.float().contiguous() if g_loss is not None else None
        dz = ops.vq_backward(g_q, g_loss, xf, q, ctx.coef)
        return dz, None, None, None, None


class Quantizer_impl(nn.Module):
    """Container with the reference's parameter / buffer names (baseline.py:24-36)."""

    def __init__(self, n_embed, embed_dim, eps):
        super().__init__()
        self.embed_dim = embed_dim
        self.n_embed = n_embed
        self.eps = eps
        self.embedding = nn.Embedding(n_embed, embed_dim)
        self.embedding.weight.requires_grad = False
        self.weight = self.embedding.weight
        self.register_buffer("N", torch.zeros(n_embed))
        self.register_buffer("embed_avg", self.weight.data.clone())

    def forward(self, x: torch.Tensor, decay: float, commitment_cost: float) -> List[torch.Tensor]:
        return list(_QuantizeFn.apply(x, self, decay, commitment_cost, self.training))

    def embed(self, embedding_indices: torch.Tensor) -> torch.Tensor:
        """baseline.py:89-91: indices [B, d, h, w] -> [B, C, d, h, w]"""
        if not embedding_indices.is_cuda:
            raise RuntimeError("synthanatomy_b200: CUDA tensors only -- there is no CPU fallback")
        idx = embedding_indices.long().contiguous()
        q = ops.vq_embed(idx.view(-1), self.weight.detach())
        q = q.view(*idx.shape, self.embed_dim)
        return ops.ndhwc_to_ncdhw(q, torch.float32)


class Quantizer(nn.Module):
    """baseline.py:94-147."""

    def __init__(self, n_embed, embed_dim, commitment_cost=0.25, decay=0.99, eps=1e-5):
        super().__init__()
        self.impl = Quantizer_impl(n_embed, embed_dim, eps)
        self.n_embed = n_embed
        self.commitment_cost = commitment_cost
        self.decay = decay
        self.perplexity_code: torch.Tensor = torch.rand(1)

    def forward(self, x):
        quantized_st, latent_loss, embed_idx = self.impl(x, self.decay, self.commitment_cost)
        self.perplexity_code = self.impl.last_perplexity   # baseline.py:110-120, from the kernel's histogram
        return quantized_st, latent_loss

    def get_ema_decay(self) -> float:
        return self.decay

    def set_ema_decay(self, decay: float) -> float:
        self.decay = decay
        return self.get_ema_decay()

    def get_commitment_cost(self) -> float:
        return self.commitment_cost

    def set_commitment_cost(self, commitment_cost) -> float:
        self.commitment_cost = commitment_cost
        return self.get_commitment_cost()

    def get_perplexity(self) -> torch.Tensor:
        return self.perplexity_code

    def embed(self, embedding_indices: torch.Tensor) -> torch.Tensor:
        return self.impl.embed(embedding_indices=embedding_indices)

    def quantize(self, encodings: torch.Tensor) -> torch.Tensor:
        return self.impl(encodings, self.decay, self.commitment_cost)


class B200VQVAE(VQVAEBase, nn.Module):
    def __init__(
        self,
        n_levels: int = 3,
        downsample_parameters: Tuple[Tuple[int, int, int, int], ...] = ((4, 2, 1, 1),) * 3,
        upsample_parameters: Tuple[Tuple[int, int, int, int, int], ...] = ((4, 2, 1, 0, 1),) * 3,
        n_embed: int = 256,
        embed_dim: int = 256,
        n_channels: int = 144,
        n_res_channels: int = 144,
        n_res_layers: int = 3,
        p_dropout: float = 0.0,
        commitment_cost: float = 0.25,
        vq_decay: float = 0.5,
        use_subpixel_conv: bool = False,
        compute_dtype: Union[torch.dtype, str, None] = None,
    ):
        super().__init__()
        assert n_levels == len(downsample_parameters) and n_levels == len(upsample_parameters), (
            f"downsample_parameters, upsample_parameters must have the same number of elements as n_levels. "
            f"But got {len(downsample_parameters)} and {len(upsample_parameters)}, instead of {n_levels}."
        )
        if use_subpixel_conv:
            raise NotImplementedError("use_subpixel_conv=True is not implemented (README.md:80 uses False); no fallback")
        if p_dropout != 0.0:
            raise NotImplementedError("dropout > 0 is not implemented (README.md:93 uses 0.0); no fallback")
        for prm in tuple(downsample_parameters) + tuple(upsample_parameters):
            if prm[-1] != 1:
                raise NotImplementedError("dilation != 1 is not implemented; no fallback")
        if embed_dim not in (8, 16, 32, 64):
            # the fused quantiser kernel keeps one latent row in registers (csrc/sa_vq.cu); README.md:83 uses 32
            raise NotImplementedError(f"embed_dim={embed_dim} is not implemented (the quantiser kernel takes 8, 16, 32 or "
                                      f"64; README.md:83 uses 32); no fallback")
        self.n_levels = n_levels
        self.downsample_parameters = downsample_parameters
        self.upsample_parameters = upsample_parameters
        self.n_embed = n_embed
        self.embed_dim = embed_dim
        self.use_subpixel_conv = use_subpixel_conv
        self.n_channels = n_channels
        self.n_res_channels = n_res_channels
        self.n_res_layers = n_res_layers
        self.p_dropout = p_dropout
        self.commitment_cost = commitment_cost
        self.vq_decay = vq_decay
        self.compute_dtype = compute_dtype

        self.encoder = self.construct_encoder()
        self.quantizer = self.construct_quantizer()
        self.decoder = self.construct_decoder()
        self._enc_ops = self._program(self.encoder[0])
        self._dec_ops = self._program(self.decoder[0])

    # ---- containers: same order / indices as baseline.py:213-299 ----
    def construct_encoder(self) -> nn.ModuleList:
        modules: List[nn.Module] = []
        for i in range(self.n_levels):
            last = i == self.n_levels - 1
            modules.append(nn.Conv3d(
                in_channels=1 if i == 0 else self.n_channels // 2,
                out_channels=self.n_channels // (1 if last else 2),
                kernel_size=self.downsample_parameters[i][0], stride=self.downsample_parameters[i][1],
                padding=self.downsample_parameters[i][2], dilation=self.downsample_parameters[i][3]))
            modules.append(nn.ReLU())
            modules.append(_Container(*[
                ResidualLayer(self.n_channels // (1 if last else 2), self.n_res_channels // (1 if last else 2),
                              self.p_dropout) for _ in range(self.n_res_layers)]))
        modules.append(nn.Conv3d(self.n_channels, self.embed_dim, 3, stride=1, padding=1))
        return nn.ModuleList([_Container(*modules)])

    def construct_quantizer(self) -> nn.ModuleList:
        return nn.ModuleList([Quantizer(self.n_embed, self.embed_dim, commitment_cost=self.commitment_cost,
                                        decay=self.vq_decay)])

    def construct_decoder(self) -> nn.ModuleList:
        modules: List[nn.Module] = [nn.Conv3d(self.embed_dim, self.n_channels, 3, stride=1, padding=1)]
        for i in range(self.n_levels):
            first, last = i == 0, i == self.n_levels - 1
            modules.append(_Container(*[
                ResidualLayer(self.n_channels // (1 if first else 2), self.n_res_channels // (1 if first else 2),
                              self.p_dropout) for _ in range(self.n_res_layers)]))
            modules.append(nn.ConvTranspose3d(
                in_channels=self.n_channels // (1 if first else 2),
                out_channels=1 if last else self.n_channels // 2,
                kernel_size=self.upsample_parameters[i][0], stride=self.upsample_parameters[i][1],
                padding=self.upsample_parameters[i][2], output_padding=self.upsample_parameters[i][3],
                dilation=self.upsample_parameters[i][4]))
            if not last:
                modules.append(nn.ReLU())
        return nn.ModuleList([_Container(*modules)])

    @staticmethod
    def _program(seq: nn.Sequential):
        mods = list(seq)
        prog = []
        for j, m in enumerate(mods):
            nxt_relu = j + 1 < len(mods) and isinstance(mods[j + 1], nn.ReLU)
            if isinstance(m, nn.ConvTranspose3d):
                prog.append(_ConvOp(m, "deconv", nxt_relu))
            elif isinstance(m, nn.Conv3d):
                prog.append(_ConvOp(m, "conv", nxt_relu))
            elif isinstance(m, _Container):
                prog.extend(_ResOp(r) for r in m)
        return prog

    def _dtype(self) -> torch.dtype:
        if self.compute_dtype is not None:
            return self.compute_dtype
        return torch.bfloat16 if torch.is_autocast_enabled() else torch.float32

    def _run(self, prog, x):
        dt, x3 = ops.resolve_dtype(self._dtype())     # BF16X3: fp32 tensors, split-bf16 tensor-core products
        h = _InFn.apply(x, dt)
        relu_in = False                               # is the op's input a post-ReLU tensor?  (mask fused into its dgrad)
        # every conv weight of the stack (forward and transposed forms) is packed by a few multi-tensor launches here
        plans = _PACK_PLANS.setdefault(self, {})
        plan = plans.setdefault(id(prog), ops.PackPlan())
        with ops.pack_plan(plan):
            for op in prog:
                h = _OpFn.apply(h, op, relu_in, x3, *op.params())
                relu_in = True if isinstance(op, _ResOp) else op.relu
        return _OutFn.apply(h, dt)

    # ---- VQVAEBase API (baseline.py:301-362) ----
    def get_ema_decay(self) -> Sequence[float]:
        return [self.quantizer[0].get_ema_decay()]

    def set_ema_decay(self, decay: Union[Sequence[float], float]) -> Sequence[float]:
        self.quantizer[0].set_ema_decay(decay[0] if isinstance(decay, list) else decay)
        return self.get_ema_decay()

    def get_commitment_cost(self) -> Sequence[float]:
        return [self.quantizer[0].get_commitment_cost()]

    def set_commitment_cost(self, commitment_factor: Union[Sequence[float], float]) -> Sequence[float]:
        self.quantizer[0].set_commitment_cost(
            commitment_factor[0] if isinstance(commitment_factor, list) else commitment_factor)
        return self.get_commitment_cost()

    def get_perplexity(self) -> Sequence[float]:
        return [self.quantizer[0].get_perplexity()]

    def get_last_layer(self) -> nn.parameter.Parameter:
        return list(self.decoder.modules())[-1].weight

    def encode(self, images: torch.Tensor) -> List[torch.Tensor]:
        return [self._run(self._enc_ops, images)]

    def quantize(self, encodings: List[torch.Tensor]) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
        x, x_loss = self.quantizer[0](encodings[0])
        return [x], [x_loss]

    def decode(self, quantizations: List[torch.Tensor]) -> torch.Tensor:
        return self._run(self._dec_ops, quantizations[0])

    def index_quantize(self, images: torch.Tensor) -> List[torch.Tensor]:
        encodings = self.encode(images)
        _, _, encoding_indices = self.quantizer[0].quantize(encodings[0])
        return [encoding_indices]

    def decode_samples(self, embedding_indices: List[torch.Tensor]) -> torch.Tensor:
        samples_codes = self.quantizer[0].embed(embedding_indices[0])
        return self.decode([samples_codes])

    def forward(self, images: torch.Tensor) -> Dict[str, List[torch.Tensor]]:
        ops.sync_deterministic()          # torch.backends.cudnn.deterministic (the reference's `deterministic`) -> ordered sums
        encodings = self.encode(images)
        quantizations, quantization_losses = self.quantize(encodings)
        reconstruction = self.decode(quantizations)
        return {"reconstruction": [reconstruction], "quantization_losses": quantization_losses}
