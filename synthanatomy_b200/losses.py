"""Losses on the hot path: mirrors of the reference's ``MSELoss`` for the VQ-VAE
(/root/reference/src/losses/vqvae/vqvae.py:14-71): ``mse(reconstruction, y) + sum(quantization_losses)``, of its
``JukeboxLoss`` (spectral amplitude MSE + pixel MSE, vqvae.py:522-640; the spectral half of the README run's
``jukebox_perceptual`` loss) and of its ``CELoss`` for the Performer
(/root/reference/src/losses/transformer/transformer.py:10-36)."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional

import torch
from torch.nn.modules.loss import _Loss

from . import _lib, ops, pf_ops

# The TensorBoard summaries hold DETACHED scalars: the handlers only read their values, and a live tensor would keep the
# whole autograd graph of the last step (and the parameters' gradient accumulators with their stream) alive between steps.


class _MSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        p = pred.detach().float().contiguous()
        t = target.detach().float().contiguous()
        sse = torch.zeros((), device=p.device, dtype=torch.float32)
        ops.mse_fwd_bwd(p, t, sse, None, 1.0)
        ctx.save_for_backward(p, t)
        return sse / p.numel()

    @staticmethod
    def backward(ctx, g):
        p, t = ctx.saved_tensors
        grad = torch.empty_like(p)
        ops.mse_fwd_bwd(p, t, None, grad, 2.0 / p.numel(), g.float().contiguous())
        return grad, None


def mse_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """F.mse_loss(pred, target) (mean reduction) through the library's fused kernel."""
    return _MSEFn.apply(pred, target)


class MSELoss(_Loss):
    def __init__(self, size_average: bool = None, reduce: bool = None, reduction: str = "mean"):
        super().__init__(size_average, reduce, reduction)
        if reduction not in ("sum", "mean"):
            raise ValueError("Reduction must be either 'sum' or 'mean'")
        self.summaries: Dict = {"scalar": dict()}

    def forward(self, network_output: Dict[str, List[torch.Tensor]], y: torch.Tensor) -> torch.Tensor:
        y_pred = network_output["reconstruction"][0]
        q_losses = network_output["quantization_losses"]
        loss = mse_loss(y_pred, y)     # the reference ignores `reduction` here as well (F.mse_loss default)
        self.summaries["scalar"]["Loss-MSE-Reconstruction"] = loss.detach()
        for idx, q_loss in enumerate(q_losses):
            q_loss = q_loss.float()
            self.summaries["scalar"][f"Loss-MSE-VQ{idx}_Commitment_Cost"] = q_loss.detach()
            loss = loss + q_loss
        return loss

    def get_summaries(self) -> Dict[str, torch.Tensor]:
        return self.summaries


# ------------------------------------------------------------------------------------------------
# spectral (Jukebox) loss: 3-D orthonormal DFT as three dense DFT-matrix products on the tensor cores (bf16x3)
# ------------------------------------------------------------------------------------------------
_DFT_CACHE: Dict = {}


def _dft_matrices(n: int, complex_in: bool, device):
    """(M, M^T) of one orthonormal DFT axis as REAL matrices over [part][index] vectors: output rows (re', k') then
    (im', k').  Real input: M = [C ; -S] (2n x n).  Complex input: M = [[C | S] ; [-S | C]] (2n x 2n).  F = C - iS,
    C = cos(2 pi k j / n) / sqrt(n).  Built in float64 on the host once per (n, kind, device)."""
    key = (n, complex_in, str(device))
    if key not in _DFT_CACHE:
        k = torch.arange(n, dtype=torch.float64)
        ang = 2.0 * math.pi * torch.outer(k, k) / n
        c, s_ = torch.cos(ang) / math.sqrt(n), torch.sin(ang) / math.sqrt(n)
        m = torch.cat((c, -s_), dim=0) if not complex_in else torch.cat((torch.cat((c, s_), dim=1), torch.cat((-s_, c), dim=1)), dim=0)
        m = m.float().to(device).contiguous()
        _DFT_CACHE[key] = (m, m.t().contiguous())
    return _DFT_CACHE[key]


def _swap(src: torch.Tensor, batch: int, A: int, M: int, Cc: int) -> torch.Tensor:
    """dst[b][c][m][a] = src[b][a][m][c]"""
    dst = torch.empty_like(src)
    _lib.check(ops.lib().sa_swap_outer_inner(ops._p(src), ops._p(dst), batch, A, M, Cc, ops._stream()), "sa_swap_outer_inner")
    return dst


def _gemm(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    out = torch.empty((a.shape[0], b.shape[0]), device=a.device, dtype=torch.float32)
    pf_ops.gemm_nt(a, b, out_f32=out)
    return out


def _spectrum(x: torch.Tensor) -> torch.Tensor:
    """x [B, D, H, W] fp32 -> its orthonormal 3-D DFT as [B * H * W, 2, D] (rows (b, h', w'), parts re | im, index d')"""
    B, D, H, W = x.shape
    dev = x.device
    t = _gemm(x.view(B * D * H, W), _dft_matrices(W, False, dev)[0])              # [b, d, h][part][w']
    t = _swap(t, B * D, H, 2, W)                                                    # [b, d][w'][part][h]
    t = _gemm(t.view(B * D * W, 2 * H), _dft_matrices(H, True, dev)[0])            # [b, d, w'][part][h']
    t = _swap(t, B, D, W * 2, H)                                                    # [b][h'][w'][part][d]
    t = _gemm(t.view(B * H * W, 2 * D), _dft_matrices(D, True, dev)[0])            # [b, h', w'][part][d']
    return t.view(B * H * W, 2, D)


def _spectrum_transposed(g: torch.Tensor, B: int, D: int, H: int, W: int) -> torch.Tensor:
    """the adjoint of `_spectrum`: gradient w.r.t. the spectrum [B * H * W, 2, D] -> gradient w.r.t. x [B, D, H, W]"""
    dev = g.device
    t = _gemm(g.view(B * H * W, 2 * D), _dft_matrices(D, True, dev)[1])            # [b, h', w'][part][d]
    t = _swap(t, B, H, W * 2, D)                                                    # [b][d][w'][part][h']
    t = _gemm(t.view(B * D * W, 2 * H), _dft_matrices(H, True, dev)[1])            # [b, d, w'][part][h]
    t = _swap(t, B * D, W, 2, H)                                                    # [b, d][h][part][w']
    t = _gemm(t.view(B * D * H, 2 * W), _dft_matrices(W, False, dev)[1])           # [b, d, h][w]
    return t.view(B, D, H, W)


class _SpectralFn(torch.autograd.Function):
    """F.mse_loss(|fftn(pred)|, |fftn(target)|) with norm="ortho" over the channel + spatial axes (vqvae.py:598-599,
    617-630), one channel.  Always evaluated in the bf16x3 arithmetic (fp32-class), like the reference's .float()."""

    @staticmethod
    def forward(ctx, pred, target):
        if not pred.is_cuda:
            raise RuntimeError("synthanatomy_b200: CUDA tensors only -- there is no CPU fallback")
        if pred.dim() != 5 or pred.shape[1] != 1:
            raise NotImplementedError("spectral loss: [B, 1, D, H, W] volumes only (README configuration); no fallback")
        B, _, D, H, W = pred.shape
        p = pred.detach().float().contiguous().view(B, D, H, W)
        t = target.detach().float().contiguous().view(B, D, H, W)
        with ops.x3_mode(True):
            ps, ts = _spectrum(p), _spectrum(t)
        sse = torch.zeros((), device=p.device, dtype=torch.float32)
        _lib.check(ops.lib().sa_spectral_amp_loss(ops._p(ps), ops._p(ts), B * H * W, D, 0.0, None, ops._p(sse), None,
                                                  ops._stream()), "sa_spectral_amp_loss")
        ctx.save_for_backward(ps, ts)
        ctx.shape = (B, D, H, W)
        return sse / p.numel()

    @staticmethod
    def backward(ctx, g):
        ps, ts = ctx.saved_tensors
        B, D, H, W = ctx.shape
        gs = torch.empty_like(ps)
        _lib.check(ops.lib().sa_spectral_amp_loss(ops._p(ps), ops._p(ts), B * H * W, D, 2.0 / (B * D * H * W),
                                                  ops._p(g.float().contiguous().view(1)), None, ops._p(gs), ops._stream()),
                   "sa_spectral_amp_loss")
        with ops.x3_mode(True):
            gx = _spectrum_transposed(gs, B, D, H, W)
        return gx.view(B, 1, D, H, W), None


def spectral_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return _SpectralFn.apply(pred, target)


class JukeboxLoss(_Loss):
    """Drop-in for the reference's ``JukeboxLoss`` (vqvae.py:522-640), default ``fft_kwargs`` only."""

    def __init__(self, dimensions: int, include_pixel_loss: bool = True, fft_kwargs: Optional[Dict] = None,
                 size_average: bool = True, reduce: bool = True, reduction: str = "mean"):
        super().__init__(size_average, reduce, reduction)
        if dimensions != 3:
            raise NotImplementedError("JukeboxLoss: dimensions=3 only (the 3-D VQ-VAE of the README); no fallback")
        default = {"s": None, "dim": tuple(range(1, dimensions + 2)), "norm": "ortho"}
        if fft_kwargs is not None and dict(fft_kwargs) != default:
            raise NotImplementedError(f"JukeboxLoss: only the default fft_kwargs {default} are implemented; no fallback")
        self.dimensions = dimensions
        self.include_pixel_loss = include_pixel_loss
        self.fft_factor: float = 1.0
        self.fft_kwargs = default
        self.summaries: Dict = {"scalar": dict()}

    def forward(self, network_output: Dict[str, List[torch.Tensor]], y: torch.Tensor) -> torch.Tensor:
        y_pred = network_output["reconstruction"][0]
        q_losses = network_output["quantization_losses"]
        loss = spectral_loss(y_pred, y) * self.fft_factor                                   # :599
        self.summaries["scalar"]["Loss-Spectral-Reconstruction"] = loss.detach()
        self.summaries["scalar"]["Auxiliary-FFT_Factor"] = self.fft_factor
        if self.include_pixel_loss:                                                         # :603-607
            l2_loss = mse_loss(y_pred, y)
            self.summaries["scalar"]["Loss-MSE-Reconstruction"] = l2_loss.detach()
            loss = loss + l2_loss
        for idx, q_loss in enumerate(q_losses):                                             # :609-616
            q_loss = q_loss.float()
            self.summaries["scalar"][f"Loss-MSE-VQ{idx}_Commitment_Cost"] = q_loss.detach()
            loss = loss + q_loss
        return loss

    def get_summaries(self) -> Dict[str, torch.Tensor]:
        return self.summaries

    def get_fft_factor(self) -> float:
        return self.fft_factor

    def set_fft_factor(self, fft_factor: float) -> float:
        self.fft_factor = fft_factor
        return self.get_fft_factor()


class _CEFn(torch.autograd.Function):
    """F.cross_entropy(input [B, V, N], target [B, N]) with one fused kernel per direction (no log-softmax tensor)."""

    @staticmethod
    def forward(ctx, y_pred, y, reduction):
        if not y_pred.is_cuda:
            raise RuntimeError("synthanatomy_b200: CUDA tensors only -- there is no CPU fallback")
        logits = y_pred.detach().float().transpose(1, 2).contiguous()      # [B, N, V]; a view for TransformerTrainingInferer output
        B, N, V = logits.shape
        target = y.detach().long().contiguous().view(-1)
        loss_sum = torch.zeros((1,), device=logits.device, dtype=torch.float32)
        pf_ops.ce_fwd_bwd(logits.view(B * N, V), target, 0.0, None, loss_sum, None)
        ctx.save_for_backward(logits, target)
        ctx.scale = 1.0 / (B * N) if reduction == "mean" else 1.0
        return loss_sum[0] * ctx.scale

    @staticmethod
    def backward(ctx, g):
        logits, target = ctx.saved_tensors
        B, N, V = logits.shape
        dl = torch.empty_like(logits)
        pf_ops.ce_fwd_bwd(logits.view(B * N, V), target, ctx.scale, g.float().contiguous().view(1), None, dl.view(B * N, V))
        return dl.transpose(1, 2), None, None


class CELoss(_Loss):
    def __init__(self, weight=None, size_average: bool = None, reduce: bool = None, reduction: str = "mean"):
        super().__init__(size_average, reduce, reduction)
        if reduction not in ("sum", "mean"):
            raise ValueError("Reduction must be either 'sum' or 'mean'")
        if weight is not None:
            raise NotImplementedError("class weights are not implemented (the reference passes none); no fallback")
        self.summaries: Dict = {"scalar": dict()}

    def forward(self, y_pred: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        loss = _CEFn.apply(y_pred, y, self.reduction)
        self.summaries["scalar"]["Loss-CE-Prediction"] = loss.detach()
        return loss

    def get_summaries(self) -> Dict[str, torch.Tensor]:
        return self.summaries


# ------------------------------------------------------------------------------------------------
# adversarial criteria on the discriminator's patch map (/root/reference/src/losses/adversarial/adversarial.py:11-122).
# The map is ~7e4 values at the README size: a handful of elementwise torch ops, device-agnostic host logic.
# ------------------------------------------------------------------------------------------------
ADVERSARIAL_CRITERIA = ("vanilla", "hinge", "least_square")          # src/losses/adversarial/utils.py


def adversarial_criterion(name: str):
    """per-element loss(logits, is_real); note the reference's naming: "vanilla" is the relu hinge, "hinge" the softplus
    form (adversarial.py:85-120)"""
    sign = lambda is_real: -1.0 if is_real else 1.0                   # noqa: E731
    if name == "vanilla":
        return lambda logits, is_real: torch.relu(1.0 + sign(is_real) * logits)
    if name == "hinge":
        return lambda logits, is_real: torch.nn.functional.softplus(sign(is_real) * logits)
    if name == "least_square":
        return lambda logits, is_real: (logits - (1.0 if is_real else 0.0)) ** 2
    raise ValueError(f"Unknown adversarial loss. Available losses are {list(ADVERSARIAL_CRITERIA)} but received {name}")


class AdversarialLoss(_Loss):
    """generator side (is_discriminator=False): weight * mean(criterion(D(fake), real=True));
    discriminator side: weight * 0.5 * (mean(criterion(D(fake), False)) + mean(criterion(D(real), True)))"""

    def __init__(self, criterion: str = "least_square", is_discriminator: bool = True, weight=None, size_average: bool = None,
                 reduce: bool = None, reduction: str = "mean"):
        super().__init__(size_average, reduce, reduction)
        if reduction not in ("sum", "mean"):
            raise ValueError("Reduction must be either 'sum' or 'mean'")
        self.criterion = criterion
        self.is_discriminator = is_discriminator
        self.criterion_function = adversarial_criterion(criterion)
        self._weight = weight
        self.summaries: Dict = {"scalar": dict()}

    def forward(self, logits_fake: torch.Tensor, logits_real: torch.Tensor = None) -> torch.Tensor:
        side = "Discriminator" if self.is_discriminator else "Generator"
        loss = torch.mean(self.criterion_function(logits_fake.float(), not self.is_discriminator))
        self.summaries["scalar"][f"Loss-Adversarial_{side}-Reconstruction"] = loss.detach()
        if self.is_discriminator:
            loss_real = torch.mean(self.criterion_function(logits_real.float(), True))
            self.summaries["scalar"]["Loss-Adversarial_Discriminator-Originals"] = loss_real.detach()
            loss = 0.5 * (loss + loss_real)
        return self._weight * loss

    def get_summaries(self) -> Dict[str, torch.Tensor]:
        return self.summaries

    def get_weight(self) -> float:
        return self._weight

    def set_weight(self, weight: float) -> float:
        self._weight = weight
        return self.get_weight()


def get_discriminator_loss(config: dict) -> AdversarialLoss:
    """src/losses/adversarial/configure.py:10-22"""
    adversarial_criterion(config["discriminator_loss"])
    return AdversarialLoss(criterion=config["discriminator_loss"], is_discriminator=True, weight=0.005)


def get_generator_loss(config: dict) -> AdversarialLoss:
    """src/losses/adversarial/configure.py:25-37"""
    adversarial_criterion(config["generator_loss"])
    return AdversarialLoss(criterion=config["generator_loss"], is_discriminator=False, weight=0.005)
