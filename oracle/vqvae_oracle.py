"""CPU oracle for the VQ-VAE hot path (TEST INFRASTRUCTURE -- never imported by the product).

A functional restatement of ``/root/reference/src/networks/vqvae/baseline.py`` that works
on a plain ``state_dict`` (same keys/shapes as the reference ``BaselineVQVAE``), so that it
can be run on the GPU box where ``/root/reference`` does not exist.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified reference
``BaselineVQVAE`` (with a stub for the unused ``monai`` import) in the build container and
stores its inputs/outputs/grads under ``tests/golden/``; ``tests/test_oracle.py`` checks this
restatement against those fixtures bit-for-bit (same torch CPU kernels => exact equality)
and checks ``conv3d_naive`` / ``conv_transpose3d_naive`` (pure numpy loops) against the
``torch.nn.functional`` calls used here, so the oracle does not silently depend on mkldnn.

Each function cites the reference lines it follows.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------
# configuration mirror of BaselineVQVAE.__init__ (baseline.py:164-211)
# ----------------------------------------------------------------------------------------
class VQVAEConfig:
    def __init__(
        self,
        n_levels: int = 3,
        downsample_parameters: Sequence[Sequence[int]] = ((4, 2, 1, 1),) * 3,
        upsample_parameters: Sequence[Sequence[int]] = ((4, 2, 1, 0, 1),) * 3,
        n_embed: int = 256,
        embed_dim: int = 256,
        n_channels: int = 144,
        n_res_channels: int = 144,
        n_res_layers: int = 3,
        p_dropout: float = 0.0,
        commitment_cost: float = 0.25,
        vq_decay: float = 0.5,
        eps: float = 1e-5,
    ):
        assert n_levels == len(downsample_parameters) == len(upsample_parameters)
        assert p_dropout == 0.0, "oracle restates the p=0 path only (README config)"
        self.n_levels = n_levels
        self.downsample_parameters = tuple(tuple(p) for p in downsample_parameters)
        self.upsample_parameters = tuple(tuple(p) for p in upsample_parameters)
        self.n_embed = n_embed
        self.embed_dim = embed_dim
        self.n_channels = n_channels
        self.n_res_channels = n_res_channels
        self.n_res_layers = n_res_layers
        self.commitment_cost = commitment_cost
        self.vq_decay = vq_decay
        self.eps = eps  # Quantizer.__init__ default, baseline.py:95


# ----------------------------------------------------------------------------------------
# layer programme: the nn.Sequential indices the reference builds
# ----------------------------------------------------------------------------------------
def encoder_program(cfg: VQVAEConfig) -> List[Tuple]:
    """construct_encoder, baseline.py:213-246.  Returns [(kind, seq_index, meta)]."""
    prog = []
    idx = 0
    for i in range(cfg.n_levels):
        cin = 1 if i == 0 else cfg.n_channels // 2
        cout = cfg.n_channels // (1 if i == cfg.n_levels - 1 else 2)
        k, s, p, d = cfg.downsample_parameters[i]
        prog.append(("conv", idx, dict(cin=cin, cout=cout, k=k, s=s, p=p, d=d, relu=True)))
        idx += 2  # conv, ReLU
        prog.append(("res", idx, dict(c=cout, n=cfg.n_res_layers)))
        idx += 1
    prog.append(("conv", idx, dict(cin=cfg.n_channels, cout=cfg.embed_dim, k=3, s=1, p=1, d=1, relu=False)))
    return prog


def decoder_program(cfg: VQVAEConfig) -> List[Tuple]:
    """construct_decoder, baseline.py:257-299."""
    prog = [("conv", 0, dict(cin=cfg.embed_dim, cout=cfg.n_channels, k=3, s=1, p=1, d=1, relu=False))]
    idx = 1
    for i in range(cfg.n_levels):
        c = cfg.n_channels // (1 if i == 0 else 2)
        prog.append(("res", idx, dict(c=c, n=cfg.n_res_layers)))
        idx += 1
        k, s, p, op, d = cfg.upsample_parameters[i]
        last = i == cfg.n_levels - 1
        cout = 1 if last else cfg.n_channels // 2
        prog.append(("deconv", idx, dict(cin=c, cout=cout, k=k, s=s, p=p, op=op, d=d, relu=not last)))
        idx += 1 if last else 2
    return prog


# ----------------------------------------------------------------------------------------
# numpy reference loops (tiny sizes only) used to pin the F.conv* calls below
# ----------------------------------------------------------------------------------------
def conv3d_naive(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray], stride: int, pad: int) -> np.ndarray:
    """y[n,co,o] = b[co] + sum_{ci,t} x[n,ci,o*s-p+t] w[co,ci,t]  (nn.Conv3d semantics, cross-correlation)."""
    n, ci, d, h, wd = x.shape
    co, _, k, _, _ = w.shape
    xp = np.zeros((n, ci, d + 2 * pad, h + 2 * pad, wd + 2 * pad), dtype=np.float64)
    xp[:, :, pad:pad + d, pad:pad + h, pad:pad + wd] = x
    od = (d + 2 * pad - k) // stride + 1
    oh = (h + 2 * pad - k) // stride + 1
    ow = (wd + 2 * pad - k) // stride + 1
    y = np.zeros((n, co, od, oh, ow), dtype=np.float64)
    for kd in range(k):
        for kh in range(k):
            for kw in range(k):
                patch = xp[:, :, kd:kd + od * stride:stride, kh:kh + oh * stride:stride, kw:kw + ow * stride:stride]
                y += np.einsum("ncdhw,oc->nodhw", patch, w[:, :, kd, kh, kw].astype(np.float64))
    if b is not None:
        y += b.reshape(1, -1, 1, 1, 1)
    return y


def conv_transpose3d_naive(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray], stride: int, pad: int) -> np.ndarray:
    """y[n,co,i*s-p+t] += x[n,ci,i] w[ci,co,t]  (nn.ConvTranspose3d semantics, output_padding=0)."""
    n, ci, d, h, wd = x.shape
    _, co, k, _, _ = w.shape
    fd, fh, fw = (d - 1) * stride + k, (h - 1) * stride + k, (wd - 1) * stride + k
    full = np.zeros((n, co, fd, fh, fw), dtype=np.float64)
    for kd in range(k):
        for kh in range(k):
            for kw in range(k):
                contrib = np.einsum("ncdhw,co->nodhw", x.astype(np.float64), w[:, :, kd, kh, kw].astype(np.float64))
                full[:, :, kd:kd + d * stride:stride, kh:kh + h * stride:stride, kw:kw + wd * stride:stride] += contrib
    y = full[:, :, pad:fd - pad, pad:fh - pad, pad:fw - pad]
    if b is not None:
        y = y + b.reshape(1, -1, 1, 1, 1)
    return y


# ----------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------
def residual_layer(x: torch.Tensor, w3, b3, w1, b1) -> torch.Tensor:
    """ResidualLayer.forward, baseline.py:150-160 (Dropout3d p=0 is the identity)."""
    h = F.relu(F.conv3d(x, w3, b3, padding=1))
    h = F.conv3d(h, w1, b1)
    return F.relu(x + h)


def encode(sd: Dict[str, torch.Tensor], cfg: VQVAEConfig, images: torch.Tensor) -> torch.Tensor:
    """BaselineVQVAE.encode, baseline.py:329-330."""
    x = images
    for kind, i, m in encoder_program(cfg):
        p = f"encoder.0.{i}"
        if kind == "conv":
            x = F.conv3d(x, sd[f"{p}.weight"], sd[f"{p}.bias"], stride=m["s"], padding=m["p"], dilation=m["d"])
            if m["relu"]:
                x = F.relu(x)
        else:
            for r in range(m["n"]):
                x = residual_layer(x, sd[f"{p}.{r}.0.weight"], sd[f"{p}.{r}.0.bias"],
                                   sd[f"{p}.{r}.3.weight"], sd[f"{p}.{r}.3.bias"])
    return x


def decode(sd: Dict[str, torch.Tensor], cfg: VQVAEConfig, q: torch.Tensor) -> torch.Tensor:
    """BaselineVQVAE.decode, baseline.py:338-340."""
    x = q
    for kind, i, m in decoder_program(cfg):
        p = f"decoder.0.{i}"
        if kind == "conv":
            x = F.conv3d(x, sd[f"{p}.weight"], sd[f"{p}.bias"], stride=m["s"], padding=m["p"])
        elif kind == "deconv":
            x = F.conv_transpose3d(x, sd[f"{p}.weight"], sd[f"{p}.bias"], stride=m["s"], padding=m["p"],
                                   output_padding=m["op"], dilation=m["d"])
            if m["relu"]:
                x = F.relu(x)
        else:
            for r in range(m["n"]):
                x = residual_layer(x, sd[f"{p}.{r}.0.weight"], sd[f"{p}.{r}.0.bias"],
                                   sd[f"{p}.{r}.3.weight"], sd[f"{p}.{r}.3.bias"])
    return x


def vq_distances(flat: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """baseline.py:49-53: ||x||^2 - 2 x W^T + ||W||^2, fp32, in that association order."""
    return (
        (flat ** 2).sum(dim=1, keepdim=True)
        - 2 * torch.mm(flat, weight.t())
        + (weight ** 2).sum(dim=1, keepdim=True).t()
    )


def vq_argmin(flat: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """baseline.py:56: argmax of -d (first index on ties)."""
    return torch.max(-vq_distances(flat, weight), dim=1)[1]


def vq_argmin_exact(flat: np.ndarray, weight: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """float64 'true' nearest code and the best/second-best gap; used by tests to decide which
    rows are numerically decidable (gap well above fp32 rounding of the expansion)."""
    f = flat.astype(np.float64)
    w = weight.astype(np.float64)
    d = (f * f).sum(1, keepdims=True) - 2.0 * f @ w.T + (w * w).sum(1)[None, :]
    idx = d.argmin(1)
    part = np.partition(d, 1, axis=1)
    return idx, part[:, 1] - part[:, 0]


def quantize(
    sd: Dict[str, torch.Tensor],
    cfg: VQVAEConfig,
    x: torch.Tensor,
    training: bool,
    decay: Optional[float] = None,
    commitment_cost: Optional[float] = None,
    all_reduce=None,
):
    """Quantizer_impl.forward, baseline.py:38-87.

    Returns (quantized_st, latent_loss, embed_idx, new_state) where new_state holds the
    post-EMA ``N``, ``embed_avg``, ``weight`` (the reference mutates its buffers in place,
    baseline.py:75-80; the oracle returns them instead).  ``all_reduce`` is an optional
    callable(tensor)->tensor standing in for dist.all_reduce(SUM) (baseline.py:70-72).
    """
    decay = cfg.vq_decay if decay is None else decay
    beta = cfg.commitment_cost if commitment_cost is None else commitment_cost
    weight = sd["quantizer.0.impl.weight"]
    b, c, h, w, d = x.shape
    x = x.float()
    flat = x.permute(0, 2, 3, 4, 1).contiguous().view(-1, cfg.embed_dim)            # :46
    embed_idx = vq_argmin(flat, weight)                                             # :49-56
    onehot = F.one_hot(embed_idx, cfg.n_embed).type_as(flat)                        # :57
    embed_idx = embed_idx.view(b, h, w, d)                                          # :60
    quantized = F.embedding(embed_idx, weight).permute(0, 4, 1, 2, 3).contiguous()  # :63
    new_state = None
    if training:                                                                    # :66
        with torch.no_grad():
            enc_sum = onehot.sum(0)                                                 # :68
            dw = torch.mm(onehot.t(), flat.detach())                                # :69
            if all_reduce is not None:
                enc_sum = all_reduce(enc_sum)
                dw = all_reduce(dw)
            N = sd["quantizer.0.impl.N"] * decay + enc_sum * (1 - decay)            # :75
            embed_avg = sd["quantizer.0.impl.embed_avg"] * decay + dw * (1 - decay)  # :76
            n = N.sum()                                                             # :78
            W = (N + cfg.eps) / (n + cfg.n_embed * cfg.eps) * n                     # :79
            new_weight = embed_avg / W.unsqueeze(1)                                 # :80
            new_state = {"N": N, "embed_avg": embed_avg, "weight": new_weight}
    # NOTE (:63 vs :80): `quantized` was gathered BEFORE the EMA update, so the loss and the
    # decoder input use the pre-update codebook.
    latent_loss = beta * F.mse_loss(quantized.detach(), x)                          # :82
    quantized_st = (quantized - x).detach() + x                                     # :85
    return quantized_st, latent_loss, embed_idx, new_state


def perplexity(embed_idx: torch.Tensor, n_embed: int) -> torch.Tensor:
    """Quantizer.forward, baseline.py:110-120."""
    avg = torch.histc(embed_idx.float(), bins=n_embed, max=n_embed).float().div(embed_idx.numel())
    return torch.exp(-torch.sum(avg * torch.log(avg + 1e-10)))


def embed(sd: Dict[str, torch.Tensor], idx: torch.Tensor) -> torch.Tensor:
    """Quantizer_impl.embed, baseline.py:89-91."""
    return F.embedding(idx, sd["quantizer.0.impl.weight"]).permute(0, 4, 1, 2, 3).contiguous()


def forward(sd, cfg: VQVAEConfig, images: torch.Tensor, training: bool = True, all_reduce=None):
    """BaselineVQVAE.forward, baseline.py:354-362.  Returns dict with the reference's two keys
    plus the oracle-only extras (indices, encodings, new codebook state)."""
    z = encode(sd, cfg, images)
    q_st, q_loss, idx, new_state = quantize(sd, cfg, z, training, all_reduce=all_reduce)
    recon = decode(sd, cfg, q_st)
    return {
        "reconstruction": [recon],
        "quantization_losses": [q_loss],
        "indices": idx,
        "encodings": z,
        "new_state": new_state,
    }


def train_step_grads(sd, cfg: VQVAEConfig, images: torch.Tensor, all_reduce=None):
    """loss = mse(recon, x) + q_loss (MSELoss.forward, src/losses/vqvae/vqvae.py:52-70);
    returns (loss, grads dict, forward outputs).  Codebook / buffers take no grad."""
    leaf = {}
    for k, v in sd.items():
        if k.startswith("quantizer."):
            leaf[k] = v.detach().clone()
        else:
            leaf[k] = v.detach().clone().requires_grad_(True)
    out = forward(leaf, cfg, images, training=True, all_reduce=all_reduce)
    loss = F.mse_loss(out["reconstruction"][0].float(), images.float()) + out["quantization_losses"][0].float()
    loss.backward()
    grads = {k: v.grad.detach().clone() for k, v in leaf.items() if v.requires_grad}
    return loss.detach(), grads, out


def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int,
              lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
    """torch.optim.Adam (run_vqvae.py:82; no weight decay, no amsgrad), single tensor."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / (bc2 ** 0.5)) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def init_state_dict(cfg: VQVAEConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random parameters with the reference's key set and shapes.  The values are
    NOT the reference's init stream; goldens store the reference's own tensors instead."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(prefix, cout, cin, k, transposed=False):
        fan_in = (cout if transposed else cin) * k ** 3
        bound = 1.0 / fan_in ** 0.5
        shape = (cin, cout, k, k, k) if transposed else (cout, cin, k, k, k)
        sd[f"{prefix}.weight"] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[f"{prefix}.bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    for kind, i, m in encoder_program(cfg):
        if kind == "conv":
            conv(f"encoder.0.{i}", m["cout"], m["cin"], m["k"])
        else:
            for r in range(m["n"]):
                conv(f"encoder.0.{i}.{r}.0", m["c"], m["c"], 3)
                conv(f"encoder.0.{i}.{r}.3", m["c"], m["c"], 1)
    w = torch.randn(cfg.n_embed, cfg.embed_dim, generator=g)
    sd["quantizer.0.impl.weight"] = w
    sd["quantizer.0.impl.N"] = torch.zeros(cfg.n_embed)
    sd["quantizer.0.impl.embed_avg"] = w.clone()
    sd["quantizer.0.impl.embedding.weight"] = w
    for kind, i, m in decoder_program(cfg):
        if kind == "conv":
            conv(f"decoder.0.{i}", m["cout"], m["cin"], m["k"])
        elif kind == "deconv":
            conv(f"decoder.0.{i}", m["cout"], m["cin"], m["k"], transposed=True)
        else:
            for r in range(m["n"]):
                conv(f"decoder.0.{i}.{r}.0", m["c"], m["c"], 3)
                conv(f"decoder.0.{i}.{r}.3", m["c"], m["c"], 1)
    return sd
