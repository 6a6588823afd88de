"""CPU oracle for the Performer prior hot path (TEST INFRASTRUCTURE -- never imported by the product).

What is restated, in call order:

* ``/root/reference/src/networks/transformers/performer.py:229-288`` (in-tree wrapper: token embedding,
  three absolute spatial-position embeddings with a zero BOS row, absolute position embedding, the
  third-party ``performer_pytorch.Performer`` stack, LayerNorm, ``to_out``) -- restated line by line.
* ``/root/reference/src/utils/transformer.py:259-282`` (``prepare_batch``: flatten, gather by the ordering,
  left-pad BOS = vocab_size, shift) and ``/root/reference/src/losses/transformer/transformer.py:24-33`` (CE, mean)
  with the ``[B, V, N]`` transpose of ``/root/reference/src/inferer/transformer.py:28-29``.
* the THIRD-PARTY arithmetic that ``performer.py:194-219`` constructs and ``:270`` calls.  Those packages are
  absent from /root/reference and from this image (no network): ``performer-pytorch==1.0.11``
  (pinned, docker/requirements.txt:10), ``local-attention`` (unpinned transitive dependency, mid-2021 = rotary
  ``SinusoidalEmbeddings`` inside the local heads) and ``pytorch-fast-transformers`` (unpinned,
  docker/Dockerfile:20; ``CausalDotProduct``).  Their published algorithms are restated here from the
  package sources as released (function names kept so that a maintainer can diff):
  ``softmax_kernel``, ``gaussian_orthogonal_random_matrix`` / ``orthogonal_matrix_chunk``,
  ``causal_linear_attention`` (+ ``CausalDotProduct`` = prefix sum of k (x) v), ``LocalAttention.forward``
  (``look_around``, ``apply_rotary_pos_emb``), ``SelfAttention.forward``, ``FeedForward``, ``ReZero``,
  ``SequentialSequence``, ``ProjectionUpdater.redraw_projections``.

PARITY STATUS: **parity unpinned** for the third-party part -- the reference holds no tests, golden vectors
or fixtures for this path and the packages cannot be executed here.  What IS pinned:
``ordering_restated`` and ``prepare_batch`` against the unmodified reference modules (they import cleanly;
``oracle/make_golden.py`` stores their outputs in tests/golden/performer_host.npz), the chunk-free O(N^2)
definitions below against each other (tests/test_oracle_performer.py: causal prefix-sum == masked quadratic form,
local attention == dense masked softmax, strict causality of the local heads, the documented non-causal coupling
through the global key stabiliser).

All tensors are torch CPU fp32; gradients come from torch autograd over this restatement.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------
# configuration mirror of Performer.__init__ (performer.py:75-115), README.md:124-141 values as defaults
# ----------------------------------------------------------------------------------------
@dataclass
class PerformerConfig:
    num_tokens: int = 2049
    max_seq_len: int = 1401
    dim: int = 512
    depth: int = 24
    heads: int = 16
    dim_head: int = 64                      # wrapper default, performer.py:84
    local_attn_heads: int = 8
    local_window_size: int = 420
    ff_mult: int = 4
    nb_features: Optional[int] = None       # default int(d * ln d) = 266 for d = 64 (performer-pytorch FastAttention)
    spatial_shape: Optional[Tuple[int, ...]] = None
    spatial_position_emb: Optional[str] = "absolute"
    use_rezero: bool = True
    local_rel_pos: str = "rotary"           # local-attention >= 1.1 (mid-2021); "none" = no positional term
    key_stabiliser: str = "global"          # 1.0.11: torch.max over the whole key tensor; "per_head" = later 1.1.x
    eps_feature: float = 1e-4               # softmax_kernel eps
    eps_cumsum: float = 1e-6                # causal_linear_attention eps
    fixed_position_emb: bool = False        # performer.py:136-138: sinusoidal pos_emb (performer-pytorch FixedPositionalEmbedding)
    conditioning_num_tokens: Optional[Tuple[int, ...]] = None     # performer.py:183-187
    conditioning_type: str = "none"         # "none" | "bos_replacement" | "prepending" (src/utils/transformer.py:21-24)

    @property
    def inner(self) -> int:
        return self.heads * self.dim_head

    @property
    def global_heads(self) -> int:
        return self.heads - self.local_attn_heads

    @property
    def m(self) -> int:
        return self.nb_features if self.nb_features is not None else int(self.dim_head * math.log(self.dim_head))


# ----------------------------------------------------------------------------------------
# host-side sequence logic
# ----------------------------------------------------------------------------------------
def ordering_restated(ordering_type: str, dimensions: Sequence[int], reflected: Sequence[bool],
                      transpositions_axes: Sequence[Sequence[int]], rot90_axes: Sequence[Sequence[int]],
                      transformation_order: Sequence[str] = ("transpose", "rotate_90", "reflect")) -> np.ndarray:
    """img2seq_ordering.py:24-140: template = arange(prod).reshape(spatial) -> transformations in the given
    order -> read out along the scan path.  Returns ``_sequence_ordering``."""
    spatial = tuple(dimensions[1:])
    template = np.arange(int(np.prod(spatial))).reshape(*spatial)                     # :86-90
    for tr in transformation_order:                                                   # :92-101
        if tr == "transpose":
            for axes in transpositions_axes:                                          # :102-106
                template = np.transpose(template, axes=axes)
        elif tr == "rotate_90":
            for axes in rot90_axes:                                                   # :114-118
                template = np.rot90(template, axes=axes)
        elif tr == "reflect":
            for axis, flag in enumerate(reflected):                                   # :108-112
                template = np.flip(template, axis=axis) if flag else template
        else:
            raise ValueError(tr)
    if ordering_type == "raster_scan":                                                # :143-157
        return np.ascontiguousarray(template).reshape(-1).copy()
    if ordering_type == "s_curve":                                                    # :159-179 (3-D: depth direction
        out = []                                                                      #  flips on odd COLUMN index)
        rows, cols = template.shape[0], template.shape[1]
        for r in range(rows):
            col_idx = range(cols) if r % 2 == 0 else range(cols - 1, -1, -1)
            for c in col_idx:
                if template.ndim == 3:
                    line = template[r, c]
                    out.extend(line if c % 2 == 0 else line[::-1])
                else:
                    out.append(template[r, c])
        return np.array(out)
    if ordering_type == "hilbert_curve":                                              # :196-201
        coords = list(gilbert_curve(template.shape))
        return np.array([template[c] for c in coords])
    raise NotImplementedError(ordering_type)


def gilbert_curve(shape: Sequence[int]):
    """Generalised Hilbert curve over a rectangle / cuboid: the visiting order the reference obtains from the vendored
    third-party generators gilbert2d / gilbert3d (J. Cerveny, BSD-2-Clause; img2seq_ordering.py:7-8,196-201), restated as
    ONE recursion on tuples for both ranks.  A box = corner + axis vectors (major first).  Pinned by the golden sequences
    of oracle/make_golden_performer.py, which come from the vendored generators through the reference's Ordering class."""
    def add(*vs):
        return tuple(sum(c) for c in zip(*vs))

    def neg(v):
        return tuple(-c for c in v)

    def sub(u, v):
        return add(u, neg(v))

    def unit(v):
        return tuple((c > 0) - (c < 0) for c in v)

    def size(v):
        return abs(sum(v))

    def half(v, even_above_2):
        h = tuple(c // 2 for c in v)                       # floor, also for negative components
        if size(h) % 2 and even_above_2:
            h = add(h, unit(v))
        return h

    def rec(p, vs):
        n = [size(v) for v in vs]
        thick = [i for i, k in enumerate(n) if k != 1]
        if len(thick) <= 1:
            i = thick[0] if thick else 0
            for _ in range(n[i]):
                yield p
                p = add(p, unit(vs[i]))
            return
        if len(vs) == 2:
            a, b = vs
            w, h = n
            if 2 * w > 3 * h:
                a2 = half(a, w > 2)
                parts = [(p, (a2, b)), (add(p, a2), (sub(a, a2), b))]
            else:
                a2, b2 = tuple(c // 2 for c in a), half(b, h > 2)
                parts = [(p, (b2, a2)), (add(p, b2), (a, sub(b, b2))),
                         (add(p, sub(a, unit(a)), sub(b2, unit(b))), (neg(b2), neg(sub(a, a2))))]
        else:
            a, b, c = vs
            w, h, d = n
            a2, b2, c2 = half(a, w > 2), half(b, h > 2), half(c, d > 2)
            ea, eb, ec = sub(a, unit(a)), sub(b2, unit(b)), sub(c, unit(c))
            if 2 * w > 3 * h and 2 * w > 3 * d:
                parts = [(p, (a2, b, c)), (add(p, a2), (sub(a, a2), b, c))]
            elif 3 * h > 4 * d:
                parts = [(p, (b2, c, a2)), (add(p, b2), (a, sub(b, b2), c)), (add(p, ea, eb), (neg(b2), c, neg(sub(a, a2))))]
            elif 3 * d > 4 * h:
                parts = [(p, (c2, a2, b)), (add(p, c2), (a, b, sub(c, c2))),
                         (add(p, ea, sub(c2, unit(c))), (neg(c2), neg(sub(a, a2)), b))]
            else:
                parts = [(p, (b2, c2, a2)), (add(p, b2), (c, a2, sub(b, b2))),
                         (add(p, eb, ec), (a, neg(b2), neg(sub(c, c2)))),
                         (add(p, ea, b2, ec), (neg(c), neg(sub(a, a2)), sub(b, b2))),
                         (add(p, ea, eb), (neg(b2), c2, neg(sub(a, a2))))]
        for q, sub_vs in parts:
            yield from rec(q, sub_vs)

    rank = len(shape)
    lead = max(range(rank), key=lambda i: (shape[i], -i))
    order = [lead] + [i for i in range(rank) if i != lead]
    vecs = tuple(tuple(shape[i] if j == i else 0 for j in range(rank)) for i in order)
    yield from rec((0,) * rank, vecs)


def prepare_batch(quantization: np.ndarray, index_sequence: np.ndarray, vocab_size: int):
    """src/utils/transformer.py:259-282 -> (x_input, x_target) int64 [B, N]."""
    enc = quantization.reshape(quantization.shape[0], -1)          # :259-260
    enc = enc[:, index_sequence]                                   # :261
    enc = np.pad(enc, ((0, 0), (1, 0)), constant_values=vocab_size).astype(np.int64)   # :262-263
    return enc[:, :-1], enc[:, 1:]                                 # :279-280


def spatial_index_sequences(spatial_shape: Sequence[int], sequence_ordering: np.ndarray) -> List[np.ndarray]:
    """performer.py:159-176: per axis, the coordinate value of every sequence position (ordering applied)."""
    coords = np.array(np.meshgrid(*tuple(np.arange(0, s) for s in spatial_shape), indexing="ij"))
    return [coords[i].flatten()[sequence_ordering] for i in range(len(spatial_shape))]


# ----------------------------------------------------------------------------------------
# performer-pytorch 1.0.11: random features
# ----------------------------------------------------------------------------------------
def orthogonal_matrix_chunk(cols: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    block = torch.randn((cols, cols), generator=generator)
    q, _ = torch.linalg.qr(block, mode="reduced")      # torch.qr(some=True) in the 2021 source
    return q.t()


def gaussian_orthogonal_random_matrix(nb_rows: int, nb_columns: int, scaling: int = 0,
                                      generator: Optional[torch.Generator] = None) -> torch.Tensor:
    nb_full_blocks = int(nb_rows / nb_columns)
    blocks = [orthogonal_matrix_chunk(nb_columns, generator) for _ in range(nb_full_blocks)]
    remaining = nb_rows - nb_full_blocks * nb_columns
    if remaining > 0:
        blocks.append(orthogonal_matrix_chunk(nb_columns, generator)[:remaining])
    final = torch.cat(blocks)
    if scaling == 0:
        multiplier = torch.randn((nb_rows, nb_columns), generator=generator).norm(dim=1)
    elif scaling == 1:
        multiplier = math.sqrt(float(nb_columns)) * torch.ones((nb_rows,))
    else:
        raise ValueError(scaling)
    return torch.diag(multiplier) @ final


def softmax_kernel(data: torch.Tensor, projection_matrix: torch.Tensor, is_query: bool, eps: float = 1e-4,
                   key_stabiliser: str = "global") -> torch.Tensor:
    """data [B, H, N, d], projection [m, d] -> [B, H, N, m] (performer-pytorch 1.0.11 ``softmax_kernel``,
    normalize_data=True).  The max terms are NOT detached (1.0.11 does not detach them)."""
    d = data.shape[-1]
    data_normalizer = d ** -0.25
    ratio = projection_matrix.shape[0] ** -0.5
    data_dash = torch.einsum("...id,jd->...ij", data_normalizer * data, projection_matrix)
    diag_data = (data ** 2).sum(dim=-1)
    diag_data = (diag_data / 2.0) * (data_normalizer ** 2)
    diag_data = diag_data.unsqueeze(-1)
    if is_query:
        return ratio * (torch.exp(data_dash - diag_data - torch.max(data_dash, dim=-1, keepdim=True).values) + eps)
    if key_stabiliser == "global":
        return ratio * (torch.exp(data_dash - diag_data - torch.max(data_dash)) + eps)
    return ratio * (torch.exp(data_dash - diag_data - torch.amax(data_dash, dim=(-1, -2), keepdim=True).detach()) + eps)


def causal_dot_product(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, chunk: int = 128) -> torch.Tensor:
    """fast_transformers.causal_product.CausalDotProduct: out[n] = q[n] . sum_{j<=n} k[j] (x) v[j].
    Evaluated chunk-wise (exactly the same sum, associativity aside) so that N = 14 000 fits in memory."""
    B, H, N, m = q.shape
    e = v.shape[-1]
    state = q.new_zeros(B, H, m, e)
    outs = []
    for s in range(0, N, chunk):
        qc, kc, vc = q[:, :, s:s + chunk], k[:, :, s:s + chunk], v[:, :, s:s + chunk]
        a = torch.einsum("bhim,bhjm->bhij", qc, kc).tril()
        outs.append(torch.einsum("bhij,bhje->bhie", a, vc) + torch.einsum("bhim,bhme->bhie", qc, state))
        state = state + torch.einsum("bhjm,bhje->bhme", kc, vc)
    return torch.cat(outs, dim=2)


def causal_linear_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    k_cumsum = k.cumsum(dim=-2) + eps
    d_inv = 1.0 / torch.einsum("...nd,...nd->...n", q, k_cumsum)
    out = causal_dot_product(q, k, v)
    return torch.einsum("...nd,...n->...nd", out, d_inv)


# ----------------------------------------------------------------------------------------
# local-attention (mid-2021): rotary + bucketed causal window attention
# ----------------------------------------------------------------------------------------
def sinusoidal_embeddings(n: int, dim: int) -> torch.Tensor:
    inv_freq = 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim))
    t = torch.arange(n).float()
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    return torch.cat((freqs, freqs), dim=-1)            # [n, dim]


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb(q, k, freqs):
    return q * freqs.cos() + rotate_half(q) * freqs.sin(), k * freqs.cos() + rotate_half(k) * freqs.sin()


def look_around(x: torch.Tensor, backward: int = 1, pad_value: float = -1.0) -> torch.Tensor:
    """[b, windows, w, ...] -> previous window concatenated in front of each window (dim 2)."""
    t = x.shape[1]
    dims = (len(x.shape) - 2) * (0, 0)
    padded = F.pad(x, (*dims, backward, 0), value=pad_value)
    tensors = [padded[:, ind:(ind + t), ...] for ind in range(backward + 1)]
    return torch.cat(tensors, dim=2)


def local_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, window_size: int,
                    rel_pos: str = "rotary") -> torch.Tensor:
    """LocalAttention(window_size, causal=True, autopad=True, look_backward=1, look_forward=0, dropout=0,
    rel_pos_emb_config=(dim_head, local_heads)).forward(q, k, v); q, k, v: [B, H, N, d]."""
    shape = q.shape
    q, k, v = (t.reshape(-1, *t.shape[-2:]) for t in (q, k, v))
    if rel_pos == "rotary":
        q, k = apply_rotary_pos_emb(q, k, sinusoidal_embeddings(q.shape[1], q.shape[2])[None])
    orig_t = q.shape[1]
    rem = (-orig_t) % window_size
    if rem:                                                                              # autopad
        q, k, v = (F.pad(t, (0, 0, 0, rem), value=0.0) for t in (q, k, v))
    b, t, e = q.shape
    windows = t // window_size
    ticker = torch.arange(t, dtype=q.dtype)[None, :]
    b_t = ticker.reshape(1, windows, window_size)
    bq, bk, bv = (x.reshape(b, windows, window_size, -1) for x in (q, k, v))
    bk, bv = look_around(bk), look_around(bv)
    bq_t = b_t
    bq_k = look_around(b_t)
    dots = torch.einsum("bhie,bhje->bhij", bq, bk) * (e ** -0.5)
    mask_value = -torch.finfo(dots.dtype).max
    dots = dots.masked_fill(bq_t[:, :, :, None] < bq_k[:, :, None, :], mask_value)       # causal
    dots = dots.masked_fill(bq_k[:, :, None, :] == -1, mask_value)                       # look-back padding
    attn = dots.softmax(dim=-1)
    out = torch.einsum("bhij,bhje->bhie", attn, bv).reshape(-1, t, e)
    return out[:, :orig_t, :].reshape(*shape)


def local_attention_dense(q, k, v, window_size: int, rel_pos: str = "rotary") -> torch.Tensor:
    """Same function written as one dense masked softmax over [N, N] (small N only): query p sees keys j with
    (floor(p / w) - 1) * w <= j <= p.  Used to pin the bucketed form above."""
    B, H, N, d = q.shape
    if rel_pos == "rotary":
        f = sinusoidal_embeddings(N, d)[None, None]
        q, k = apply_rotary_pos_emb(q, k, f)
    pos = torch.arange(N)
    lo = (pos // window_size - 1).clamp(min=0) * window_size
    allowed = (pos[None, :] <= pos[:, None]) & (pos[None, :] >= lo[:, None])
    dots = torch.einsum("bhie,bhje->bhij", q, k) * (d ** -0.5)
    dots = dots.masked_fill(~allowed, -torch.finfo(dots.dtype).max)
    return torch.einsum("bhij,bhje->bhie", dots.softmax(-1), v)


# ----------------------------------------------------------------------------------------
# state dict (the reference module tree's keys)
# ----------------------------------------------------------------------------------------
def layer_prefix(i: int) -> str:
    return f"performer.net.layers.{i}."


def init_state_dict(cfg: PerformerConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init parameters / buffers under the reference's state_dict keys (performer.py:117-221 module tree +
    performer-pytorch 1.0.11 sub-modules).  Initialisers follow torch defaults (nn.Embedding N(0,1), nn.Linear
    kaiming-uniform, LayerNorm ones / zeros, ReZero g = 1e-3)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def linear(name, out_f, in_f, bias):
        bound = 1.0 / math.sqrt(in_f)
        sd[name + ".weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
        if bias:
            sd[name + ".bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * bound

    n_seq = cfg.max_seq_len - 1                       # = prod(spatial_shape); the wrapper is built with max_seq_len = N + 1
    n_cond = len(cfg.conditioning_num_tokens) if (cfg.conditioning_num_tokens and cfg.conditioning_type == "prepending") else 0
    sd["token_emb.weight"] = torch.randn(cfg.num_tokens, cfg.dim, generator=g)
    if cfg.fixed_position_emb:
        sd["pos_emb.emb"] = sinusoid_table(torch.arange(0, cfg.max_seq_len + n_cond), cfg.dim)
    else:
        sd["pos_emb.emb.weight"] = torch.randn(cfg.max_seq_len + n_cond, cfg.dim, generator=g)     # performer.py:119-125
    if cfg.spatial_position_emb == "absolute":
        for a in range(len(cfg.spatial_shape)):
            sd[f"spatial_position_emb.{a}.emb.weight"] = torch.randn(n_seq - 1, cfg.dim, generator=g)  # performer.py:27-33
    for i, cnt in enumerate(cfg.conditioning_num_tokens or ()):
        sd[f"conditioning_emb.{i}.weight"] = torch.randn(cnt, cfg.dim, generator=g)
    for i in range(cfg.depth):
        p = layer_prefix(i)
        sd[p + "0.g"] = torch.tensor(1e-3)
        linear(p + "0.fn.to_q", cfg.inner, cfg.dim, False)        # qkv_bias=False, performer.py:109
        linear(p + "0.fn.to_k", cfg.inner, cfg.dim, False)
        linear(p + "0.fn.to_v", cfg.inner, cfg.dim, False)
        linear(p + "0.fn.to_out", cfg.dim, cfg.inner, False)      # attn_out_bias=False, performer.py:110
        if cfg.global_heads > 0:
            sd[p + "0.fn.fast_attention.projection_matrix"] = gaussian_orthogonal_random_matrix(cfg.m, cfg.dim_head, 0, g)
        sd[p + "1.g"] = torch.tensor(1e-3)
        linear(p + "1.fn.fn.w1", cfg.dim * cfg.ff_mult, cfg.dim, True)
        linear(p + "1.fn.fn.w2", cfg.dim, cfg.dim * cfg.ff_mult, True)
    sd["norm.weight"] = torch.ones(cfg.dim)
    sd["norm.bias"] = torch.zeros(cfg.dim)
    linear("to_out", cfg.num_tokens, cfg.dim, True)               # performer.py:221
    return sd


BUFFER_SUFFIXES = ("projection_matrix", "pos_emb.emb")


def trainable_keys(sd: Dict[str, torch.Tensor]) -> List[str]:
    return [k for k in sd if not k.endswith(BUFFER_SUFFIXES)]


# ----------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------
def self_attention(x: torch.Tensor, sd: Dict[str, torch.Tensor], p: str, cfg: PerformerConfig) -> torch.Tensor:
    """performer-pytorch 1.0.11 ``SelfAttention.forward`` (no mask / context / layer_pos_emb at this config)."""
    b, n, _ = x.shape
    h, gh = cfg.heads, cfg.global_heads
    q, k, v = F.linear(x, sd[p + "to_q.weight"]), F.linear(x, sd[p + "to_k.weight"]), F.linear(x, sd[p + "to_v.weight"])
    q, k, v = (t.reshape(b, n, h, cfg.dim_head).permute(0, 2, 1, 3) for t in (q, k, v))      # b n (h d) -> b h n d
    outs = []
    if gh > 0:
        P = sd[p + "fast_attention.projection_matrix"]
        qp = softmax_kernel(q[:, :gh], P, True, cfg.eps_feature)
        kp = softmax_kernel(k[:, :gh], P, False, cfg.eps_feature, cfg.key_stabiliser)
        outs.append(causal_linear_attention(qp, kp, v[:, :gh], cfg.eps_cumsum))
    if h - gh > 0:
        outs.append(local_attention(q[:, gh:], k[:, gh:], v[:, gh:], cfg.local_window_size, cfg.local_rel_pos))
    out = torch.cat(outs, dim=1).permute(0, 2, 1, 3).reshape(b, n, h * cfg.dim_head)          # b h n d -> b n (h d)
    return F.linear(out, sd[p + "to_out.weight"])


def feed_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], p: str) -> torch.Tensor:
    """FeedForward(dim, mult=4, glu=False, dropout=0): w2(gelu(w1(x)))  (exact erf GELU = nn.GELU())."""
    return F.linear(F.gelu(F.linear(x, sd[p + "w1.weight"], sd[p + "w1.bias"])), sd[p + "w2.weight"], sd[p + "w2.bias"])


def sinusoid_table(positions: torch.Tensor, dim: int) -> torch.Tensor:
    """performer.py:46-57 / performer-pytorch FixedPositionalEmbedding: cat(sin(p f), cos(p f)), f = 10000^(-2i/dim)"""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim))
    sinusoid_inp = torch.einsum("i,j->ij", positions.float(), inv_freq)
    return torch.cat((sinusoid_inp.sin(), sinusoid_inp.cos()), dim=-1)


def embed(sd: Dict[str, torch.Tensor], cfg: PerformerConfig, tokens: torch.Tensor,
          spatial_seqs: Optional[Sequence[torch.Tensor]], conditionings: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
    """performer.py:241-268 (dropout 0)."""
    n = tokens.shape[1]
    x = F.embedding(tokens, sd["token_emb.weight"])                                  # :241
    if cfg.spatial_position_emb == "absolute":
        for a, seq in enumerate(spatial_seqs):                                       # :243-244
            sc = F.embedding(seq[:-1], sd[f"spatial_position_emb.{a}.emb.weight"])   # :27-33, 36
            sc = sc[None, : n - 1, :]                                                # :37
            sc = F.pad(sc, (0, 0, 1, 0, 0, 0), "constant", 0)                        # :31, 38 (zero BOS row)
            x = x + sc
    elif cfg.spatial_position_emb == "fixed":
        for seq in spatial_seqs:                                                     # :43-67
            table = sinusoid_table(torch.arange(0, int(seq.max()) + 1), cfg.dim)     # :48-53
            sc = table[seq.long(), :][:-1]                                           # :54-58
            sc = F.pad(sc[None, : n - 1, :], (0, 0, 1, 0, 0, 0), "constant", 0)      # :63-66
            x = x + sc
    if conditionings and cfg.conditioning_type != "none":                            # :248-264
        if cfg.conditioning_type == "bos_replacement":
            c = torch.zeros_like(x[:, 0, :]).unsqueeze(1)
            for i, cond in enumerate(conditionings):
                c = c + F.embedding(cond, sd[f"conditioning_emb.{i}.weight"])
            x = torch.cat((c[:, :1, :], x[:, 1:, :]), dim=1)                         # x[:, 0, :] = c[:, 0, :]
        else:
            for i, cond in enumerate(conditionings):
                x = torch.cat((F.embedding(cond, sd[f"conditioning_emb.{i}.weight"]), x), dim=1)
    pos = sd["pos_emb.emb"] if cfg.fixed_position_emb else sd["pos_emb.emb.weight"]
    x = x + pos[: x.shape[1]]                                                        # :266 (Absolute | Fixed)PositionalEmbedding
    return x


def forward(sd: Dict[str, torch.Tensor], cfg: PerformerConfig, tokens: torch.Tensor,
            spatial_seqs: Optional[Sequence[torch.Tensor]] = None, return_encodings: bool = False,
            conditionings: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
    """Performer.forward, performer.py:229-288 -> logits [B, N, num_tokens]."""
    assert tokens.shape[1] <= cfg.max_seq_len                                        # :237-239
    x = embed(sd, cfg, tokens, spatial_seqs, conditionings)
    for i in range(cfg.depth):                                                       # SequentialSequence
        p = layer_prefix(i)
        x = x + self_attention(x, sd, p + "0.fn.", cfg) * sd[p + "0.g"]              # ReZero(SelfAttention)
        x = x + feed_forward(x, sd, p + "1.fn.fn.") * sd[p + "1.g"]                  # ReZero(Chunk(FeedForward))
    x = F.layer_norm(x, (cfg.dim,), sd["norm.weight"], sd["norm.bias"], 1e-5)        # :273
    if conditionings and cfg.conditioning_type == "prepending":                      # :275-280
        x = x[:, len(conditionings):, :]
    if return_encodings:
        return x
    return F.linear(x, sd["to_out.weight"], sd["to_out.bias"])                       # :286


def ce_loss(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """inferer/transformer.py:29 (transpose to [B, V, N]) + losses/transformer/transformer.py:24-33 (mean CE)."""
    return F.cross_entropy(logits.transpose(1, 2).float(), target.long(), reduction="mean")


def train_step_grads(sd: Dict[str, torch.Tensor], cfg: PerformerConfig, x_in: torch.Tensor, y: torch.Tensor,
                     spatial_seqs: Optional[Sequence[torch.Tensor]] = None,
                     conditionings: Optional[Sequence[torch.Tensor]] = None):
    """-> (loss, {key: grad}, logits) with torch autograd over the restatement."""
    leaves = {k: (v.clone().requires_grad_(True) if k in set(trainable_keys(sd)) else v) for k, v in sd.items()}
    logits = forward(leaves, cfg, x_in, spatial_seqs, conditionings=conditionings)
    loss = ce_loss(logits, y)
    keys = [k for k in trainable_keys(sd)]
    grads = torch.autograd.grad(loss, [leaves[k] for k in keys], allow_unused=True)
    out = {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(keys, grads)}
    return loss.detach(), out, logits.detach()


def adam_step(p, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam (run_transformer.py:109), no weight decay, no amsgrad."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)) + eps
    return p - (lr / bc1) * (m / denom), m, v


# ----------------------------------------------------------------------------------------
# redraw rule
# ----------------------------------------------------------------------------------------
class ProjectionUpdaterState:
    """performer-pytorch 1.0.11 ProjectionUpdater.redraw_projections: in training only; if calls >= interval: redraw
    every layer and reset the counter, ELSE increment it (so interval = 1 redraws on every second forward)."""

    def __init__(self, interval: Optional[int]):
        self.interval = interval
        self.calls_since_last_redraw = 0

    def step(self, training: bool) -> bool:
        if not training:
            return False
        if self.interval is not None and self.calls_since_last_redraw >= self.interval:
            self.calls_since_last_redraw = 0
            return True
        self.calls_since_last_redraw += 1
        return False
