"""B200-native Performer prior: drop-in for the reference's ``Performer``
(/root/reference/src/networks/transformers/performer.py:70-288) and for the third-party stack it instantiates
(performer-pytorch==1.0.11 ``Performer`` / ``SelfAttention`` / ``FastAttention`` / ``FeedForward`` / ``ReZero`` /
``ProjectionUpdater``, local-attention ``LocalAttention``, fast-transformers ``CausalDotProduct``).

Same keyword-only constructor, same module tree (=> same ``state_dict`` keys), same methods -- but no ``nn.Linear`` /
``nn.Embedding`` forward is ever executed: the ``nn`` modules are parameter containers, the whole network runs as
one hand-scheduled forward / backward programme over the C ABI in ``include/synthanatomy_b200_performer.h``.

Precision: ``compute_dtype=None`` follows the caller like the reference (fp32 normally = CUDA-core fp32 "parity"
path; bf16 operands + fp32 accumulation on tcgen05 under ``torch.autocast``).  ``compute_dtype=ops.BF16X3`` keeps fp32
tensors and runs every dense layer (75 % of the FLOPs) on the bf16 tensor cores as hi.hi + lo.hi + hi.lo of split
operands (csrc/sa_x3.cu; the attention kernels stay on the exact fp32 CUDA-core path): the reference's own precision
class (fp32 storage, TF32-or-better products; run_transformer.py:165 amp=False).  The residual stream, LayerNorm
statistics, softmax / feature-map statistics, logits and all gradients of parameters are fp32 in every mode.
"""
from __future__ import annotations

import math
import os
import weakref
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import nn

from ... import ops, pf_ops
from ...pf_ops import SA_ACT_GELU_FWD, SA_ACT_GELU_FWD_D, SA_ACT_MUL_PRE
from .transformer import TransformerBase

# TransformerConditioningType values (src/utils/transformer.py:21-24)
_NONE, _BOS_REPLACEMENT, _PREPENDING = "none", "bos_replacement", "prepending"


def _no_exec(*_a, **_k):
    raise RuntimeError("synthanatomy_b200: parameter container, not executable (no eager fallback)")


# ------------------------------------------------------------------------------------------------
# parameter containers: the reference's module tree
# ------------------------------------------------------------------------------------------------
class AbsolutePositionalEmbedding(nn.Module):
    """performer-pytorch: emb(arange(n))"""

    def __init__(self, dim, max_seq_len):
        super().__init__()
        self.emb = nn.Embedding(max_seq_len, dim)

    forward = _no_exec


class AbsoluteSpatialPositionalEmbedding(nn.Module):
    """performer.py:23-40"""

    def __init__(self, dim: int, spatial_indices_sequence: torch.Tensor):
        super().__init__()
        self.register_buffer("spatial_indices_sequence", spatial_indices_sequence)
        self.spatial_indices_sequence = self.spatial_indices_sequence[:-1]     # the last element is the predicted one
        self.emb = nn.Embedding(len(self.spatial_indices_sequence), dim)

    forward = _no_exec


def _sinusoid_table(positions: torch.Tensor, dim: int) -> torch.Tensor:
    """[len(positions), dim] = cat(sin(p f), cos(p f)), f = 10000^(-2i / dim)"""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim))
    sinusoid_inp = torch.einsum("i,j->ij", positions.float(), inv_freq)
    return torch.cat((sinusoid_inp.sin(), sinusoid_inp.cos()), dim=-1)


class FixedPositionalEmbedding(nn.Module):
    """performer-pytorch: sinusoidal table over the sequence positions (buffer ``emb``), emb[:n]"""

    def __init__(self, dim, max_seq_len):
        super().__init__()
        self.register_buffer("emb", _sinusoid_table(torch.arange(0, max_seq_len), dim))

    forward = _no_exec


class FixedSpatialPositionalEmbedding(nn.Module):
    """performer.py:43-67: sinusoidal code of the COORDINATE of each sequence position (buffer ``emb``, one row per
    position, last one dropped: it is the predicted position)"""

    def __init__(self, dim: int, spatial_indices_sequence: torch.Tensor):
        super().__init__()
        max_position = int(torch.max(spatial_indices_sequence))
        table = _sinusoid_table(torch.arange(0, max_position + 1), dim)
        self.register_buffer("emb", table[spatial_indices_sequence.long(), :][:-1])

    forward = _no_exec


class SinusoidalEmbeddings(nn.Module):
    """local-attention rotary frequencies (buffer ``inv_freq``)"""

    def __init__(self, dim):
        super().__init__()
        inv_freq = 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer("inv_freq", inv_freq)

    forward = _no_exec


class LocalAttention(nn.Module):
    def __init__(self, window_size, dim_head, rel_pos: bool):
        super().__init__()
        self.window_size = window_size
        self.rel_pos = SinusoidalEmbeddings(dim_head) if rel_pos else None

    forward = _no_exec


def orthogonal_matrix_chunk(cols, generator=None):
    block = torch.randn((cols, cols), generator=generator)
    q, _ = torch.linalg.qr(block, mode="reduced")
    return q.t()


def gaussian_orthogonal_random_matrix(nb_rows, nb_columns, scaling=0, generator=None):
    """performer-pytorch 1.0.11 (QR on the host, as the reference does; off the GPU's critical path)."""
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
        raise ValueError(f"Invalid scaling {scaling}")
    return torch.diag(multiplier) @ final


class FastAttention(nn.Module):
    def __init__(self, dim_heads, nb_features=None, ortho_scaling=0):
        super().__init__()
        self.dim_heads = dim_heads
        self.nb_features = nb_features if nb_features is not None else int(dim_heads * math.log(dim_heads))
        self.ortho_scaling = ortho_scaling
        self.register_buffer("projection_matrix", self._draw())

    def _draw(self):
        return gaussian_orthogonal_random_matrix(self.nb_features, self.dim_heads, self.ortho_scaling)

    @torch.no_grad()
    def redraw_projection_matrix(self, device=None):
        self.projection_matrix.copy_(self._draw().to(self.projection_matrix.device, non_blocking=True))

    forward = _no_exec


class SelfAttention(nn.Module):
    def __init__(self, dim, heads, dim_head, local_heads, local_window_size, nb_features, local_rel_pos: bool):
        super().__init__()
        inner = dim_head * heads
        self.heads = heads
        self.global_heads = heads - local_heads
        if self.global_heads > 0:
            self.fast_attention = FastAttention(dim_head, nb_features)
        self.local_attn = LocalAttention(local_window_size, dim_head, local_rel_pos) if local_heads > 0 else None
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)

    forward = _no_exec


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.w1 = nn.Linear(dim, dim * mult)
        self.w2 = nn.Linear(dim * mult, dim)

    forward = _no_exec


class Chunk(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    forward = _no_exec


class ReZero(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.g = nn.Parameter(torch.tensor(1e-3))
        self.fn = fn

    forward = _no_exec


class SequentialSequence(nn.Module):
    def __init__(self, layers):
        super().__init__()
        self.layers = layers

    forward = _no_exec


class ProjectionUpdater(nn.Module):
    """performer-pytorch 1.0.11 (the shared ``instance`` sub-module duplicates the keys in state_dict there too)."""

    def __init__(self, instance, feature_redraw_interval):
        super().__init__()
        self.instance = instance
        self.feature_redraw_interval = feature_redraw_interval
        self.register_buffer("calls_since_last_redraw", torch.tensor(0))
        self._calls = 0     # host mirror of the counter: no device -> host sync on the hot path

    def fix_projections_(self):
        self.feature_redraw_interval = None

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        # resume the redraw phase where the checkpoint left it (one .item() at load time, none on the hot path)
        self._calls = int(self.calls_since_last_redraw.item())

    def __getstate__(self):      # the prefetch thread / pinned buffers are rebuilt on demand (deepcopy, pickling)
        st = dict(self.__dict__)
        for k in ("_pins", "_pin_events", "_gen", "_pool", "_next"):
            st.pop(k, None)
        return st

    def _draw_all_into(self, slot: int, mods):
        """host side of a redraw: QR draws of every layer into pinned buffer `slot` (runs on the prefetch thread)"""
        evs = self._pin_events
        if evs[slot] is not None:
            evs[slot].synchronize()            # the copies issued from this buffer two redraws ago have executed
        for i, m in enumerate(mods):
            self._pins[slot][i].copy_(gaussian_orthogonal_random_matrix(m.nb_features, m.dim_heads, m.ortho_scaling,
                                                                        generator=self._gen))
        return slot

    def _redraw_all(self):
        """Redraw every layer's projection matrix.  The QR factorisations run on the host like the reference's
        (performer-pytorch draws on the CPU and copies), but off the critical path: the matrices of the NEXT redraw are
        drawn by a background thread into pinned memory while the device works, and a redraw itself is only a batch of
        asynchronous H2D copies.  Draws come from a private generator seeded once from torch's global RNG, so a run is
        reproducible under torch.manual_seed."""
        mods = [m for m in self.instance.modules() if isinstance(m, FastAttention)]
        if not mods:
            return
        dev = mods[0].projection_matrix.device
        if dev.type != "cuda":
            for m in mods:
                m.redraw_projection_matrix()
            return
        shape = (len(mods),) + tuple(mods[0].projection_matrix.shape)
        st = self.__dict__
        if st.get("_pins") is None or tuple(st["_pins"][0].shape) != shape:
            import concurrent.futures as cf
            st["_pins"] = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(2)]
            st["_pin_events"] = [None, None]
            st["_gen"] = torch.Generator().manual_seed(int(torch.randint(0, 2 ** 62, (1,)).item()))
            st["_pool"] = cf.ThreadPoolExecutor(max_workers=1)
            st["_next"] = st["_pool"].submit(self._draw_all_into, 0, mods)
        slot = self._next.result()
        with torch.no_grad():
            for i, m in enumerate(mods):
                m.projection_matrix.copy_(self._pins[slot][i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pin_events[slot] = ev
        self._next = self._pool.submit(self._draw_all_into, 1 - slot, mods)

    def redraw_projections(self):
        if not self.training:
            return
        if self.feature_redraw_interval is not None and self._calls >= self.feature_redraw_interval:
            self._redraw_all()
            self._calls = 0
            self.calls_since_last_redraw.zero_()
            return
        self._calls += 1
        self.calls_since_last_redraw += 1

    forward = _no_exec


class PerformerStack(nn.Module):
    """container for performer_pytorch.Performer(dim, depth, heads, dim_head, ...) as built at performer.py:194-219"""

    def __init__(self, dim, depth, heads, dim_head, local_attn_heads, local_window_size, ff_mult, nb_features,
                 feature_redraw_interval, auto_check_redraw, local_rel_pos):
        super().__init__()
        layers = nn.ModuleList([])
        for _ in range(depth):
            layers.append(nn.ModuleList([
                ReZero(SelfAttention(dim, heads, dim_head, local_attn_heads, local_window_size, nb_features,
                                     local_rel_pos)),
                ReZero(Chunk(FeedForward(dim, ff_mult))),
            ]))
        self.net = SequentialSequence(layers)
        self.auto_check_redraw = auto_check_redraw
        self.proj_updater = ProjectionUpdater(self.net, feature_redraw_interval)

    def fix_projection_matrices_(self):
        self.proj_updater.feature_redraw_interval = None

    def check_redraw_projections(self):
        self.proj_updater.redraw_projections()

    forward = _no_exec


# ------------------------------------------------------------------------------------------------
# the programme
# ------------------------------------------------------------------------------------------------
class _Dims:
    def __init__(self, net: "Performer", B: int, N: int, dt: torch.dtype):
        self.B, self.N, self.M = B, N, B * N
        self.dim, self.depth = net.dim, net.depth
        self.heads, self.dh = net.heads, net.dim_head
        self.lh = net.local_attn_heads
        self.gh = net.heads - net.local_attn_heads
        self.inner = net.heads * net.dim_head
        self.ff = net.dim * net.ff_mult
        self.W = net.local_window_size
        self.m = net.nb_features
        self.mp = ((self.m + 15) // 16) * 16
        self.V = net.num_tokens
        self.dt = dt
        self.n_axes = len(net.spatial_position_emb)


_WS = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (device.index,)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


def _as(t: torch.Tensor, dt: torch.dtype) -> torch.Tensor:
    """dtype-converted contiguous copy made by the library's cast kernel (weights are tiny)."""
    t = t.contiguous()
    if t.dtype == dt:
        return t
    out = torch.empty_like(t, dtype=dt)
    ops._lib.check(ops.lib().sa_cast(ops._p(t), ops._dt(t.dtype), ops._p(out), ops._dt(dt), t.numel(), ops._stream()),
                   "sa_cast")
    return out


def _t(w: torch.Tensor) -> torch.Tensor:
    """contiguous transpose of a 2-D weight (same dtype), one launch of the library's tiled transpose kernel"""
    n, k = w.shape
    return ops.ncdhw_to_ndhwc(w.contiguous().view(1, n, k), w.dtype).view(k, n)


class _WeightPrep:
    """bf16 copies of every layer's dense weights (q | k | v concatenated) and their transposes, refreshed once per forward
    pass by ONE multi-tensor launch (`sa_weight_prep`) into buffers that live as long as the module's parameters stay where
    they are.  Before: a concat, four casts and four transposes per layer and step (218 launches at depth 24)."""

    def __init__(self, net: "Performer", D: "_Dims", dev):
        layers = net.performer.net.layers
        self.key = self._key(net)
        inner, dim, ff = D.inner, D.dim, D.ff
        per = 2 * (3 * inner * dim + inner * dim + 2 * ff * dim)
        self.buf = torch.empty((len(layers) * per,), device=dev, dtype=torch.bfloat16)
        self.fwd, self.bwd = [], []
        items = []
        off = 0

        def take(rows, cols):
            nonlocal off
            t = self.buf[off:off + rows * cols].view(rows, cols)
            off += rows * cols
            return t

        for layer in layers:
            att, ffn = layer[0].fn, layer[1].fn.fn
            Wq, Wk, Wv, Wo = att.to_q.weight, att.to_k.weight, att.to_v.weight, att.to_out.weight
            W1, W2 = ffn.w1.weight, ffn.w2.weight
            qkv, qkv_t = take(3 * inner, dim), take(dim, 3 * inner)
            o, o_t = take(dim, inner), take(inner, dim)
            w1, w1_t = take(ff, dim), take(dim, ff)
            w2, w2_t = take(dim, ff), take(ff, dim)
            for i, W in enumerate((Wq, Wk, Wv)):
                items.append((W, qkv[i * inner:], qkv_t[:, i * inner:], inner, dim, dim, 3 * inner))
            items.append((Wo, o, o_t, dim, inner, inner, dim))
            items.append((W1, w1, w1_t, ff, dim, dim, ff))
            items.append((W2, w2, w2_t, dim, ff, ff, dim))
            self.fwd.append((qkv, o, w1, w2))
            self.bwd.append((qkv_t, o_t, w1_t, w2_t))
        arr = (ops._lib.WPrepItem * len(items))()
        for a, (W, d, dt_, rows, cols, ld, ldt) in zip(arr, items):
            assert tuple(W.shape) == (rows, cols) and W.dtype == torch.float32 and W.is_contiguous()
            a.src, a.dst, a.dst_t = W.data_ptr(), d.data_ptr(), dt_.data_ptr()
            a.rows, a.cols, a.dst_ld, a.dst_t_ld = rows, cols, ld, ldt
        self.items, self.n = arr, len(items)

    @staticmethod
    def _key(net):
        return tuple(p.data_ptr() for p in net.performer.net.layers.parameters())

    def refresh(self):
        ops._lib.check(ops.lib().sa_weight_prep(self.items, self.n, ops._stream()), "sa_weight_prep")


_WPREP: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _weight_prep(net: "Performer", D: "_Dims", dev) -> Optional[_WeightPrep]:
    """the module's prepared bf16 weights for this pass, or None where the per-layer path applies (fp32 / bf16x3 arithmetic)"""
    if D.dt != torch.bfloat16:
        return None
    wp = _WPREP.get(net)
    if wp is None or wp.buf.device != dev or wp.key != _WeightPrep._key(net):
        wp = _WeightPrep(net, D, dev)
        _WPREP[net] = wp               # (not a module attribute: the ctypes table must not travel with deepcopy / pickle)
    wp.refresh()
    return wp


# (cos, sin) tables of the local heads' rotary term, read by the local-attention backward kernels' epilogues.  One entry
# per frequency buffer (every layer owns one); the entry keeps the buffer alive, so its address cannot be reused, and a
# load_state_dict / in-place write moves _version.
_ROT_TABLES = {}


def _rot_table(inv_freq: torch.Tensor, N: int, dh: int) -> torch.Tensor:
    key = (str(inv_freq.device), inv_freq.data_ptr(), inv_freq._version, N, dh)       # (addresses are per device)
    hit = _ROT_TABLES.get(key)
    if hit is None:
        if len(_ROT_TABLES) >= 256:
            _ROT_TABLES.clear()
        hit = (inv_freq, pf_ops.rotary_table(inv_freq, N, dh))
        _ROT_TABLES[key] = hit
    return hit[1]


def _FAVOR_FUSED_BWD() -> bool:
    """A/B switch (0: scan_bwd + two featmap_bwd launches): the feature-map backward as the epilogue of the dq' / dk' kernels"""
    return os.environ.get("SA_FAVOR_FUSED_BWD", "1") != "0"


class _Ctx:
    """what every piece of the programme shares for one forward / backward pass"""

    def __init__(self, net: "Performer", B: int, N: int, dt: torch.dtype, x3: bool, dev):
        self.net, self.x3, self.dev = net, x3, dev
        self.lead = 0
        self.D = D = _Dims(net, B, N, dt)
        self.wp = None if x3 else _weight_prep(net, D, dev)
        self.fd = pf_ops.favor_desc(B, N, D.gh, D.dh, D.m, D.mp, 3 * D.inner, dt) if D.gh > 0 else None
        self.ld = pf_ops.local_desc(B, N, D.lh, D.dh, D.W, 3 * D.inner, D.inner, dt) if D.lh > 0 else None

    def ws(self, backward: bool):
        return _workspace(pf_ops.favor_scan_workspace(self.fd, backward), self.dev) if self.fd is not None else None


# The bf16 copy of a gradient that the GEMM epilogue of one piece's backward writes next to the fp32 gradient it returns:
# the next piece's backward (the layer below) looks it up by the storage of the gradient autograd hands it, and casts
# only if it is not there (a hook or an accumulation in between made a new tensor).
_GRAD_COPY = {}


def _grad_copy(g: torch.Tensor, dt: torch.dtype) -> torch.Tensor:
    hit = _GRAD_COPY.pop("last", None)
    if hit is not None and hit[0] == g.data_ptr() and hit[1].dtype == dt and hit[1].shape == g.shape:
        return hit[1]
    return _as(g, dt)


class _EmbedFn(torch.autograd.Function):
    """performer.py:241-268: token + spatial + absolute position embeddings -> (x fp32, x in the activation dtype)"""

    @staticmethod
    def forward(ctx, tokens, C, *tables):
        D = C.D
        f32 = torch.float32
        p = [t.detach() for t in tables]
        tokens = tokens.long().contiguous()
        x32 = torch.empty((D.M, D.dim), device=C.dev, dtype=f32)
        xa = x32 if D.dt == f32 else torch.empty((D.M, D.dim), device=C.dev, dtype=D.dt)
        sp_idx = C.net._sp_idx(D.N - C.lead, C.dev, C.lead)
        pf_ops.embed_fwd(tokens, sp_idx, p[0], p[2:], p[1], x32, None if D.dt == f32 else xa)
        ctx.C, ctx.tokens, ctx.sp_idx, ctx.shapes = C, tokens, sp_idx, [t.shape for t in p]
        ctx.set_materialize_grads(False)      # no zero tensor for the (non-differentiable) activation-dtype copy
        if D.dt == f32:
            return x32, None              # fp32: the residual stream itself is the operand
        ctx.mark_non_differentiable(xa)
        return x32, xa

    @staticmethod
    def backward(ctx, dx32, _dxa):
        C = ctx.C
        _GRAD_COPY.pop("last", None)
        dev = dx32.device
        d_tabs = [torch.zeros(s, device=dev, dtype=torch.float32) for s in ctx.shapes]
        pf_ops.embed_bwd(dx32.contiguous(), ctx.tokens, ctx.sp_idx, d_tabs[0], d_tabs[2:], d_tabs[1])
        return (None, None, *d_tabs)


class _LayerFn(torch.autograd.Function):
    """one performer-pytorch layer: x + g_a * SelfAttention(x), then x + g_f * FeedForward(x) (ReZero, no LayerNorm).
    params: g_a, Wq, Wk, Wv, Wo, g_f, W1, b1, W2, b2.  Returning the gradients layer by layer lets DistributedDataParallel
    start reducing a layer's bucket while the layers below are still in their backward pass."""

    @staticmethod
    def forward(ctx, x32, xa, C, li, *params):
        with ops.x3_mode(C.x3):
            return _LayerFn._forward(ctx, x32, xa, C, li, *params)

    @staticmethod
    def _forward(ctx, x32, xa, C, li, *params):
        D, net, dev = C.D, C.net, C.dev
        dt, M, B, N = D.dt, D.M, D.B, D.N
        f32 = torch.float32
        is32 = dt == f32
        need_grad = any(ctx.needs_input_grad)
        ctx.set_materialize_grads(False)      # no zero tensor for the (non-differentiable) activation-dtype copy
        g_a, Wq, Wk, Wv, Wo, g_f, W1, b1, W2, b2 = [q.detach() for q in params]
        x32 = x32.detach()
        xa = x32 if (is32 or xa is None) else xa.detach()
        attn_mod = net.performer.net.layers[li][0].fn
        if C.wp is not None:                   # prepared for all layers by one launch at the top of the pass
            Wqkv, Wo_, W1_, W2_ = C.wp.fwd[li]
        else:
            Wqkv = _as(torch.cat((Wq, Wk, Wv), dim=0), dt)
            Wo_, W1_, W2_ = _as(Wo, dt), _as(W1, dt), _as(W2, dt)
        # ---- attention sub-layer
        xa_attn = xa
        qkv = torch.empty((M, 3 * D.inner), device=dev, dtype=dt)
        pf_ops.gemm_nt(xa, Wqkv, out_act=qkv)
        attn = torch.empty((M, D.inner), device=dev, dtype=dt)
        qf = kf = argq = kmax = den = lse = proj = inv_freq = states = None
        if D.gh > 0:
            fd = C.fd
            proj = attn_mod.fast_attention.projection_matrix
            kmax = torch.zeros((1,), device=dev, dtype=torch.int64)
            pf_ops.favor_kmax(fd, qkv, D.inner, proj, kmax)
            qf = torch.empty((B, D.gh, N, D.mp), device=dev, dtype=dt)
            kf = torch.empty((B, D.gh, N, D.mp), device=dev, dtype=dt)
            argq = torch.empty((B, D.gh, N), device=dev, dtype=torch.int32)
            pf_ops.favor_featmap_fwd(fd, qkv, 0, proj, True, None, net.eps_feature, qf, argq)
            pf_ops.favor_featmap_fwd(fd, qkv, D.inner, proj, False, kmax, net.eps_feature, kf, None)
            den = torch.empty((B, D.gh, N), device=dev, dtype=f32)
            # the tcgen05 path can keep its per-chunk prefix states for the backward scan (no recomputation)
            nst = pf_ops.favor_scan_states_bytes(fd) if need_grad else 0
            states = torch.empty((nst,), device=dev, dtype=torch.uint8) if nst else None
            pf_ops.favor_scan_fwd(fd, qf, kf, qkv, 2 * D.inner, net.eps_cumsum, attn, 0, den, C.ws(need_grad), states)
        if D.lh > 0:
            inv_freq = attn_mod.local_attn.rel_pos.inv_freq if attn_mod.local_attn.rel_pos is not None else None
            lse = torch.empty((B, D.lh, N), device=dev, dtype=f32)
            c0 = D.gh * D.dh
            if inv_freq is not None:       # rotary position term, in place: the saved q / k of the local heads are rotated
                pf_ops.rotary_qk(qkv, c0, D.inner + c0, B, N, D.lh, D.dh, inv_freq, False)
            pf_ops.local_attn_fwd(C.ld, qkv, c0, D.inner + c0, 2 * D.inner + c0, None, attn, c0, lse)
        x_mid = torch.empty((M, D.dim), device=dev, dtype=f32)
        xa_mid = x_mid if is32 else torch.empty((M, D.dim), device=dev, dtype=dt)
        pf_ops.gemm_nt(attn, Wo_, scale_dev=g_a, resid=x32, out_f32=x_mid, out_act=None if is32 else xa_mid)
        # ---- feed-forward sub-layer
        u = torch.empty((M, D.ff), device=dev, dtype=dt)
        h = torch.empty((M, D.ff), device=dev, dtype=dt)
        # `u` receives gelu'(x W1^T + b1) when a backward pass will follow (it needs nothing else of the pre-activation)
        pf_ops.gemm_nt(xa_mid, W1_, bias=b1, act=SA_ACT_GELU_FWD_D if need_grad else SA_ACT_GELU_FWD, pre=u, out_act=h)
        x_out = torch.empty((M, D.dim), device=dev, dtype=f32)
        xa_out = None if is32 else torch.empty((M, D.dim), device=dev, dtype=dt)
        pf_ops.gemm_nt(h, W2_, bias=b2, scale_dev=g_f, resid=x_mid, out_f32=x_out, out_act=xa_out)
        if need_grad:
            ctx.C = C
            ctx.saved = (xa_attn, qkv, qf, kf, argq, kmax, den, lse, attn, xa_mid, u, h, proj, inv_freq, states)
            ctx.weights = (Wqkv, Wo_, W1_, W2_)
            ctx.weights_t = C.wp.bwd[li] if C.wp is not None else None
            ctx.scalars = (g_a, g_f, b1, b2)
            ctx.masters = (Wo, W2)
        if is32:
            return x_out, None
        ctx.mark_non_differentiable(xa_out)
        return x_out, xa_out

    @staticmethod
    def backward(ctx, g32, _gxa):
        with ops.x3_mode(ctx.C.x3):
            return _LayerFn._backward(ctx, g32)

    @staticmethod
    def _backward(ctx, g32):
        C = ctx.C
        D, net, dev = C.D, C.net, C.dev
        dt, M, B, N = D.dt, D.M, D.B, D.N
        f32 = torch.float32
        is32 = dt == f32
        if ctx.saved is None:
            raise RuntimeError("Performer: second backward through a layer whose activations were released")
        xa_attn, qkv, qf, kf, argq, kmax, den, lse, attn, xa_ffn, u, h, proj, inv_freq, states = ctx.saved
        Wqkv, Wo_, W1_, W2_ = ctx.weights
        if ctx.weights_t is not None:
            Wqkv_t, Wo_t, W1_t, W2_t = ctx.weights_t
        else:
            Wqkv_t, Wo_t, W1_t, W2_t = _t(Wqkv), _t(Wo_), _t(W1_), _t(W2_)
        g_a, g_f, b1, b2 = ctx.scalars
        Wo, W2 = ctx.masters
        ctx.saved = ctx.weights = ctx.masters = None
        g32 = g32.contiguous()
        dxa = g32 if is32 else _grad_copy(g32, dt)
        # ---- feed-forward sub-layer: x_out = x + g_f * (gelu(x W1^T + b1) W2^T + b2)
        # (the gate gradient sum (dx W2) . h is read off the unscaled weight gradient dx^T h below: the data-gradient
        #  GEMM's epilogue does not have to stream h again)
        dot = torch.zeros((1,), device=dev, dtype=f32)
        du = torch.empty((M, D.ff), device=dev, dtype=dt)
        pf_ops.gemm_nt(dxa, W2_t, scale_dev=g_f, act=SA_ACT_MUL_PRE, pre=u, out_act=du)
        colsum = torch.empty((D.dim,), device=dev, dtype=f32)
        dW2 = torch.empty((D.dim, D.ff), device=dev, dtype=f32)
        pf_ops.gemm_tn(dxa, h, dW2, colsum=colsum)         # + the column sums of dx (bias gradient) from the same tiles
        pf_ops.gate_wgrad(dW2, W2, g_f, dot)               # dot = sum (dx^T h) . W2;  dW2 *= g_f
        db2 = torch.empty_like(b2)
        dg_f = torch.empty((), device=dev, dtype=f32)
        pf_ops.rezero_finish(colsum, b2, g_f, dot, db2, dg_f)
        dW1 = torch.empty((D.ff, D.dim), device=dev, dtype=f32)
        db1 = torch.empty((D.ff,), device=dev, dtype=f32)
        pf_ops.gemm_tn(du, xa_ffn, dW1, colsum=db1)
        d_mid = torch.empty((M, D.dim), device=dev, dtype=f32)
        dxa_mid = d_mid if is32 else torch.empty((M, D.dim), device=dev, dtype=dt)
        pf_ops.gemm_nt(du, W1_t, resid=g32, out_f32=d_mid, out_act=None if is32 else dxa_mid)
        del du, u, h
        # ---- attention sub-layer: x_mid = x + g_a * (attn Wo^T)
        dot = torch.zeros((1,), device=dev, dtype=f32)
        dattn = torch.empty((M, D.inner), device=dev, dtype=dt)
        pf_ops.gemm_nt(dxa_mid, Wo_t, scale_dev=g_a, out_act=dattn)
        dWo = torch.empty((D.dim, D.inner), device=dev, dtype=f32)
        pf_ops.gemm_tn(dxa_mid, attn, dWo)
        pf_ops.gate_wgrad(dWo, Wo, g_a, dot)               # dg_a = sum (dx^T attn) . Wo;  dWo *= g_a
        dqkv = torch.empty((M, 3 * D.inner), device=dev, dtype=dt)
        if D.lh > 0:
            c0 = D.gh * D.dh
            if inv_freq is not None and os.environ.get("SA_LOCAL_ROT_FUSED", "1") == "0":      # A/B: the two-pass form
                pf_ops.local_attn_bwd(C.ld, qkv, c0, D.inner + c0, 2 * D.inner + c0, None, attn, dattn, c0, lse, dqkv)
                pf_ops.rotary_qk(dqkv, c0, D.inner + c0, B, N, D.lh, D.dh, inv_freq, True)
            elif inv_freq is not None:     # the gradients of the rotated q / k leave through the transpose of the rotary map
                pf_ops.local_attn_bwd_rot(C.ld, qkv, c0, D.inner + c0, 2 * D.inner + c0, inv_freq,
                                          _rot_table(inv_freq, N, D.dh), attn, dattn, c0, lse, dqkv)
            else:
                pf_ops.local_attn_bwd(C.ld, qkv, c0, D.inner + c0, 2 * D.inner + c0, None, attn, dattn, c0, lse, dqkv)
        if D.gh > 0:
            fd = C.fd
            if states is not None and pf_ops.favor_scan_states_bytes(fd) == 0:
                states = None          # the dispatch changed since the forward pass: recompute
            gsum = torch.zeros((1,), device=dev, dtype=f32)
            if _FAVOR_FUSED_BWD() and pf_ops.favor_scan_bwd_fused_supported(fd):
                # dq' / dk' are turned into dq / dk block by block inside the kernels that produce them
                pf_ops.favor_scan_bwd_fused(fd, qf, kf, qkv, 0, D.inner, 2 * D.inner, proj, net.eps_cumsum, net.eps_feature,
                                            attn, dattn, 0, den, argq, dqkv, gsum, C.ws(True), states)
            else:
                dqf = torch.empty_like(qf)
                dkf = torch.empty_like(kf)
                pf_ops.favor_scan_bwd(fd, qf, kf, qkv, 2 * D.inner, net.eps_cumsum, attn, dattn, 0, den, dqf, dkf, dqkv,
                                      2 * D.inner, C.ws(True), states)
                pf_ops.favor_featmap_bwd(fd, qkv, 0, proj, True, net.eps_feature, qf, dqf, argq, dqkv, 0, None)
                pf_ops.favor_featmap_bwd(fd, qkv, D.inner, proj, False, net.eps_feature, kf, dkf, None, dqkv, D.inner, gsum)
                del dqf, dkf
            pf_ops.favor_kmax_fixup(fd, proj, kmax, gsum, dqkv, D.inner)
        dWqkv = torch.empty((3 * D.inner, D.dim), device=dev, dtype=f32)
        pf_ops.gemm_tn(dqkv, xa_attn, dWqkv)
        dx = torch.empty((M, D.dim), device=dev, dtype=f32)
        dxa_in = None if is32 else torch.empty((M, D.dim), device=dev, dtype=dt)
        pf_ops.gemm_nt(dqkv, Wqkv_t, resid=d_mid, out_f32=dx, out_act=dxa_in)
        if dxa_in is not None:
            _GRAD_COPY["last"] = (dx.data_ptr(), dxa_in)
        return (dx, None, None, None, dot.view(()), dWqkv[:D.inner], dWqkv[D.inner:2 * D.inner], dWqkv[2 * D.inner:], dWo,
                dg_f, dW1, db1, dW2, db2)


class _HeadFn(torch.autograd.Function):
    """performer.py:273-286: final LayerNorm (+ to_out).  params: norm.weight, norm.bias, to_out.weight, to_out.bias"""

    @staticmethod
    def forward(ctx, x32, C, return_encodings, *params):
        with ops.x3_mode(C.x3):
            return _HeadFn._forward(ctx, x32, C, return_encodings, *params)

    @staticmethod
    def _forward(ctx, x32, C, return_encodings, *params):
        D, dev = C.D, C.dev
        dt, M, B, N = D.dt, D.M, D.B, D.N
        f32 = torch.float32
        is32 = dt == f32
        nw, nb, Wout, bout = [q.detach() for q in params]
        x32 = x32.detach()
        mean = torch.empty((M,), device=dev, dtype=f32)
        rstd = torch.empty((M,), device=dev, dtype=f32)
        enc32 = torch.empty((M, D.dim), device=dev, dtype=f32) if (return_encodings or is32) else None
        xn = enc32 if is32 else torch.empty((M, D.dim), device=dev, dtype=dt)
        pf_ops.layernorm_fwd(x32, nw, nb, 1e-5, enc32, None if is32 else xn, mean, rstd)
        if return_encodings:
            out = enc32.view(B, N, D.dim)
        else:
            logits = torch.empty((M, D.V), device=dev, dtype=f32)
            pf_ops.gemm_nt(xn, _as(Wout, dt), bias=bout, out_f32=logits)
            out = logits.view(B, N, D.V)
        if any(ctx.needs_input_grad):
            ctx.C, ctx.return_encodings = C, return_encodings
            ctx.saved = (x32, xn, mean, rstd, nw, nb, Wout)
        return out

    @staticmethod
    def backward(ctx, gout):
        with ops.x3_mode(ctx.C.x3):
            return _HeadFn._backward(ctx, gout)

    @staticmethod
    def _backward(ctx, gout):
        C = ctx.C
        D, dev = C.D, C.dev
        dt, M = D.dt, D.M
        f32 = torch.float32
        is32 = dt == f32
        x32, xn, mean, rstd, nw, nb, Wout = ctx.saved
        ctx.saved = None
        dWout = dbout = None
        if ctx.return_encodings:
            dxn = gout.float().contiguous().view(M, D.dim)
        else:
            dl = gout.float().contiguous().view(M, D.V)
            dbout = ops.bias_grad(dl)
            if is32:
                dlb, Wout_t = dl, Wout.t().contiguous()
            else:
                Vp = ((D.V + 63) // 64) * 64
                dlb = torch.empty((M, Vp), device=dev, dtype=dt)
                pf_ops.cast2d(dl, dlb, D.V)
                Wout_t = torch.empty((D.dim, Vp), device=dev, dtype=dt)
                pf_ops.cast2d(Wout.t().contiguous(), Wout_t, D.V)
            dxn = torch.empty((M, D.dim), device=dev, dtype=f32)
            pf_ops.gemm_nt(dlb, Wout_t, out_f32=dxn)
            dWout = torch.empty((D.V, D.dim), device=dev, dtype=f32)
            pf_ops.gemm_tn(dlb[:, :D.V], xn, dWout)
            del dl, dlb
        dx32 = torch.empty((M, D.dim), device=dev, dtype=f32)
        dnw = torch.zeros_like(nw)
        dnb = torch.zeros_like(nb)
        pf_ops.layernorm_bwd(dxn, x32, nw, mean, rstd, dx32, dnw, dnb)
        _GRAD_COPY.pop("last", None)
        return (dx32, None, None, dnw, dnb, dWout, dbout)


def _run_programme(net: "Performer", tokens: torch.Tensor, dt, return_encodings: bool, tok_table=None, lead: int = 0):
    """embeddings -> depth x layer -> LayerNorm / logits, one autograd node per piece.  tok_table: the token table with
    conditioning rows appended (None: the plain one); lead: number of prepended conditioning positions"""
    if not tokens.is_cuda:
        raise RuntimeError("synthanatomy_b200: CUDA tensors only -- there is no CPU fallback")
    ops.sync_deterministic()             # torch.backends.cudnn.deterministic -> ordered sums in the kernels
    dt, x3 = ops.resolve_dtype(dt)       # BF16X3: fp32 tensors, dense layers as split-bf16 tensor-core products
    B, N = tokens.shape
    C = _Ctx(net, B, N, dt, x3, tokens.device)
    C.lead = lead
    ps = net._params()
    if tok_table is not None:
        ps[0] = tok_table
    n_tab = 2 + C.D.n_axes
    x32, xa = _EmbedFn.apply(tokens, C, *ps[:n_tab])
    for li in range(C.D.depth):
        x32, xa = _LayerFn.apply(x32, xa, C, li, *ps[n_tab + 10 * li: n_tab + 10 * li + 10])
    return _HeadFn.apply(x32, C, return_encodings, *ps[n_tab + 10 * C.D.depth:])


class _Decoder:
    """Recurrent-state evaluation of the network, one position per call.

    ``step(tokens_t, t)`` returns the logits the reference obtains from a forward over the prefix ``x[:, :t + 1]`` at
    its last position (src/networks/transformers/transformer.py:20-56) -- including the prefix-dependent key stabiliser
    of the FAVOR+ heads -- at O(1) attention work per token: the global heads carry S = sum k' (x) v in a split form that
    can be re-normalised exactly when the stabiliser moves, the local heads keep a key / value cache.  Eval mode only
    (no projection redraw, no dropout)."""

    def __init__(self, net: "Performer", batch: int, max_len: int, dt: torch.dtype):
        dev = net.token_emb.weight.device
        self.net, self.B, self.max_len, self.dt = net, batch, max_len, dt
        D = self.D = _Dims(net, batch, 1, dt)
        f32 = torch.float32
        self.sp_idx = net._sp_idx(max_len, dev)            # [n_axes, max_len] (coordinate of position n - 1, -1 at BOS)
        self.layers = []
        for li in range(D.depth):
            att = net.performer.net.layers[li][0]
            ff = net.performer.net.layers[li][1]
            a = att.fn
            st = dict(
                g_a=att.g.detach(), g_f=ff.g.detach(),
                Wqkv=_as(torch.cat((a.to_q.weight, a.to_k.weight, a.to_v.weight), dim=0).detach(), dt),
                Wo=_as(a.to_out.weight.detach(), dt), W1=_as(ff.fn.fn.w1.weight.detach(), dt), b1=ff.fn.fn.w1.bias.detach(),
                W2=_as(ff.fn.fn.w2.weight.detach(), dt), b2=ff.fn.fn.w2.bias.detach())
            if D.gh > 0:
                st["proj"] = a.fast_attention.projection_matrix
                mh = torch.zeros((max_len + 2,), device=dev, dtype=torch.int32)
                mh[0] = 0x007FFFFF                           # ordered-uint encoding of -inf
                st["mhist"] = mh
                st["Se"] = torch.zeros((batch * D.gh, D.m, D.dh), device=dev, dtype=f32)
                st["ze"] = torch.zeros((batch * D.gh, D.m), device=dev, dtype=f32)
                st["S1"] = torch.zeros((batch * D.gh, D.dh), device=dev, dtype=f32)
            if D.lh > 0:
                st["inv_freq"] = a.local_attn.rel_pos.inv_freq if a.local_attn.rel_pos is not None else None
                st["kc"] = torch.zeros((batch, max_len, D.lh * D.dh), device=dev, dtype=dt)
                st["vc"] = torch.zeros((batch, max_len, D.lh * D.dh), device=dev, dtype=dt)
            self.layers.append(st)
        self.scratch = torch.empty((2 * batch * max(D.gh, 1) * D.m + batch * max(D.gh, 1),), device=dev, dtype=f32)
        self.Wout = _as(net.to_out.weight.detach(), dt)
        # static buffers of one step (the step is captured into a CUDA graph after a few eager calls and replayed:
        # ~200 launches per position would otherwise be bound by the host)
        is32 = dt == f32
        self.tok = torch.zeros((batch,), device=dev, dtype=torch.int64)
        self.t_dev = torch.zeros((1,), device=dev, dtype=torch.int32)
        self.x32 = [torch.empty((batch, D.dim), device=dev, dtype=f32) for _ in range(2)]
        self.xa = None if is32 else torch.empty((batch, D.dim), device=dev, dtype=dt)
        self.qkv = torch.empty((batch, 3 * D.inner), device=dev, dtype=dt)
        self.attn = torch.empty((batch, D.inner), device=dev, dtype=dt)
        self.u = torch.empty((batch, D.ff), device=dev, dtype=dt)
        self.h = torch.empty((batch, D.ff), device=dev, dtype=dt)
        self.mean = torch.empty((batch,), device=dev, dtype=f32)
        self.rstd = torch.empty((batch,), device=dev, dtype=f32)
        self.xn = torch.empty((batch, D.dim), device=dev, dtype=f32 if is32 else dt)
        self.logits = torch.empty((batch, D.V), device=dev, dtype=f32)
        self.tok_w = net.token_emb.weight.detach()
        self.sp_ws = [(m.emb if isinstance(m, FixedSpatialPositionalEmbedding) else m.emb.weight).detach()
                      for m in net.spatial_position_emb]
        self.pos_w = net._pos_table().detach()
        self.graph = None
        self.calls = 0
        import os
        self.use_graph = os.environ.get("SA_DECODE_GRAPH", "1") != "0"

    def _step_impl(self) -> None:
        """one position: everything reads the position from self.t_dev and the tokens from self.tok"""
        net, D, B, dt = self.net, self.D, self.B, self.dt
        is32 = dt == torch.float32
        cur = 0
        x32 = self.x32[cur]
        pf_ops.embed_step(self.tok, self.sp_idx, self.tok_w, self.sp_ws, self.pos_w, 0, self.t_dev, x32,
                          None if is32 else self.xa)
        xa = x32 if is32 else self.xa
        for st in self.layers:
            pf_ops.gemm_nt(xa, st["Wqkv"], out_act=self.qkv)
            if D.gh > 0:
                pf_ops.favor_decode_step(B, D.gh, D.m, 0, self.t_dev, self.qkv, 0, D.inner, 2 * D.inner, st["proj"],
                                         net.eps_feature, net.eps_cumsum, st["mhist"], self.scratch, st["Se"], st["ze"],
                                         st["S1"], self.attn, 0)
            if D.lh > 0:
                c0 = D.gh * D.dh
                pf_ops.local_decode_step(B, D.lh, D.W, 0, self.t_dev, self.max_len, self.qkv, c0, D.inner + c0,
                                         2 * D.inner + c0, st["inv_freq"], st["kc"], st["vc"], self.attn, c0)
            if is32:
                nxt = self.x32[cur ^ 1]
                pf_ops.gemm_nt(self.attn, st["Wo"], scale_dev=st["g_a"], resid=x32, out_f32=nxt)
                cur ^= 1
                x32 = xa = nxt
            else:
                pf_ops.gemm_nt(self.attn, st["Wo"], scale_dev=st["g_a"], resid=x32, out_f32=x32, out_act=self.xa)
            pf_ops.gemm_nt(xa, st["W1"], bias=st["b1"], act=SA_ACT_GELU_FWD, pre=self.u, out_act=self.h)
            if is32:
                nxt = self.x32[cur ^ 1]
                pf_ops.gemm_nt(self.h, st["W2"], bias=st["b2"], scale_dev=st["g_f"], resid=x32, out_f32=nxt)
                cur ^= 1
                x32 = xa = nxt
            else:
                pf_ops.gemm_nt(self.h, st["W2"], bias=st["b2"], scale_dev=st["g_f"], resid=x32, out_f32=x32, out_act=self.xa)
        pf_ops.layernorm_fwd(x32, net.norm.weight.detach(), net.norm.bias.detach(), 1e-5, self.xn if is32 else None,
                             None if is32 else self.xn, self.mean, self.rstd)
        pf_ops.gemm_nt(self.xn, self.Wout, bias=net.to_out.bias.detach(), out_f32=self.logits)

    @torch.no_grad()
    def step(self, tokens_t: torch.Tensor, t: int) -> torch.Tensor:
        """tokens_t [B] (the token at position t) -> logits [B, num_tokens] of position t.  Positions must be fed in
        order 0, 1, 2, ...; the returned tensor is a static buffer that the next call overwrites."""
        assert 0 <= t < self.max_len
        self.tok.copy_(tokens_t.view(-1))
        self.t_dev.fill_(t)
        if self.graph is not None:
            self.graph.replay()
            return self.logits
        self._step_impl()
        self.calls += 1
        if self.use_graph and self.calls == 3:       # kernels / allocations are warm: capture one step for replay
            try:
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(g):
                    self._step_impl()
                self.graph = g
            except Exception:                        # capture is an optimisation only
                self.graph = None
                self.use_graph = False
                torch.cuda.synchronize()
        return self.logits


class Performer(TransformerBase):
    """NOTE: All tensor logic assumes the following ordering [Batch, Length, Channel] (as the reference)."""

    def __init__(
        self,
        *,
        num_tokens: int,
        max_seq_len: int,
        dim: int,
        depth: int,
        heads: int,
        ordering,
        dim_head: int = 64,
        local_attn_heads: int = 0,
        local_window_size: int = 256,
        causal: bool = True,
        ff_mult: int = 4,
        nb_features: Optional[int] = None,
        feature_redraw_interval: int = 1000,
        reversible: bool = False,
        ff_chunks: int = 1,
        ff_glu: bool = False,
        emb_dropout: float = 0.0,
        ff_dropout: float = 0.0,
        attn_dropout: float = 0.0,
        generalized_attention: bool = False,
        kernel_fn: torch.nn.Module = nn.ReLU(),
        use_scalenorm: bool = False,
        use_rezero: bool = False,
        cross_attend: bool = False,
        no_projection: bool = False,
        tie_embed: bool = False,
        rotary_position_emb: bool = False,
        fixed_position_emb: bool = False,
        axial_position_emb: bool = False,
        axial_position_shape: Tuple[int, int] = None,
        auto_check_redraw: bool = True,
        qkv_bias: bool = False,
        attn_out_bias: bool = False,
        spatial_position_emb: str = None,
        spatial_shape: Union[Tuple[int, int], Tuple[int, int, int]] = None,
        conditioning_num_tokens: Optional[Tuple[int, ...]] = None,
        conditioning_type: str = _NONE,
        # ---- extensions (not in the reference signature)
        compute_dtype: Union[torch.dtype, str, None] = None,
        local_rel_pos: str = "rotary",
    ):
        super().__init__()
        unsupported = {
            "causal=False": not causal, "reversible": reversible, "ff_chunks != 1": ff_chunks != 1, "ff_glu": ff_glu,
            "dropout > 0": (emb_dropout, ff_dropout, attn_dropout) != (0.0, 0.0, 0.0),
            "generalized_attention": generalized_attention, "use_scalenorm": use_scalenorm,
            "use_rezero=False": not use_rezero, "cross_attend": cross_attend, "no_projection": no_projection,
            "tie_embed": tie_embed, "rotary / axial position_emb": rotary_position_emb or axial_position_emb,
            "qkv_bias / attn_out_bias": qkv_bias or attn_out_bias,
        }
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(f"synthanatomy_b200 Performer implements the README.md:124-141 configuration; "
                                      f"not implemented: {bad} (no fallback)")
        if isinstance(local_attn_heads, (tuple, list)):
            assert len(local_attn_heads) == 1, "per-layer local head counts are not implemented"
            local_attn_heads = local_attn_heads[0]
        assert local_rel_pos in ("rotary", "none")
        assert conditioning_type in (_NONE, _BOS_REPLACEMENT, _PREPENDING), conditioning_type
        # accounting for the number of prepended conditionings (performer.py:119-125)
        max_seq_len = max_seq_len + (len(conditioning_num_tokens)
                                     if conditioning_num_tokens and conditioning_type == _PREPENDING else 0)
        self.num_tokens, self.max_seq_len = num_tokens, max_seq_len
        self.dim, self.depth, self.heads, self.dim_head = dim, depth, heads, dim_head
        self.local_attn_heads, self.local_window_size, self.ff_mult = local_attn_heads, local_window_size, ff_mult
        self.nb_features = nb_features if nb_features is not None else int(dim_head * math.log(dim_head))
        self.eps_feature, self.eps_cumsum = 1e-4, 1e-6
        self.compute_dtype = compute_dtype
        self.conditioning_type = conditioning_type

        self.token_emb = nn.Embedding(num_tokens, dim)
        self.pos_emb = (FixedPositionalEmbedding(dim, self.max_seq_len) if fixed_position_emb
                        else AbsolutePositionalEmbedding(dim, self.max_seq_len))
        self.ordering = ordering
        self.spatial_position_emb = nn.ModuleList()
        if spatial_position_emb:
            assert spatial_position_emb in ["fixed", "absolute"], \
                f"spatial_position_emb must be either 'fixed' or  'absolute', but got {spatial_position_emb}"
            axis = (0, 1, 2) if len(spatial_shape) == 3 else (0, 1)
            coord_channels = np.array(np.meshgrid(*tuple(np.arange(0, s) for s in spatial_shape), indexing="ij"))
            coord_channels = coord_channels[[s for s in axis]]
            for i in axis:
                seq = torch.from_numpy(coord_channels[i, ...].flatten())
                seq = self.ordering(seq)
                cls = FixedSpatialPositionalEmbedding if spatial_position_emb == "fixed" else AbsoluteSpatialPositionalEmbedding
                self.spatial_position_emb.append(cls(dim=dim, spatial_indices_sequence=seq))
        self.conditioning_emb = nn.ModuleList()
        if conditioning_num_tokens:
            for cnt in conditioning_num_tokens:
                self.conditioning_emb.append(nn.Embedding(cnt, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.performer = PerformerStack(dim, depth, heads, dim_head, local_attn_heads, local_window_size, ff_mult,
                                        nb_features, feature_redraw_interval, auto_check_redraw,
                                        local_rel_pos == "rotary")
        self.norm = nn.LayerNorm(dim)
        self.to_out = nn.Linear(dim, num_tokens)
        self._sp_cache = {}

    # ---- reference API
    def check_redraw_projections(self):
        self.performer.check_redraw_projections()

    def fix_projection_matrices_(self):
        self.performer.fix_projection_matrices_()

    def _sp_idx(self, n: int, device, lead: int = 0) -> Optional[torch.Tensor]:
        """[n_axes, lead + n] int32 row of each axis' table that sequence position j adds (-1: none): nothing at the
        `lead` prepended conditioning positions and at the BOS position; then, for an absolute table (one row per
        coordinate VALUE, performer.py:27-33) the coordinate of position j - 1, for a fixed table (one row per sequence
        position, performer.py:43-57) j - 1 itself."""
        if len(self.spatial_position_emb) == 0:
            return None
        key = (n, lead, str(device))
        if key not in self._sp_cache:
            rows = []
            for mod in self.spatial_position_emb:
                if isinstance(mod, FixedSpatialPositionalEmbedding):
                    seq = torch.arange(0, n - 1, dtype=torch.int32)
                else:
                    seq = mod.spatial_indices_sequence[: n - 1].to(torch.int32).cpu()
                rows.append(torch.cat((torch.full((lead + 1,), -1, dtype=torch.int32), seq)))
            self._sp_cache = {key: torch.stack(rows).contiguous().to(device)}
        return self._sp_cache[key]

    def _pos_table(self) -> torch.Tensor:
        return self.pos_emb.emb if isinstance(self.pos_emb, FixedPositionalEmbedding) else self.pos_emb.emb.weight

    def _params(self) -> List[torch.Tensor]:
        ps = [self.token_emb.weight, self._pos_table()]
        ps += [m.emb if isinstance(m, FixedSpatialPositionalEmbedding) else m.emb.weight for m in self.spatial_position_emb]
        for layer in self.performer.net.layers:
            a, f = layer[0], layer[1]
            ps += [a.g, a.fn.to_q.weight, a.fn.to_k.weight, a.fn.to_v.weight, a.fn.to_out.weight,
                   f.g, f.fn.fn.w1.weight, f.fn.fn.w1.bias, f.fn.fn.w2.weight, f.fn.fn.w2.bias]
        ps += [self.norm.weight, self.norm.bias, self.to_out.weight, self.to_out.bias]
        return ps

    def _dtype(self) -> torch.dtype:
        if self.compute_dtype is not None:
            return self.compute_dtype
        return torch.bfloat16 if torch.is_autocast_enabled() else torch.float32

    def make_decoder(self, batch: int, max_len: int) -> _Decoder:
        """recurrent-state evaluator for sampling: ``decoder.step(tokens_t, t) -> logits [batch, num_tokens]``"""
        return _Decoder(self, batch, max_len, ops.resolve_dtype(self._dtype())[0])

    @torch.no_grad()
    def sample(self, prefix: torch.Tensor, conditioning: torch.Tensor = None, temperature: float = 1.0,
               sample: bool = True, top_k: Optional[int] = None, recurrent: Optional[bool] = None) -> torch.Tensor:
        """``TransformerBase.sample`` (reference transformer.py:58-101), same arguments, evaluated with recurrent
        attention state: one position per step instead of one forward over the whole prefix per step (O(N) instead of
        O(N^2) layer evaluations).  With one attention layer the result is the reference's; with deeper stacks it
        agrees up to the reference's own non-causal coupling through the prefix-dependent key stabiliser (see
        ``_Decoder`` and tests/test_gpu_performer.py::test_recurrent_decoder_matches_prefix_forward).
        The recurrent evaluator is OPT-IN (``recurrent=True`` or ``self.recurrent_sampling = True``): the default, like
        the reference, is the loop of full forwards over the growing prefix, so that a drop-in samples from exactly the
        reference's distribution (also used for conditioning / a CPU module)."""
        if recurrent is None:
            recurrent = getattr(self, "recurrent_sampling", False)
        if not recurrent or conditioning is not None or not prefix.is_cuda:
            return super().sample(prefix, conditioning=conditioning, temperature=temperature, sample=sample, top_k=top_k)
        self.eval()
        steps = int(np.prod(self.ordering.dimensions))
        B, P = prefix.shape
        dec = self.make_decoder(B, max(1, P + steps - 1))     # the last sampled token is never fed back
        logits = None
        for t in range(P):
            logits = dec.step(prefix[:, t], t)
        out = []
        for k in range(steps):
            lg = logits / temperature
            if top_k is not None:
                lg = self._top_k_logits(lg, top_k)
            probs = torch.softmax(lg, dim=-1)
            if sample:
                ix = torch.multinomial(probs, num_samples=1)
            else:
                _, ix = torch.topk(probs, k=1, dim=-1)
            out.append(ix)
            if k + 1 < steps:
                logits = dec.step(ix[:, 0], P + k)
        x = torch.cat(out, dim=1)
        x = x[:, self.ordering.get_revert_sequence_ordering()]
        x = x.reshape(x.shape[0], *self.ordering.dimensions)
        return torch.squeeze(x, 1)

    def forward(self, x: torch.Tensor, conditionings: Sequence[torch.Tensor] = None, return_encodings: bool = False,
                **kwargs):
        b, n = x.shape
        assert n <= self.max_seq_len, \
            f"sequence length {n} must be less than the max sequence length {self.max_seq_len}"
        if self.performer.auto_check_redraw:
            self.performer.proj_updater.redraw_projections()
        tok_table, lead = None, 0
        if conditionings and self.conditioning_type != _NONE:
            # performer.py:248-264.  Both forms become rows appended to the token table, so that the fused embedding
            # kernel (and its scatter backward) serve them unchanged; torch only concatenates / indexes [B, dim]-sized
            # pieces here and its autograd routes their gradients to the conditioning tables.
            V = self.token_emb.weight.shape[0]
            if self.conditioning_type == _BOS_REPLACEMENT:
                # x[:, 0] = sum_i conditioning_emb_i(c_i)   (replaces the BOS token embedding; position term still added)
                c = sum(emb(conditionings[i].long().view(b)) for i, emb in enumerate(self.conditioning_emb))
                tok_table = torch.cat((self.token_emb.weight, c), dim=0)
                x = torch.cat((V + torch.arange(b, device=x.device, dtype=x.dtype).view(b, 1), x[:, 1:]), dim=1)
            else:
                # x = cat(conditioning_emb_i(c_i), x) for i = 0, 1, ...: the LAST conditioning ends up first
                tok_table = torch.cat([self.token_emb.weight] + [emb.weight for emb in self.conditioning_emb], dim=0)
                offs, cols = V, []
                for i, emb in enumerate(self.conditioning_emb):
                    cols.append(offs + conditionings[i].long().view(b, 1))
                    offs += emb.weight.shape[0]
                x = torch.cat(cols[::-1] + [x], dim=1)
                lead = len(cols)
        out = _run_programme(self, x, self._dtype(), return_encodings, tok_table, lead)
        if lead:
            out = out[:, lead:, :]                    # performer.py:275-280
        return out
