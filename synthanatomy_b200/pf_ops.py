"""Tensor-level wrappers over the Performer C ABI (include/synthanatomy_b200_performer.h).

PyTorch is plumbing only (device memory, current stream); every number is computed inside
libsynthanatomy_b200.so.  CUDA tensors only -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from . import ops as _ops
from ._lib import (FavorDesc, GemmEpilogue, LocalDesc, SA_ACT_GELU_BWD, SA_ACT_GELU_FWD, SA_ACT_GELU_FWD_D,  # noqa: F401
                   SA_ACT_MUL_PRE, SA_ACT_NONE)
from .ops import _dt, _p, _stream, lib


def _ptr(t: Optional[torch.Tensor], offset_elems: int = 0):
    """raw device pointer (+ element offset); the tensor need not be contiguous as a whole (column blocks)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("synthanatomy_b200: CUDA tensors only -- there is no CPU fallback")
    return C.c_void_p(t.data_ptr() + offset_elems * t.element_size())


def _rowmajor(t: torch.Tensor) -> int:
    """leading dimension of a 2-D row-major (possibly column-sliced) tensor"""
    assert t.dim() == 2 and t.stride(1) == 1, (t.shape, t.stride())
    return t.stride(0)


# ------------------------------------------------------------------------------------------------
# dense layers
# ------------------------------------------------------------------------------------------------
def gemm_nt(a: torch.Tensor, b: torch.Tensor, *, bias=None, dot_with=None, dot_out=None, scale_dev=None,
            scale: float = 1.0, act: int = SA_ACT_NONE, pre=None, resid=None, out_f32=None, out_act=None) -> None:
    """v = a @ b.T with the fused epilogue of sa_gemm_nt.  a [m, k], b [n, k] (row-major, same dtype)."""
    m, k = a.shape
    n = b.shape[0]
    assert b.shape[1] == k and a.dtype == b.dtype
    ldo = None
    for t in (dot_with, pre, resid, out_f32, out_act):
        if t is not None:
            assert t.shape[0] == m and t.shape[1] == n, (t.shape, m, n)
            ld = _rowmajor(t)
            assert ldo is None or ldo == ld, "epilogue tensors must share one leading dimension"
            ldo = ld
    if ldo is None:
        ldo = n
    for t in (dot_with, pre, out_act):
        if t is not None:
            assert t.dtype == a.dtype, "act-dtype epilogue tensors must match the operand dtype"
    e = GemmEpilogue()
    e.bias = _ptr(bias); e.dot_with = _ptr(dot_with); e.dot_out = _ptr(dot_out); e.scale_dev = _ptr(scale_dev)
    e.scale = float(scale); e.act = int(act); e.pre = _ptr(pre); e.resid = _ptr(resid)
    e.out_f32 = _ptr(out_f32); e.out_act = _ptr(out_act)
    if _ops.x3_enabled() and a.dtype == torch.float32 and m > 8 and n >= 8 and k >= 8:
        # bf16x3 parity arithmetic (csrc/sa_x3.cu); m <= 8 rows (one decoding position) stay on the weight-streaming kernel
        nb = int(lib().sa_gemm_nt_x3_workspace(m, n, k))
        ws = _ops.x3_workspace(nb, a.device)
        _lib.check(lib().sa_gemm_nt_x3(m, n, k, _ptr(a), _rowmajor(a), _ptr(b), _rowmajor(b), C.byref(e), ldo, _p(ws), nb,
                                       _stream()), "sa_gemm_nt_x3")
        return
    _lib.check(lib().sa_gemm_nt(m, n, k, _dt(a.dtype), _ptr(a), _rowmajor(a), _ptr(b), _rowmajor(b), C.byref(e), ldo,
                                _stream()), "sa_gemm_nt")


def gemm_tn(a: torch.Tensor, b: torch.Tensor, d: torch.Tensor, *, scale_dev=None, scale: float = 1.0,
            accumulate: bool = False, colsum: Optional[torch.Tensor] = None) -> None:
    """d[na, nb] (+)= scale * a.T @ b;  a [m, na], b [m, nb] row-major (column slices allowed), d fp32 dense.
    colsum (fp32 [na], optional) receives a.sum(0) -- the bias gradient that goes with the weight gradient."""
    m, na = a.shape
    nb = b.shape[1]
    assert b.shape[0] == m and a.dtype == b.dtype
    assert d.dtype == torch.float32 and d.is_contiguous() and tuple(d.shape) == (na, nb)
    if colsum is not None:
        assert colsum.dtype == torch.float32 and colsum.is_contiguous() and colsum.numel() == na
        if not (_ops.x3_enabled() and a.dtype == torch.float32):
            _lib.check(lib().sa_gemm_tn_colsum(m, na, nb, _dt(a.dtype), _ptr(a), _rowmajor(a), _ptr(b), _rowmajor(b),
                                               _ptr(scale_dev), float(scale), _p(d), int(accumulate), _p(colsum), _stream()),
                       "sa_gemm_tn_colsum")
            return
        colsum.copy_(_ops.bias_grad(a))          # split operands: the sums come from the fp32 tensor itself
    if _ops.x3_enabled() and a.dtype == torch.float32 and m >= 64 and na >= 8 and nb >= 8:
        nbytes = int(lib().sa_gemm_tn_x3_workspace(m, na, nb))
        ws = _ops.x3_workspace(nbytes, a.device)
        _lib.check(lib().sa_gemm_tn_x3(m, na, nb, _ptr(a), _rowmajor(a), _ptr(b), _rowmajor(b), _ptr(scale_dev), float(scale),
                                       _p(d), int(accumulate), _p(ws), nbytes, _stream()), "sa_gemm_tn_x3")
        return
    _lib.check(lib().sa_gemm_tn(m, na, nb, _dt(a.dtype), _ptr(a), _rowmajor(a), _ptr(b), _rowmajor(b), _ptr(scale_dev),
                                float(scale), _p(d), int(accumulate), _stream()), "sa_gemm_tn")


# ------------------------------------------------------------------------------------------------
# embeddings / LayerNorm / cross-entropy
# ------------------------------------------------------------------------------------------------
def _ptr_array(ts: Sequence[torch.Tensor]):
    arr = (C.c_void_p * max(1, len(ts)))()
    for i, t in enumerate(ts):
        arr[i] = t.data_ptr()
    return arr


def embed_fwd(tokens, sp_idx, tok_w, sp_ws, pos_w, x_f32, x_act) -> None:
    B, N = tokens.shape
    dim = tok_w.shape[1]
    act = x_act if x_act is not None else x_f32
    _lib.check(lib().sa_embed_fwd(_p(tokens), _p(sp_idx), len(sp_ws), _p(tok_w), _ptr_array(sp_ws), _p(pos_w), B, N, dim,
                                  tok_w.shape[0], _p(x_f32), _p(x_act), _dt(act.dtype), _stream()), "sa_embed_fwd")


def embed_bwd(dx, tokens, sp_idx, d_tok_w, d_sp_ws, d_pos_w) -> None:
    B, N = tokens.shape
    dim = d_tok_w.shape[1]
    sp_rows = (C.c_int32 * max(1, len(d_sp_ws)))(*[int(w.shape[0]) for w in d_sp_ws])
    _lib.check(lib().sa_embed_bwd(_p(dx), _p(tokens), _p(sp_idx), len(d_sp_ws), B, N, dim, int(d_tok_w.shape[0]), sp_rows,
                                  _p(d_tok_w), _ptr_array(d_sp_ws), _p(d_pos_w), _stream()), "sa_embed_bwd")


def layernorm_fwd(x, w, b, eps, y_f32, y_act, mean, rstd) -> None:
    rows, dim = x.shape
    act = y_act if y_act is not None else y_f32
    _lib.check(lib().sa_layernorm_fwd(_p(x), _p(w), _p(b), rows, dim, float(eps), _p(y_f32), _p(y_act), _dt(act.dtype),
                                      _p(mean), _p(rstd), _stream()), "sa_layernorm_fwd")


def layernorm_bwd(dy, x, w, mean, rstd, dx, dw, db) -> None:
    rows, dim = x.shape
    _lib.check(lib().sa_layernorm_bwd(_p(dy), _p(x), _p(w), _p(mean), _p(rstd), rows, dim, _p(dx), _p(dw), _p(db),
                                      _stream()), "sa_layernorm_bwd")


def ce_fwd_bwd(logits2d, target, grad_scale, grad_scale_dev, loss_sum, dlogits2d) -> None:
    rows, V = logits2d.shape
    _lib.check(lib().sa_ce_fwd_bwd(_ptr(logits2d), _rowmajor(logits2d), _p(target), rows, V, float(grad_scale),
                                   _p(grad_scale_dev), _p(loss_sum), _ptr(dlogits2d), _stream()), "sa_ce_fwd_bwd")


def cast2d(src, dst, cols: int) -> None:
    rows = src.shape[0]
    _lib.check(lib().sa_cast2d(_ptr(src), _dt(src.dtype), _rowmajor(src), _ptr(dst), _dt(dst.dtype), _rowmajor(dst), rows,
                               cols, _stream()), "sa_cast2d")


def gate_wgrad(t, w, g, dot) -> None:
    """dot += sum t . w;  t *= g   (t: unscaled weight gradient of a ReZero-gated layer, w: its fp32 weight)"""
    assert t.dtype == torch.float32 and w.dtype == torch.float32 and t.shape == w.shape and t.is_contiguous()
    _lib.check(lib().sa_gate_wgrad(_p(t), _p(w.contiguous()), t.numel(), _p(g), _p(dot), _stream()), "sa_gate_wgrad")


def rezero_finish(colsum, bias, g, dot, dbias, dg) -> None:
    n = 0 if colsum is None else colsum.numel()
    _lib.check(lib().sa_rezero_finish(_p(colsum), _p(bias), _p(g), _p(dot), n, _p(dbias), _p(dg), _stream()),
               "sa_rezero_finish")


# ------------------------------------------------------------------------------------------------
# FAVOR+ / local attention
# ------------------------------------------------------------------------------------------------
def favor_desc(batch, seq, heads, dim_head, m, mp, ld, dtype) -> FavorDesc:
    d = FavorDesc()
    d.batch, d.seq, d.heads, d.dim_head, d.m, d.mp, d.ld, d.act_dtype = batch, seq, heads, dim_head, m, mp, ld, _dt(dtype)
    return d


def local_desc(batch, seq, heads, dim_head, window, ld, out_ld, dtype) -> LocalDesc:
    d = LocalDesc()
    d.batch, d.seq, d.heads, d.dim_head, d.window, d.ld, d.out_ld, d.act_dtype = (batch, seq, heads, dim_head, window, ld,
                                                                                  out_ld, _dt(dtype))
    return d


def favor_kmax(d, buf, col, proj, kmax) -> None:
    _lib.check(lib().sa_favor_kmax(C.byref(d), _ptr(buf, col), _p(proj), _p(kmax), _stream()), "sa_favor_kmax")


def favor_featmap_fwd(d, buf, col, proj, is_query, kmax, eps, feat, argmax) -> None:
    _lib.check(lib().sa_favor_featmap_fwd(C.byref(d), _ptr(buf, col), _p(proj), int(is_query), _p(kmax), float(eps),
                                          _p(feat), _p(argmax), _stream()), "sa_favor_featmap_fwd")


def favor_featmap_bwd(d, buf, col, proj, is_query, eps, feat, dfeat, argmax, dbuf, dcol, gsum) -> None:
    _lib.check(lib().sa_favor_featmap_bwd(C.byref(d), _ptr(buf, col), _p(proj), int(is_query), float(eps), _p(feat),
                                          _p(dfeat), _p(argmax), _ptr(dbuf, dcol), _p(gsum), _stream()),
               "sa_favor_featmap_bwd")


def favor_kmax_fixup(d, proj, kmax, gsum, dbuf, dcol) -> None:
    _lib.check(lib().sa_favor_kmax_fixup(C.byref(d), _p(proj), _p(kmax), _p(gsum), _ptr(dbuf, dcol), _stream()),
               "sa_favor_kmax_fixup")


def favor_scan_workspace(d, backward: bool) -> int:
    return int(lib().sa_favor_scan_workspace(C.byref(d), int(backward)))


def favor_scan_states_bytes(d) -> int:
    """bytes of the prefix-state buffer the forward scan can keep for the backward scan (0: the selected path recomputes)"""
    return int(lib().sa_favor_scan_states_bytes(C.byref(d)))


def favor_scan_fwd(d, qf, kf, vbuf, vcol, eps, out, ocol, den, ws, states=None) -> None:
    _lib.check(lib().sa_favor_scan_fwd_save(C.byref(d), _p(qf), _p(kf), _ptr(vbuf, vcol), float(eps), _ptr(out, ocol),
                                            _rowmajor(out), _p(den), _p(ws), ws.numel() * ws.element_size(), _p(states),
                                            0 if states is None else states.numel() * states.element_size(), _stream()),
               "sa_favor_scan_fwd")


def favor_scan_bwd(d, qf, kf, vbuf, vcol, eps, out, dout, ocol, den, dqf, dkf, dbuf, dvcol, ws, states=None) -> None:
    assert _rowmajor(out) == _rowmajor(dout)
    _lib.check(lib().sa_favor_scan_bwd_saved(C.byref(d), _p(qf), _p(kf), _ptr(vbuf, vcol), float(eps), _ptr(out, ocol),
                                             _ptr(dout, ocol), _rowmajor(out), _p(den), _p(dqf), _p(dkf),
                                             _ptr(dbuf, dvcol), _p(ws), ws.numel() * ws.element_size(), _p(states),
                                             0 if states is None else states.numel() * states.element_size(), _stream()),
               "sa_favor_scan_bwd")


def favor_scan_bwd_fused_supported(d) -> bool:
    return bool(lib().sa_favor_scan_bwd_fused_supported(C.byref(d)))


def favor_scan_bwd_fused(d, qf, kf, buf, qcol, kcol, vcol, proj, eps, eps_feature, out, dout, ocol, den, argq, dbuf, gsum, ws,
                         states=None) -> None:
    """favor_scan_bwd + favor_featmap_bwd of the queries and the keys: dq' / dk' are never stored; dbuf gets dq, dk, dv"""
    assert _rowmajor(out) == _rowmajor(dout) and _rowmajor(buf) == _rowmajor(dbuf)
    _lib.check(lib().sa_favor_scan_bwd_fused(C.byref(d), _p(qf), _p(kf), _ptr(buf, qcol), _ptr(buf, kcol), _ptr(buf, vcol),
                                             _p(proj), float(eps), float(eps_feature), _ptr(out, ocol), _ptr(dout, ocol),
                                             _rowmajor(out), _p(den), _p(argq), _ptr(dbuf, qcol), _ptr(dbuf, kcol),
                                             _ptr(dbuf, vcol), _p(gsum), _p(ws), ws.numel() * ws.element_size(), _p(states),
                                             0 if states is None else states.numel() * states.element_size(), _stream()),
               "sa_favor_scan_bwd_fused")


def rotary(buf, col, batch, seq, heads, dim_head, inv_freq, inverse: bool) -> None:
    """in-place rotary position term on `heads` head blocks of the row-major buffer starting at column `col`"""
    _lib.check(lib().sa_rotary(_ptr(buf, col), _dt(buf.dtype), _rowmajor(buf), batch, seq, heads, dim_head, _p(inv_freq),
                               int(inverse), _stream()), "sa_rotary")


def rotary_qk(buf, qcol, kcol, batch, seq, heads, dim_head, inv_freq, inverse: bool) -> None:
    """rotary position term on the q block (column qcol) and the k block (column kcol) of one buffer, one launch"""
    _lib.check(lib().sa_rotary_qk(_ptr(buf, qcol), _dt(buf.dtype), _rowmajor(buf), kcol - qcol, batch, seq, heads, dim_head,
                                  _p(inv_freq), int(inverse), _stream()), "sa_rotary_qk")


def favor_decode_step(batch, heads, m, t, t_dev, buf, qcol, kcol, vcol, proj, eps, eps_cumsum, mhist, scratch, Se, ze, S1,
                      out, ocol) -> None:
    """advance the FAVOR+ state of one layer by the position whose q / k / v rows are in `buf` ([batch, ld])"""
    _lib.check(lib().sa_favor_decode_step(batch, heads, m, _dt(buf.dtype), t, _p(t_dev), _ptr(buf, qcol), _ptr(buf, kcol),
                                          _ptr(buf, vcol), _rowmajor(buf), _p(proj), float(eps), float(eps_cumsum),
                                          _p(mhist), _p(scratch), _p(Se), _p(ze), _p(S1), _ptr(out, ocol), _rowmajor(out),
                                          _stream()), "sa_favor_decode_step")


def local_decode_step(batch, heads, window, p, p_dev, nmax, buf, qcol, kcol, vcol, inv_freq, kcache, vcache, out,
                      ocol) -> None:
    """append position p to the local-head caches of one layer and attend its window"""
    _lib.check(lib().sa_local_decode_step(batch, heads, window, _dt(buf.dtype), p, _p(p_dev), nmax, _ptr(buf, qcol),
                                          _ptr(buf, kcol), _ptr(buf, vcol), _rowmajor(buf), _p(inv_freq), _p(kcache),
                                          _p(vcache), _ptr(out, ocol), _rowmajor(out), _stream()), "sa_local_decode_step")


def embed_step(tokens, sp_idx, tok_w, sp_ws, pos_w, t, t_dev, x_f32, x_act) -> None:
    """embedding of one position (host t, or the device int t_dev) for a [batch] vector of tokens"""
    act = x_act if x_act is not None else x_f32
    _lib.check(lib().sa_embed_step(_p(tokens), _p(sp_idx), len(sp_ws), 0 if sp_idx is None else sp_idx.shape[1], _p(tok_w),
                                   _ptr_array(sp_ws), _p(pos_w), tokens.shape[0], tok_w.shape[1], t, _p(t_dev), _p(x_f32),
                                   _p(x_act), _dt(act.dtype), _stream()), "sa_embed_step")


def local_attn_fwd(d, buf, qcol, kcol, vcol, inv_freq, out, ocol, lse) -> None:
    _lib.check(lib().sa_local_attn_fwd(C.byref(d), _ptr(buf, qcol), _ptr(buf, kcol), _ptr(buf, vcol), _p(inv_freq),
                                       _ptr(out, ocol), _p(lse), _stream()), "sa_local_attn_fwd")


def local_attn_bwd(d, buf, qcol, kcol, vcol, inv_freq, out, dout, ocol, lse, dbuf) -> None:
    delta_ws = torch.empty_like(lse)       # scratch: delta = dO . O per query, handed from the dq to the dk/dv kernel
    _lib.check(lib().sa_local_attn_bwd_ws(C.byref(d), _ptr(buf, qcol), _ptr(buf, kcol), _ptr(buf, vcol), _p(inv_freq),
                                          _ptr(out, ocol), _ptr(dout, ocol), _p(lse), _ptr(dbuf, qcol), _ptr(dbuf, kcol),
                                          _ptr(dbuf, vcol), _p(delta_ws), _stream()), "sa_local_attn_bwd")


def rotary_table(inv_freq: torch.Tensor, seq: int, dim_head: int) -> torch.Tensor:
    """[seq, dim_head / 2, 2] fp32 (cos, sin) of n * inv_freq, evaluated as the rotary kernels evaluate them"""
    table = torch.empty((seq, dim_head // 2, 2), device=inv_freq.device, dtype=torch.float32)
    _lib.check(lib().sa_rotary_table(_p(inv_freq), seq, dim_head, _p(table), _stream()), "sa_rotary_table")
    return table


def local_attn_bwd_rot(d, buf, qcol, kcol, vcol, inv_freq, rot_table, out, dout, ocol, lse, dbuf) -> None:
    """local_attn_bwd for q / k rotated in place before the forward call: dq / dk leave through the rotation's transpose"""
    delta_ws = torch.empty_like(lse)
    _lib.check(lib().sa_local_attn_bwd_rot(C.byref(d), _ptr(buf, qcol), _ptr(buf, kcol), _ptr(buf, vcol), _p(inv_freq),
                                           _p(rot_table), _ptr(out, ocol), _ptr(dout, ocol), _p(lse), _ptr(dbuf, qcol),
                                           _ptr(dbuf, kcol), _ptr(dbuf, vcol), _p(delta_ws), _stream()),
               "sa_local_attn_bwd_rot")


# ------------------------------------------------------------------------------------------------
# token plumbing between the two models
# ------------------------------------------------------------------------------------------------
_TOK_DTYPES = {torch.uint16: 0, torch.int16: 0, torch.int32: 1, torch.int64: 2}     # sa_tok_dtype (int16 storage read as uint16)


def tokens_prepare(grid: torch.Tensor, order: torch.Tensor, bos: int):
    """(x_in, y) int64 [B, N] from a latent index grid [B, ...] in its stored type (uint16 / int32 / int64), one launch"""
    assert grid.is_contiguous() and order.dtype == torch.int64 and order.is_contiguous()
    B = grid.shape[0]
    n_src = grid.numel() // max(B, 1)
    n = order.numel()
    x_in = torch.empty(B, n, dtype=torch.int64, device=grid.device)
    y = torch.empty(B, n, dtype=torch.int64, device=grid.device)
    if B == 0 or n == 0:
        return x_in, y
    _lib.check(lib().sa_tokens_prepare(_ptr(grid), _TOK_DTYPES[grid.dtype], B, n_src, _ptr(order), n, int(bos), _ptr(x_in),
                                       _ptr(y), _stream()), "sa_tokens_prepare")
    return x_in, y


def tokens_gather(src: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """out[b][i] = src[b][index[i]] (int64 out); src [B, n_src] in uint16 / int32 / int64"""
    assert src.is_contiguous() and index.dtype == torch.int64 and index.is_contiguous()
    B = src.shape[0]
    n_src = src.numel() // max(B, 1)
    out = torch.empty(B, index.numel(), dtype=torch.int64, device=src.device)
    if out.numel() == 0:
        return out
    _lib.check(lib().sa_tokens_gather(_ptr(src), _TOK_DTYPES[src.dtype], B, n_src, _ptr(index), index.numel(), _ptr(out),
                                      _stream()), "sa_tokens_gather")
    return out


def tokens_narrow(idx: torch.Tensor) -> torch.Tensor:
    """int64 indices -> uint16 (the on-disk token type); raises if a value does not fit"""
    assert idx.dtype == torch.int64 and idx.is_contiguous()
    out = torch.empty(idx.shape, dtype=torch.uint16, device=idx.device)
    if idx.numel() == 0:
        return out
    bad = torch.zeros(1, dtype=torch.int32, device=idx.device)
    _lib.check(lib().sa_tokens_narrow(_ptr(idx), idx.numel(), _ptr(out), _ptr(bad), _stream()), "sa_tokens_narrow")
    if int(bad.item()):
        raise ValueError("token index outside [0, 65535] cannot be stored as uint16")
    return out


# ------------------------------------------------------------------------------------------------
# optional per-launch CUDA-event timing (bench.py: roofline of the dominant kernel, per-kernel breakdown)
# ------------------------------------------------------------------------------------------------
class KernelTimer:
    """Records (name, start event, end event) around every wrapper call whose name passes `match`.
    Events are recorded on torch's current stream, the stream every kernel of this library is enqueued on."""

    def __init__(self, match=None):
        self.match = match
        self.records = []

    def summary(self):
        out = {}
        for name, e0, e1, meta in self.records:
            ms = e0.elapsed_time(e1)
            d = out.setdefault(name, {"ms": 0.0, "launches": 0, "flop": 0.0, "bytes": 0.0})
            d["ms"] += ms; d["launches"] += 1; d["flop"] += meta[0]; d["bytes"] += meta[1]
        return out


_TIMER: Optional[KernelTimer] = None


def set_timer(timer: Optional[KernelTimer]) -> None:
    global _TIMER
    _TIMER = timer


def _meta(name, args, kwargs):
    """(algorithmic FLOPs, algorithmic HBM bytes) of one call: operands read once + every epilogue tensor the call
    names read / written once"""
    if name == "gemm_nt":
        a, b = args[0], args[1]
        m, k, n = a.shape[0], a.shape[1], b.shape[0]
        by = (m * k + n * k) * a.element_size()
        for key in ("dot_with", "pre", "resid", "out_f32", "out_act"):
            t = kwargs.get(key)
            if t is not None:
                by += m * n * t.element_size()
        if kwargs.get("bias") is not None:
            by += n * 4
        return 2.0 * m * k * n, float(by)
    if name == "gemm_tn":
        a, b = args[0], args[1]
        m, na, nb = a.shape[0], a.shape[1], b.shape[1]
        return 2.0 * m * na * nb, float(m * (na + nb) * a.element_size() + na * nb * 4)
    return 0.0, 0.0


def _instrument(name, fn):
    def wrapper(*args, **kwargs):
        t = _TIMER
        if t is None or (t.match is not None and not t.match(name)):
            return fn(*args, **kwargs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*args, **kwargs)
        e1.record()
        t.records.append((name, e0, e1, _meta(name, args, kwargs)))
        return r
    wrapper.__name__ = fn.__name__
    wrapper.__doc__ = fn.__doc__
    return wrapper


for _n in ("gemm_nt", "gemm_tn", "embed_fwd", "embed_bwd", "layernorm_fwd", "layernorm_bwd", "ce_fwd_bwd", "cast2d",
           "rotary", "rotary_qk", "favor_kmax", "favor_featmap_fwd", "favor_featmap_bwd", "favor_scan_fwd", "favor_scan_bwd", "favor_scan_bwd_fused", "local_attn_fwd", "rotary_table", "local_attn_bwd_rot",
           "local_attn_bwd"):
    globals()[_n] = _instrument(_n, globals()[_n])
