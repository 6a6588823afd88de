"""Tensor-level wrappers over the C ABI (include/synthanatomy_b200.h).

PyTorch is used here only as plumbing: device memory (caching allocator), the current CUDA stream and
dtype bookkeeping.  All arithmetic happens inside libsynthanatomy_b200.so.  Every function raises if the
tensors are not CUDA tensors -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import SA_BF16, SA_F32, ConvDesc


def lib():
    return _lib.load()


def _dt(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return SA_F32
    if dtype == torch.bfloat16:
        return SA_BF16
    raise TypeError(f"synthanatomy_b200: unsupported dtype {dtype} (float32 / bfloat16 only)")


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("synthanatomy_b200: CUDA tensors only -- there is no CPU fallback")
    if not t.is_contiguous():
        raise RuntimeError("synthanatomy_b200: tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(lib().sa_launch_count())


def reset_launch_count() -> None:
    lib().sa_launch_count_reset()


def last_path() -> int:
    return int(lib().sa_last_path())


_FORCE_SIMT = False


def set_force_simt(on: bool) -> None:
    global _FORCE_SIMT
    _FORCE_SIMT = bool(on)
    lib().sa_set_force_simt(1 if on else 0)


# Deterministic mode: the reference's `deterministic=True` (run_vqvae.py:550 -> src/utils/general.py:333) sets
# torch.backends.cudnn.deterministic; the two networks call sync_deterministic() at the top of forward, so the same
# switch makes this library add its split-K partials / column sums / loss sums in a fixed order (csrc/sa_common.cuh).
_DETERMINISTIC: Optional[bool] = None          # None: follow torch's flags


def set_deterministic(on: Optional[bool]) -> None:
    """True / False: force the mode; None: follow torch.backends.cudnn.deterministic / use_deterministic_algorithms"""
    global _DETERMINISTIC
    _DETERMINISTIC = on
    sync_deterministic()


def sync_deterministic() -> bool:
    on = _DETERMINISTIC
    if on is None:
        on = bool(torch.backends.cudnn.deterministic) or torch.are_deterministic_algorithms_enabled()
    if bool(lib().sa_get_deterministic()) != on:
        lib().sa_set_deterministic(1 if on else 0)
    return on


# ------------------------------------------------------------------------------------------------
# bf16x3 "parity" arithmetic (csrc/sa_x3.cu): fp32 tensors, products on the bf16 tensor cores as hi.hi + lo.hi + hi.lo.
# `compute_dtype=BF16X3` on the two networks selects it; while the mode is on, every fp32 conv / dense call whose shape
# the tcgen05 kernels take goes through the *_x3 entry points (the rest stays on the CUDA-core fp32 kernels).
# ------------------------------------------------------------------------------------------------
BF16X3 = "bf16x3"
_X3 = [False]
_X3_WS = {}


class x3_mode:
    def __init__(self, on: bool):
        self.on = bool(on)

    def __enter__(self):
        self.prev = _X3[0]
        _X3[0] = self.on
        return self

    def __exit__(self, *exc):
        _X3[0] = self.prev
        return False


def x3_enabled() -> bool:
    return _X3[0] and not _FORCE_SIMT


def x3_workspace(nbytes: int, device) -> torch.Tensor:
    """grow-only scratch buffer for the split operands (all launches are on one stream, so one buffer is enough)"""
    key = device.index
    buf = _X3_WS.get(key)
    if buf is None or buf.numel() < nbytes:
        _X3_WS.pop(key, None)
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _X3_WS[key] = buf
    return buf


def resolve_dtype(compute_dtype):
    """(tensor dtype, x3 flag) of a `compute_dtype` argument (a torch dtype or BF16X3)"""
    if isinstance(compute_dtype, str):
        if compute_dtype != BF16X3:
            raise ValueError(f"compute_dtype must be a torch dtype or {BF16X3!r}, got {compute_dtype!r}")
        return torch.float32, True
    return compute_dtype, False


# ------------------------------------------------------------------------------------------------
# conv geometry
# ------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ConvSpec:
    """One nn.Conv3d (kind='conv') or nn.ConvTranspose3d (kind='deconv') of the reference
    (src/networks/vqvae/baseline.py:153-156, 218-227, 242-244, 258, 283-293); cubic kernel, dilation 1."""
    kind: str
    cin: int
    cout: int
    k: int
    s: int
    p: int

    def out_dhw(self, in_dhw: Sequence[int]) -> Tuple[int, int, int]:
        if self.kind == "conv":
            return tuple((i + 2 * self.p - self.k) // self.s + 1 for i in in_dhw)
        return tuple((i - 1) * self.s - 2 * self.p + self.k for i in in_dhw)


def _desc(batch, in_dhw, out_dhw, c_in, c_out, k, s, p, transposed, dtype) -> ConvDesc:
    d = ConvDesc()
    d.batch = batch
    for i in range(3):
        d.in_dhw[i] = in_dhw[i]
        d.out_dhw[i] = out_dhw[i]
    d.c_in, d.c_out, d.ksize, d.stride, d.pad = c_in, c_out, k, s, p
    d.transposed = transposed
    d.act_dtype = _dt(dtype)
    return d


class PackPlan:
    """The packed forms (forward and transposed, per dtype) of one stack's conv weights, refreshed by a few multi-tensor
    launches at the top of a pass instead of two `sa_pack_weight` launches per conv and step.  The first pass through a
    stack records what it packs; `begin()` of every later pass re-packs all of it into the same buffers and stamps the
    entries with the pass number, and `pack_weight` hands out an entry only while its stamp is the current one -- anything
    not (or no longer) covered is packed the old way and recorded."""

    def __init__(self):
        self.entries = {}            # key -> [w, out, A, B, taps, transpose, flip, stamp]
        self.epoch = 0
        self._tables = None          # dtype -> (ctypes item array, n), rebuilt when the entries change

    def begin(self) -> None:
        self.epoch += 1
        if not self.entries:
            return
        if self._tables is None:
            by_dtype = {}
            for e in self.entries.values():
                by_dtype.setdefault(e[1].dtype, []).append(e)
            self._tables = {}
            for dt, es in by_dtype.items():
                arr = (_lib.WPackItem * len(es))()
                for a, (w, out, A, B, taps, tr, fl, _) in zip(arr, es):
                    a.src, a.dst, a.A, a.B, a.taps, a.transpose, a.flip = w.data_ptr(), out.data_ptr(), A, B, taps, int(tr), int(fl)
                self._tables[dt] = (arr, len(es))
        for dt, (arr, n) in self._tables.items():
            _lib.check(lib().sa_pack_weight_multi(arr, n, _dt(dt), _stream()), "sa_pack_weight_multi")
        for e in self.entries.values():
            e[7] = self.epoch

    def lookup(self, key):
        e = self.entries.get(key)
        return e[1] if e is not None and e[7] == self.epoch else None

    def record(self, key, w, out, A, B, taps, transpose, flip) -> None:
        if len(self.entries) >= 1024:        # weights that move every pass (a dtype-converted copy) must not pile up here
            self.entries.clear()
        self.entries[key] = [w, out, A, B, taps, transpose, flip, self.epoch]
        self._tables = None


_PACK_PLAN: Optional[PackPlan] = None


class pack_plan:
    """`with pack_plan(plan):` -- pack_weight calls inside consult / feed `plan`; begin=True refreshes it first (top of a
    forward pass), begin=False only makes it current (the backward pass of ops recorded under it)."""

    def __init__(self, plan: Optional[PackPlan], begin: bool = True):
        self.plan, self.begin = plan, begin

    def __enter__(self):
        global _PACK_PLAN
        self.prev = _PACK_PLAN
        _PACK_PLAN = self.plan
        if self.plan is not None and self.begin:
            self.plan.begin()
        return self.plan

    def __exit__(self, *exc):
        global _PACK_PLAN
        _PACK_PLAN = self.prev
        return False


def current_pack_plan() -> Optional[PackPlan]:
    return _PACK_PLAN


def pack_weight(w: torch.Tensor, transpose: bool, dtype: torch.dtype, flip: bool = False) -> torch.Tensor:
    """torch layout [A][B][k,k,k] fp32 -> packed [taps][A][B] (or [taps][B][A] if transpose) in `dtype`."""
    A, B = w.shape[0], w.shape[1]
    taps = w[0, 0].numel()
    plan = _PACK_PLAN
    key = None
    if plan is not None:
        key = (w.data_ptr(), tuple(w.shape), bool(transpose), dtype, bool(flip))
        hit = plan.lookup(key)
        if hit is not None:
            return hit
    R, Cc = (B, A) if transpose else (A, B)
    out = torch.empty((taps, R, Cc), device=w.device, dtype=dtype)
    _lib.check(lib().sa_pack_weight(_p(w), A, B, taps, int(transpose), int(flip), _p(out), _dt(dtype), _stream()),
               "sa_pack_weight")
    if plan is not None:
        plan.record(key, w, out, A, B, taps, bool(transpose), bool(flip))
    return out


def unpack_wgrad(dwp: torch.Tensor, like: torch.Tensor, transpose: bool, flip: bool = False) -> torch.Tensor:
    """packed fp32 [taps][..][..] -> torch layout gradient shaped like `like`."""
    A, B = like.shape[0], like.shape[1]
    taps = like[0, 0].numel()
    out = torch.empty_like(like, dtype=torch.float32)
    _lib.check(lib().sa_unpack_wgrad(_p(dwp), A, B, taps, int(transpose), int(flip), _p(out), 0, _stream()),
               "sa_unpack_wgrad")
    return out


class ConvTimer:
    """Optional per-launch CUDA-event timing of sa_conv3d_fwd calls whose descriptor matches `match`
    (bench.py uses it to time the dominant kernel inside the timed region, on the launching stream)."""

    def __init__(self, match):
        self.match = match
        self.events = []

    def elapsed_ms(self):
        return [a.elapsed_time(b) for a, b in self.events]


_TIMER: Optional[ConvTimer] = None


def set_conv_timer(timer: Optional[ConvTimer]) -> None:
    global _TIMER
    _TIMER = timer


def _run_fwd(d: ConvDesc, x, wp, bias, addend, mask, relu, out_shape):
    y = torch.empty(out_shape, device=x.device, dtype=x.dtype)
    timed = _TIMER is not None and _TIMER.match(d)
    if timed:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if x3_enabled() and x.dtype == torch.float32 and lib().sa_conv3d_x3_supported(C.byref(d), 0):
        nb = int(lib().sa_conv3d_x3_workspace(C.byref(d), 0))
        ws = x3_workspace(nb, x.device)
        _lib.check(lib().sa_conv3d_fwd_x3(C.byref(d), _p(x), _p(wp), _p(bias), _p(addend), _p(mask), int(relu), _p(y),
                                          _p(ws), nb, _stream()), "sa_conv3d_fwd_x3")
    else:
        _lib.check(lib().sa_conv3d_fwd(C.byref(d), _p(x), _p(wp), _p(bias), _p(addend), _p(mask), int(relu), _p(y),
                                       _stream()), "sa_conv3d_fwd")
    if timed:
        e1.record()
        _TIMER.events.append((e0, e1))
    return y


def conv_forward(spec: ConvSpec, x: torch.Tensor, wp: torch.Tensor, bias: Optional[torch.Tensor],
                 addend: Optional[torch.Tensor] = None, relu: bool = False,
                 mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: NDHWC [B, D, H, W, cin]; wp: pack_weight(weight, transpose=(kind == 'deconv')).
    Returns act( conv(x) + bias (+ addend) ) as NDHWC [B, oD, oH, oW, cout]."""
    B, in_dhw = x.shape[0], tuple(x.shape[1:4])
    assert x.shape[4] == spec.cin, (x.shape, spec)
    out_dhw = spec.out_dhw(in_dhw)
    d = _desc(B, in_dhw, out_dhw, spec.cin, spec.cout, spec.k, spec.s, spec.p, int(spec.kind == "deconv"), x.dtype)
    return _run_fwd(d, x, wp, bias, addend, mask, relu, (B, *out_dhw, spec.cout))


def conv_dgrad(spec: ConvSpec, dy: torch.Tensor, wp_t: torch.Tensor, in_dhw: Sequence[int],
               addend: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Data gradient: dx = (dgrad(dy) (+ addend)) * (mask > 0).
    wp_t: pack_weight(weight, transpose=(kind == 'conv'))  -- the opposite packing of the forward."""
    B, out_dhw = dy.shape[0], tuple(dy.shape[1:4])
    assert dy.shape[4] == spec.cout
    # the dgrad of FORM_CONV is FORM_TCONV over dy and vice versa; X of the primitive is dy
    d = _desc(B, out_dhw, tuple(in_dhw), spec.cout, spec.cin, spec.k, spec.s, spec.p, int(spec.kind == "conv"), dy.dtype)
    return _run_fwd(d, dy, wp_t, None, addend, mask, False, (B, *in_dhw, spec.cin))


def conv_wgrad(spec: ConvSpec, x: torch.Tensor, dy: torch.Tensor, weight_like: torch.Tensor) -> torch.Tensor:
    """Weight gradient in the torch layout of `weight_like` (fp32)."""
    B = x.shape[0]
    in_dhw, out_dhw = tuple(x.shape[1:4]), tuple(dy.shape[1:4])
    taps = spec.k ** 3
    if spec.kind == "conv":
        d = _desc(B, in_dhw, out_dhw, spec.cin, spec.cout, spec.k, spec.s, spec.p, 0, x.dtype)
        pp, qq = dy, x
    else:  # the strided gather runs over dy; P = x
        d = _desc(B, out_dhw, in_dhw, spec.cout, spec.cin, spec.k, spec.s, spec.p, 0, x.dtype)
        pp, qq = x, dy
    dwp = torch.empty((taps, d.c_out, d.c_in), device=x.device, dtype=torch.float32)
    if x3_enabled() and x.dtype == torch.float32 and lib().sa_conv3d_x3_supported(C.byref(d), 1):
        nb = int(lib().sa_conv3d_x3_workspace(C.byref(d), 1))
        ws = x3_workspace(nb, x.device)
        _lib.check(lib().sa_conv3d_wgrad_x3(C.byref(d), _p(pp), _p(qq), _p(dwp), 0, _p(ws), nb, _stream()),
                   "sa_conv3d_wgrad_x3")
    else:
        _lib.check(lib().sa_conv3d_wgrad(C.byref(d), _p(pp), _p(qq), _p(dwp), 0, _stream()), "sa_conv3d_wgrad")
    return unpack_wgrad(dwp, weight_like, transpose=False)


def conv1x1_bwd_fused_supported(spec: ConvSpec, g: torch.Tensor) -> bool:
    return (spec.k == 1 and spec.s == 1 and spec.p == 0 and spec.kind == "conv" and spec.cin == 128 and spec.cout == 128
            and g.dtype == torch.bfloat16 and not _FORCE_SIMT)


def conv1x1_bwd_fused(spec: ConvSpec, g: torch.Tensor, h: torch.Tensor, wp_t: torch.Tensor, weight_like: torch.Tensor,
                      with_dbh: bool = False):
    """One pass over g and h: dh = dgrad(g) * (h > 0), dW (torch layout of `weight_like`, fp32) and db; with_dbh: also the
    column sums of dh (the bias gradient of the conv that produced h) as a fourth result."""
    m = g.numel() // spec.cout
    dh = torch.empty_like(h)
    # one zeroed buffer: dW | db | dbh
    acc = torch.zeros((spec.cout * spec.cin + spec.cout + spec.cin,), device=g.device, dtype=torch.float32)
    dwp = acc[: spec.cout * spec.cin].view(1, spec.cout, spec.cin)
    db = acc[spec.cout * spec.cin: spec.cout * spec.cin + spec.cout]
    dbh = acc[spec.cout * spec.cin + spec.cout:]
    _lib.check(lib().sa_conv1x1_bwd_fused_dbh(m, spec.cout, spec.cin, _p(g), _p(h), _p(wp_t), _p(dh), _p(dwp), _p(db),
                                              _p(dbh) if with_dbh else None, _stream()), "sa_conv1x1_bwd_fused")
    dw = unpack_wgrad(dwp, weight_like, transpose=False)
    return (dh, dw, db, dbh) if with_dbh else (dh, dw, db)


def conv1x1_fwd_fused(spec: ConvSpec, x: torch.Tensor, wp: torch.Tensor, bias: Optional[torch.Tensor],
                      addend: torch.Tensor, relu: bool) -> torch.Tensor:
    """y = relu?(conv1x1(x) + bias + addend) through the streaming pointwise kernel (128 -> 128 channels, bf16)."""
    m = x.numel() // spec.cin
    y = torch.empty_like(addend)
    _lib.check(lib().sa_conv1x1_fwd_fused(m, spec.cout, spec.cin, _p(x), _p(wp), _p(bias), _p(addend), int(relu), _p(y),
                                          _stream()), "sa_conv1x1_fwd_fused")
    return y


def _i3(v):
    return (C.c_int * 3)(*[int(a) for a in v])


def im2col_c1(x: torch.Tensor, k: int, s: int, p: int, out_dhw: Sequence[int]) -> torch.Tensor:
    """x: [B, D, H, W, 1] -> cols [B, oD, oH, oW, k^3] with cols[b, o, t] = x[b, o*s - p + t]."""
    assert x.shape[-1] == 1
    B, in_dhw = x.shape[0], tuple(x.shape[1:4])
    cols = torch.empty((B, *out_dhw, k ** 3), device=x.device, dtype=x.dtype)
    _lib.check(lib().sa_im2col_c1(_p(x), _dt(x.dtype), B, _i3(in_dhw), _i3(out_dhw), k, s, p, _p(cols), _stream()),
               "sa_im2col_c1")
    return cols


def col2im_c1(cols: torch.Tensor, k: int, s: int, p: int, out_dhw: Sequence[int],
              bias: Optional[torch.Tensor]) -> torch.Tensor:
    """cols: [B, iD, iH, iW, k^3] -> y [B, oD, oH, oW, 1] (transposed-conv scatter written as a gather)."""
    assert cols.shape[-1] == k ** 3
    B, in_dhw = cols.shape[0], tuple(cols.shape[1:4])
    y = torch.empty((B, *out_dhw, 1), device=cols.device, dtype=cols.dtype)
    _lib.check(lib().sa_col2im_c1(_p(cols), _dt(cols.dtype), B, _i3(in_dhw), _i3(out_dhw), k, s, p, _p(bias), _p(y),
                                  _stream()), "sa_col2im_c1")
    return y


def bias_grad(dy: torch.Tensor) -> torch.Tensor:
    c = dy.shape[-1]
    rows = dy.numel() // c
    db = torch.empty((c,), device=dy.device, dtype=torch.float32)
    _lib.check(lib().sa_bias_grad(_p(dy), rows, c, _dt(dy.dtype), _p(db), 0, _stream()), "sa_bias_grad")
    return db


# ------------------------------------------------------------------------------------------------
# BatchNorm3d + LeakyReLU (PatchGAN discriminator blocks), channels-last activations
# ------------------------------------------------------------------------------------------------
def _bn_ws(c: int, device) -> torch.Tensor:
    return torch.empty(int(lib().sa_bn_workspace(c)), dtype=torch.uint8, device=device)


def bn_stats(x: torch.Tensor, eps: float, momentum: float, running_mean: Optional[torch.Tensor],
             running_var: Optional[torch.Tensor]):
    """training-mode statistics of x [..., C]: (mean, rstd) fp32 [C]; the running buffers are updated in place"""
    c = x.shape[-1]
    rows = x.numel() // c
    mean = torch.empty(c, device=x.device, dtype=torch.float32)
    rstd = torch.empty(c, device=x.device, dtype=torch.float32)
    _lib.check(lib().sa_bn_stats(_p(x), _dt(x.dtype), rows, c, _p(_bn_ws(c, x.device)), float(eps), float(momentum), _p(mean),
                                 _p(rstd), _p(running_mean), _p(running_var), _stream()), "sa_bn_stats")
    return mean, rstd


def bn_eval_stats(running_mean: torch.Tensor, running_var: torch.Tensor, eps: float):
    c = running_mean.numel()
    mean = torch.empty_like(running_mean)
    rstd = torch.empty_like(running_mean)
    _lib.check(lib().sa_bn_eval_stats(_p(running_mean), _p(running_var), c, float(eps), _p(mean), _p(rstd), _stream()),
               "sa_bn_eval_stats")
    return mean, rstd


def bn_lrelu_fwd(x, mean, rstd, gamma, beta, slope: float) -> torch.Tensor:
    c = x.shape[-1]
    y = torch.empty_like(x)
    _lib.check(lib().sa_bn_lrelu_fwd(_p(x), _dt(x.dtype), x.numel() // c, c, _p(mean), _p(rstd), _p(gamma), _p(beta),
                                     float(slope), _p(y), _stream()), "sa_bn_lrelu_fwd")
    return y


def bn_lrelu_bwd(g, x, y, mean, rstd, gamma, slope: float, need_dx: bool = True):
    """(dx | None, dgamma, dbeta) of y = lrelu(bn(x)) in training mode"""
    c = x.shape[-1]
    dgamma = torch.empty(c, device=x.device, dtype=torch.float32)
    dbeta = torch.empty(c, device=x.device, dtype=torch.float32)
    dx = torch.empty_like(x) if need_dx else None
    _lib.check(lib().sa_bn_lrelu_bwd(_p(g), _p(x), _p(y), _dt(x.dtype), x.numel() // c, c, _p(mean), _p(rstd), _p(gamma),
                                     float(slope), _p(_bn_ws(c, x.device)), _p(dgamma), _p(dbeta), _p(dx), _stream()),
               "sa_bn_lrelu_bwd")
    return dx, dgamma, dbeta


def lrelu_fwd_(x: torch.Tensor, slope: float) -> torch.Tensor:
    _lib.check(lib().sa_lrelu_fwd(_p(x), _dt(x.dtype), x.numel(), float(slope), _stream()), "sa_lrelu_fwd")
    return x


def lrelu_bwd_(g: torch.Tensor, y: torch.Tensor, slope: float) -> torch.Tensor:
    _lib.check(lib().sa_lrelu_bwd(_p(g), _p(y), _dt(g.dtype), g.numel(), float(slope), _stream()), "sa_lrelu_bwd")
    return g


# ------------------------------------------------------------------------------------------------
# layout
# ------------------------------------------------------------------------------------------------
def ncdhw_to_ndhwc(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    B, Cc = x.shape[0], x.shape[1]
    sp = tuple(x.shape[2:])
    S = 1
    for v in sp:
        S *= v
    out = torch.empty((B, *sp, Cc), device=x.device, dtype=dtype)
    _lib.check(lib().sa_nchw_to_nhwc(_p(x), _dt(x.dtype), _p(out), _dt(dtype), B, Cc, S, _stream()), "sa_nchw_to_nhwc")
    return out


def ndhwc_to_ncdhw(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    B, Cc = x.shape[0], x.shape[-1]
    sp = tuple(x.shape[1:-1])
    S = 1
    for v in sp:
        S *= v
    out = torch.empty((B, Cc, *sp), device=x.device, dtype=dtype)
    _lib.check(lib().sa_nhwc_to_nchw(_p(x), _dt(x.dtype), _p(out), _dt(dtype), B, Cc, S, _stream()), "sa_nhwc_to_nchw")
    return out


# ------------------------------------------------------------------------------------------------
# vector quantiser
# ------------------------------------------------------------------------------------------------
def vq_forward(z_flat: torch.Tensor, codebook: torch.Tensor, stats: Optional[torch.Tensor], straight_through: bool):
    """z_flat [rows, dim] fp32, codebook [K, dim] fp32.  stats: zeroed fp32 [K + K*dim + 1] receiving
    (counts | dw | sse), or None.  Returns (idx int64 [rows], q [rows, dim])."""
    rows, dim = z_flat.shape
    K = codebook.shape[0]
    idx = torch.empty((rows,), device=z_flat.device, dtype=torch.int64)
    q = torch.empty_like(z_flat)
    counts = dw = sse = None
    if stats is not None:
        assert stats.numel() == K + K * dim + 1 and stats.dtype == torch.float32
        base = stats.data_ptr()
        counts, dw, sse = C.c_void_p(base), C.c_void_p(base + 4 * K), C.c_void_p(base + 4 * (K + K * dim))
    _lib.check(lib().sa_vq_forward(_p(z_flat), _p(codebook), rows, dim, K, _p(idx), _p(q), int(straight_through), counts,
                                   dw, sse, _stream()), "sa_vq_forward")
    return idx, q


def vq_ema_update(N, embed_avg, codebook, stats, decay: float, eps: float) -> None:
    K, dim = codebook.shape
    base = stats.data_ptr()
    _lib.check(lib().sa_vq_ema_update(_p(N), _p(embed_avg), _p(codebook), C.c_void_p(base), C.c_void_p(base + 4 * K), K,
                                      dim, float(decay), float(eps), None, _stream()), "sa_vq_ema_update")


def vq_perplexity(counts: torch.Tensor, total: int) -> torch.Tensor:
    out = torch.empty((), device=counts.device, dtype=torch.float32)
    _lib.check(lib().sa_vq_perplexity(_p(counts), counts.numel(), float(total), _p(out), _stream()), "sa_vq_perplexity")
    return out


def vq_embed(idx: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    K, dim = codebook.shape
    rows = idx.numel()
    q = torch.empty((rows, dim), device=codebook.device, dtype=torch.float32)
    _lib.check(lib().sa_vq_embed(_p(idx), _p(codebook), rows, dim, K, _p(q), _stream()), "sa_vq_embed")
    return q


def vq_backward(g_q, g_loss, z, q, coef: float) -> torch.Tensor:
    dz = torch.empty_like(z)
    _lib.check(lib().sa_vq_backward(_p(g_q), _p(g_loss), _p(z), _p(q), float(coef), z.numel(), _p(dz), _stream()),
               "sa_vq_backward")
    return dz


# ------------------------------------------------------------------------------------------------
# losses / optimiser
# ------------------------------------------------------------------------------------------------
def mse_fwd_bwd(pred: torch.Tensor, target: torch.Tensor, sse: Optional[torch.Tensor], grad: Optional[torch.Tensor],
                scale: float, scale_dev: Optional[torch.Tensor] = None) -> None:
    _lib.check(lib().sa_mse_fwd_bwd(_p(pred), _dt(pred.dtype), _p(target), pred.numel(), float(scale), _p(scale_dev),
                                    _p(sse), _p(grad), _stream()), "sa_mse_fwd_bwd")


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step) -> None:
    _lib.check(lib().sa_adam_step(_p(p), _p(g), _p(m), _p(v), p.numel(), float(lr), float(beta1), float(beta2),
                                  float(eps), int(step), _stream()), "sa_adam_step")
