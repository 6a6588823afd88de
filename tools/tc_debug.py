"""GPU diagnostic: tcgen05 kernels vs the CUDA-core kernels on identical bf16 inputs, with error structure dumps.
Usage (on the GPU box):  python tools/tc_debug.py [fwd|wgrad|all]  > gpurun_out/tc_debug.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthanatomy_b200 import ops  # noqa: E402


def report(tag, a, b):
    a, b = a.float(), b.float()
    err = (a - b).abs()
    scale = b.abs().max().item() + 1e-12
    bad = err > 2e-2 * scale
    print(f"{tag}: max_err={err.max().item():.4e} scale={scale:.3e} bad_frac={bad.float().mean().item():.4f} "
          f"finite={torch.isfinite(a).all().item()}", flush=True)
    return bad


def structure(bad, names):
    # fraction of bad entries along each axis index (first 16 entries)
    for ax, nm in enumerate(names):
        dims = [d for d in range(bad.dim()) if d != ax]
        fr = bad.float().mean(dim=dims)
        print(f"   bad by {nm}[{bad.shape[ax]}]: " + " ".join(f"{v:.2f}" for v in fr[:32].tolist()), flush=True)


def fwd_case(kind, cin, cout, k, s, p, shape, dgrad=False):
    g = torch.Generator().manual_seed(3)
    B, D, H, W = shape
    spec = ops.ConvSpec(kind, cin, cout, k, s, p)
    wshape = (cout, cin, k, k, k) if kind == "conv" else (cin, cout, k, k, k)
    w = (torch.randn(wshape, generator=g) * 0.05).cuda()
    if not dgrad:
        x = torch.randn(B, D, H, W, cin, generator=g).to(torch.bfloat16).cuda()
        wp = ops.pack_weight(w, transpose=(kind == "deconv"), dtype=torch.bfloat16)
        f = lambda: ops.conv_forward(spec, x, wp, None, None, False)
    else:
        od = spec.out_dhw((D, H, W))
        x = torch.randn(B, *od, cout, generator=g).to(torch.bfloat16).cuda()
        wp = ops.pack_weight(w, transpose=(kind == "conv"), dtype=torch.bfloat16)
        f = lambda: ops.conv_dgrad(spec, x, wp, (D, H, W))
    y = f(); path = ops.last_path()
    torch.cuda.synchronize()
    ops.set_force_simt(True)
    ys = f()
    ops.set_force_simt(False)
    torch.cuda.synchronize()
    tag = f"{'dgrad' if dgrad else 'fwd'} {kind} cin={cin} cout={cout} k={k} s={s} shape={shape} path={path}"
    bad = report(tag, y, ys)
    if bad.any():
        structure(bad, ["b", "d", "h", "w", "c"])


def wgrad_case(kind, cin, cout, k, s, p, shape):
    g = torch.Generator().manual_seed(4)
    B, D, H, W = shape
    spec = ops.ConvSpec(kind, cin, cout, k, s, p)
    od = spec.out_dhw((D, H, W))
    x = torch.randn(B, D, H, W, cin, generator=g).to(torch.bfloat16).cuda()
    gy = torch.randn(B, *od, cout, generator=g).to(torch.bfloat16).cuda()
    wshape = (cout, cin, k, k, k) if kind == "conv" else (cin, cout, k, k, k)
    w = torch.zeros(wshape).cuda()
    dw = ops.conv_wgrad(spec, x, gy, w); path = ops.last_path()
    torch.cuda.synchronize()
    ops.set_force_simt(True)
    dws = ops.conv_wgrad(spec, x, gy, w)
    ops.set_force_simt(False)
    torch.cuda.synchronize()
    bad = report(f"wgrad {kind} cin={cin} cout={cout} k={k} s={s} shape={shape} path={path}"
                 , dw, dws)
    if bad.any():
        structure(bad.reshape(wshape[0], wshape[1], -1), ["a", "b", "tap"])


FWD = [
    ("conv", 64, 64, 1, 1, 0, (1, 4, 8, 8)),
    ("conv", 64, 16, 1, 1, 0, (1, 1, 8, 16)),
    ("conv", 128, 128, 1, 1, 0, (1, 2, 8, 8)),
    ("conv", 128, 128, 3, 1, 1, (1, 8, 8, 16)),
    ("conv", 128, 128, 3, 1, 1, (2, 5, 7, 10)),
    ("conv", 256, 256, 3, 1, 1, (1, 4, 6, 10)),
    ("conv", 256, 32, 3, 1, 1, (2, 5, 7, 5)),
    ("conv", 128, 128, 4, 2, 1, (1, 8, 16, 16)),
    ("conv", 128, 256, 4, 2, 1, (2, 8, 12, 20)),
    ("deconv", 256, 128, 4, 2, 1, (1, 4, 6, 10)),
    ("deconv", 128, 128, 4, 2, 1, (2, 4, 8, 8)),
]
WG = [
    ("conv", 128, 128, 1, 1, 0, (1, 4, 8, 16)),
    ("conv", 128, 128, 3, 1, 1, (2, 6, 8, 16)),
    ("conv", 64, 128, 3, 1, 1, (1, 4, 8, 8)),
    ("conv", 256, 256, 3, 1, 1, (1, 4, 4, 8)),
    ("conv", 128, 128, 4, 2, 1, (1, 8, 8, 16)),
    ("deconv", 128, 128, 4, 2, 1, (1, 4, 4, 8)),
]

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), flush=True)
    if what in ("fwd", "all"):
        for c in FWD:
            try:
                fwd_case(*c)
                fwd_case(*c, dgrad=True)
            except Exception as e:  # keep going: one broken shape must not hide the others
                print("EXC", c, repr(e), flush=True)
    if what in ("wgrad", "all"):
        for c in WG:
            try:
                wgrad_case(*c)
            except Exception as e:
                print("EXC", c, repr(e), flush=True)
