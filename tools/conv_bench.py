"""GPU micro-benchmark of the conv primitives at the level-1 shapes of the headline config (events, L2-exceeding
tensors).  Usage: python tools/conv_bench.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthanatomy_b200 import ops  # noqa: E402


def timeit(f, n=5, warm=2):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    D, H, W, C = 80, 112, 80, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, D, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(B, D, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    w3 = torch.randn(C, C, 3, 3, 3, device="cuda", generator=g) * 0.02
    w1 = torch.randn(C, C, 1, 1, 1, device="cuda", generator=g) * 0.02
    bias = torch.randn(C, device="cuda", generator=g)
    s3, s1 = ops.ConvSpec("conv", C, C, 3, 1, 1), ops.ConvSpec("conv", C, C, 1, 1, 0)
    wp3, wp1 = ops.pack_weight(w3, False, torch.bfloat16), ops.pack_weight(w1, False, torch.bfloat16)
    pos = B * D * H * W
    f3 = 2.0 * pos * 27 * C * C
    f1 = 2.0 * pos * C * C
    tag = f"B={B} stages_env={os.environ.get('SA_TC_MAX_STAGES', '-')}"
    t = timeit(lambda: ops.conv_forward(s3, x, wp3, bias, None, True))
    print(f"{tag} fwd3x3x3 relu          {t:8.3f} ms  {f3 / t / 1e9:8.1f} TFLOP/s", flush=True)
    t = timeit(lambda: ops.conv_forward(s3, x, wp3, None, dy, False, x))
    print(f"{tag} dgrad3-like add+mask   {t:8.3f} ms  {f3 / t / 1e9:8.1f} TFLOP/s", flush=True)
    t = timeit(lambda: ops.conv_forward(s1, x, wp1, bias, dy, True))
    print(f"{tag} fwd1x1x1 add+relu      {t:8.3f} ms  {f1 / t / 1e9:8.1f} TFLOP/s  {3 * pos * C * 2 / t / 1e6:8.1f} GB/s",
          flush=True)
    t = timeit(lambda: ops.conv_wgrad(s3, x, dy, w3))
    print(f"{tag} wgrad3x3x3             {t:8.3f} ms  {f3 / t / 1e9:8.1f} TFLOP/s", flush=True)
    t = timeit(lambda: ops.conv_wgrad(s1, x, dy, w1))
    print(f"{tag} wgrad1x1x1             {t:8.3f} ms  {f1 / t / 1e9:8.1f} TFLOP/s", flush=True)
    wp1t = ops.pack_weight(w1, True, torch.bfloat16)
    t = timeit(lambda: ops.conv1x1_bwd_fused(s1, dy, x, wp1t, w1))
    print(f"{tag} 1x1x1 fused bwd        {t:8.3f} ms  {2 * f1 / t / 1e9:8.1f} TFLOP/s  {3 * pos * C * 2 / t / 1e6:8.1f} GB/s", flush=True)
    t = timeit(lambda: ops.bias_grad(dy))
    print(f"{tag} bias_grad              {t:8.3f} ms  {pos * C * 2 / t / 1e6:8.1f} GB/s", flush=True)
    # strided pair
    sd = ops.ConvSpec("conv", C, C, 4, 2, 1)
    wd = torch.randn(C, C, 4, 4, 4, device="cuda", generator=g) * 0.02
    wpd = ops.pack_weight(wd, False, torch.bfloat16)
    fd = 2.0 * (pos / 8) * 64 * C * C
    t = timeit(lambda: ops.conv_forward(sd, x, wpd, bias, None, True))
    print(f"{tag} fwd k4s2               {t:8.3f} ms  {fd / t / 1e9:8.1f} TFLOP/s", flush=True)
    ys = ops.conv_forward(sd, x, wpd, bias, None, True)
    wpdt = ops.pack_weight(wd, True, torch.bfloat16)
    t = timeit(lambda: ops.conv_dgrad(sd, ys, wpdt, (D, H, W), None, x))
    print(f"{tag} dgrad k4s2 (8 phases)  {t:8.3f} ms  {fd / t / 1e9:8.1f} TFLOP/s", flush=True)
    t = timeit(lambda: ops.conv_wgrad(sd, x, ys, wd))
    print(f"{tag} wgrad k4s2             {t:8.3f} ms  {fd / t / 1e9:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
