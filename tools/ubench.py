"""GPU micro-benchmarks of tcgen05.mma rate and TMA box-load rate (tuning aid; see csrc/sa_ubench.cu)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthanatomy_b200 import _lib  # noqa: E402

lib = _lib.load()
out = torch.zeros(4, dtype=torch.int64, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
vp = lambda t: C.c_void_p(t.data_ptr())

print("== tcgen05.mma: cycles per 128xNx16 MMA (grid 148, 2000 x 4 MMAs)")
for a_mn, b_mn in ((0, 0), (1, 1), (0, 1), (1, 0)):
    for N in (64, 128, 256):
        for dep in (0, 1):
            iters = 2000
            rc = lib.sa_ubench_mma(a_mn, b_mn, N, iters, dep, 148, vp(out), st)
            assert rc == 0, lib.sa_last_error()
            torch.cuda.synchronize()
            cyc = out[0].item() / (iters * 4)
            print(f"a_mn={a_mn} b_mn={b_mn} N={N:3d} commit_each_stage={dep}: {cyc:7.1f} cycles/MMA "
                  f"(ideal {128 * N / 256:.0f})", flush=True)

print("== TMA: single producer thread per CTA, consumer frees slots immediately")
x = torch.randn(1, 80, 112, 80, 128, device="cuda").to(torch.bfloat16)
for (tw, th, td, lps, stages, grid) in [
        (16, 8, 1, 1, 6, 148), (16, 8, 1, 2, 6, 148), (16, 8, 1, 1, 6, 444),
        (8, 4, 1, 1, 8, 148), (8, 4, 1, 4, 8, 148), (8, 4, 1, 10, 5, 148),
        (16, 8, 2, 1, 4, 148), (16, 8, 2, 2, 3, 148), (16, 4, 1, 2, 8, 148)]:
    iters = 2000
    rc = lib.sa_ubench_tma(vp(x), 128, 80, 112, 80, tw, th, td, lps, stages, iters, grid, vp(out), st)
    assert rc == 0, lib.sa_last_error()
    torch.cuda.synchronize()
    tot, iss = out[0].item(), out[1].item()
    box = tw * th * td * 128
    print(f"box {tw}x{th}x{td} ({box // 1024:3d} KB) loads/stage={lps:2d} stages={stages} grid={grid}: "
          f"{tot / iters:8.1f} cyc/stage, {iss / (iters * lps):6.1f} cyc/TMA issue, "
          f"{lps * box * iters / tot:6.1f} B/cyc/CTA", flush=True)
