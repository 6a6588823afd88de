"""Local-window attention micro-benchmark at the Performer shape (CUDA events): forward and backward."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthanatomy_b200 import pf_ops as pf


def timeit(f, n=10, warm=3):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


B, N, H, d, W = 6, 14000, 8, 64, 420
g = torch.Generator(device="cuda").manual_seed(0)
buf = (torch.randn(B * N, 3 * H * d, device="cuda", generator=g) * 0.5).bfloat16()
dbuf = torch.empty_like(buf)
out = torch.empty(B * N, H * d, device="cuda", dtype=torch.bfloat16)
dout = torch.randn(B * N, H * d, device="cuda", generator=g).bfloat16()
lse = torch.empty(B * H, N, device="cuda")
desc = pf.local_desc(B, N, H, d, W, 3 * H * d, H * d, torch.bfloat16)
tf = timeit(lambda: pf.local_attn_fwd(desc, buf, 0, H * d, 2 * H * d, None, out, 0, lse))
tb = timeit(lambda: pf.local_attn_bwd(desc, buf, 0, H * d, 2 * H * d, None, out, dout, 0, lse, dbuf))
print(f"local attention B={B} N={N} H={H} W={W} fast={os.environ.get('SA_LOCAL_FASTMASK', '1')}: fwd {tf:.3f} ms  bwd {tb:.3f} ms", flush=True)
