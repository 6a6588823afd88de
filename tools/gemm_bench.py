"""GEMM micro-benchmark at the Performer shapes (CUDA events)."""
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

M = 84000
for (n, k, what) in [(3072, 512, "qkv"), (512, 1024, "out"), (2048, 512, "w1"), (512, 2048, "w2")]:
    a = torch.randn(M, k, device="cuda").bfloat16(); b = (torch.randn(n, k, device="cuda") * 0.05).bfloat16()
    o = torch.empty(M, n, device="cuda", dtype=torch.bfloat16)
    t = timeit(lambda: pf.gemm_nt(a, b, out_act=o))
    tt = timeit(lambda: torch.matmul(a, b.t(), out=o))
    print(f"{os.environ.get('SA_GEMM_NOEPI','-')} {what:4s} M={M} N={n} K={k}: {t:.3f} ms {2*M*n*k/t/1e9:7.1f} TF   cublas {tt:.3f} ms {2*M*n*k/tt/1e9:7.1f} TF", flush=True)
