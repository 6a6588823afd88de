"""Sampling latency at the BASELINE.json Performer config: recurrent-state decoder (ms per token) vs the reference-style
loop of full forwards over the growing prefix (measured at a few prefix lengths, integrated over the sequence)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthanatomy_b200.networks.transformers import Ordering, Performer

grid, B = (20, 28, 25), int(os.environ.get("B", "6"))
n = int(np.prod(grid))
order = Ordering("raster_scan", 3, (1, *grid), (False,) * 3, ((2, 0, 1),), ((0, 1),), ("rotate_90", "transpose"))
net = Performer(num_tokens=2049, dim=512, heads=16, dim_head=64, local_attn_heads=8, local_window_size=420,
                max_seq_len=n + 1, depth=24, ordering=order, causal=True, feature_redraw_interval=1,
                generalized_attention=False, use_rezero=True, spatial_position_emb="absolute", spatial_shape=grid,
                compute_dtype=torch.bfloat16).cuda().eval()
tok = torch.randint(0, 2048, (B, n), device="cuda")
dec = net.make_decoder(B, n)
steps = int(os.environ.get("STEPS", "400"))
with torch.no_grad():
    for t in range(20):
        dec.step(tok[:, t], t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(20, 20 + steps):
        dec.step(tok[:, t], t)
    torch.cuda.synchronize()
    ms_tok = (time.perf_counter() - t0) / steps * 1e3
    print(f"recurrent decoder: {ms_tok:.3f} ms per position (batch {B}) -> {ms_tok * n / 1e3:.1f} s per {n}-token sample")
    tot = 0.0
    pts = []
    for L in (1000, 4000, 8000, 14000):
        x = tok[:, :L]
        net(x); torch.cuda.synchronize()
        t0 = time.perf_counter(); net(x); torch.cuda.synchronize()
        pts.append((L, (time.perf_counter() - t0) * 1e3))
    print("full forward over a prefix (ms):", pts)
    # integrate the piecewise-linear cost over prefix lengths 1..n
    Ls = [0] + [p[0] for p in pts]; Ts = [0.0] + [p[1] for p in pts]
    total = float(np.trapz(np.interp(np.arange(1, n + 1), Ls, Ts))) / 1e3
    print(f"reference-style loop (one full forward per token): ~{total:.0f} s per sample -> speed-up {total / (ms_tok * n / 1e3):.0f}x")
