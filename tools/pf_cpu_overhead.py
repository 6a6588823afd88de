"""How long does the host take to enqueue one Performer training step (vs the device time)?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthanatomy_b200.losses import CELoss
from synthanatomy_b200.networks.transformers import Ordering, Performer
from synthanatomy_b200.optim import Adam
from synthanatomy_b200.utils.transformer import prepare_batch

grid, B, depth = (20, 28, 25), 6, int(os.environ.get("DEPTH", "24"))
n = int(np.prod(grid))
dev = torch.device("cuda", 0)
order = Ordering("raster_scan", 3, (1, *grid), (False,) * 3, ((2, 0, 1),), ((0, 1),), ("rotate_90", "transpose"))
net = Performer(num_tokens=2049, dim=512, heads=16, dim_head=64, local_attn_heads=8, local_window_size=420,
                max_seq_len=n + 1, depth=depth, ordering=order, causal=True, feature_redraw_interval=1,
                generalized_attention=False, use_rezero=True, spatial_position_emb="absolute", spatial_shape=grid,
                compute_dtype=torch.bfloat16).to(dev).train()
opt = Adam(net.parameters(), lr=1e-3)
crit = CELoss()
quant = torch.randint(0, 2048, (B, *grid))
(x, _), y = prepare_batch({"quantization": quant}, order.get_sequence_ordering(), 2048)
x, y = x.to(dev), y.to(dev)

def step():
    logits = net(x)
    loss = crit(logits.transpose(1, 2), y)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)

for _ in range(3):
    step()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host enqueue {1e3 * (t1 - t0):8.1f} ms   device tail {1e3 * (t2 - t1):8.1f} ms   total {1e3 * (t2 - t0):8.1f} ms")
