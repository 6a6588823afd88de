import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_performer as T

base = dict(num_tokens=65, dim=128, depth=2, heads=4, dim_head=64, local_attn_heads=2, local_window_size=20)
variants = {
    "tiny": base,
    "depth1": dict(base, depth=1),
    "nolocal": dict(base, local_attn_heads=0),
    "alllocal": dict(base, local_attn_heads=4),
    "dim64": dict(base, dim=64),
    "heads2": dict(base, heads=2, local_attn_heads=1),
}
for name, kw in variants.items():
    cfg, sd, net, seqs, x_in, y = T._build(kw, (4, 5, 6), 21)
    net = net.cuda().eval(); x = x_in.cuda(); n = x.shape[1]
    dec = net.make_decoder(x.shape[0], n)
    errs = []
    with torch.no_grad():
        for t in range(8):
            lg = dec.step(x[:, t], t)
            ref = net(x[:, :t + 1])[:, -1]
            errs.append(float((lg - ref).abs().max()))
    print(name, ["%.1e" % e for e in errs])
