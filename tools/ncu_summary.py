#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one line per launch with the metrics the roofline needs."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def g(r, name, default=""):
    i = col.get(name)
    return r[i] if i is not None else default
print(f"{'kernel':44s} {'grid':>7s} {'ms':>8s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>7s} {'tens%':>6s} {'warps%':>6s} {'regs':>4s} {'smemKB':>6s}")
for r in rows[2:]:
    name = g(r, "Kernel Name").replace("<unnamed>::", "").replace("void ", "")[:44]
    def f(n):
        try: return float(g(r, n).replace(",", ""))
        except Exception: return float("nan")
    def scaled(n):
        v = f(n); u = units[col[n]] if n in col else ""
        mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
        return v * mul
    dur = f("gpu__time_duration.sum"); du = units[col["gpu__time_duration.sum"]]
    ms = dur * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(du, 1)
    rd, wr = scaled("dram__bytes_read.sum"), scaled("dram__bytes_write.sum")
    print(f"{name:44s} {g(r,'launch__grid_size'):>7s} {ms:8.3f} {rd/1e6:8.1f} {wr/1e6:8.1f} {(rd+wr)/ms/1e6:7.0f} "
          f"{f('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{f('sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} {g(r,'launch__registers_per_thread'):>4s} "
          f"{scaled('launch__shared_mem_per_block_dynamic')/1024 if 'launch__shared_mem_per_block_dynamic' in col else float('nan'):6.0f}")
