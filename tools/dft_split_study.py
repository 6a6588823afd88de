"""Numerics study for the planned spectral-loss kernel: a 3-D orthonormal DFT computed as
dense DFT-matrix products whose operands are split into bf16 terms (x = hi + lo [+ lo2]) and multiplied with fp32
accumulation, i.e. what tcgen05 kind::f16 MMAs would do.  CPU only (torch emulation); compares the amplitude and the
spectral loss with torch.fft in float64.  Result at 40 x 56 x 40 (the radix structure of 160 x 224 x 160):
hi*hi only: amplitude error 0.6 of the rms amplitude (useless); hi*hi + hi*lo + lo*hi (3 MMAs): 6.6e-4 of the rms amplitude,
spectral loss within 5e-7 relative; 6 MMAs: 2.9e-5 / 9e-8."""
import math

import torch


def dft_mats(n):
    k = torch.arange(n, dtype=torch.float64)
    ang = -2 * math.pi * torch.outer(k, k) / n
    return (torch.cos(ang) / math.sqrt(n)).float(), (torch.sin(ang) / math.sqrt(n)).float()


def split(x, terms):
    parts, r = [], x.clone()
    for _ in range(terms):
        p = r.to(torch.bfloat16).to(torch.float32)
        parts.append(p)
        r = r - p
    return parts


def mm_split(a_parts, b_parts, terms):
    out = 0
    for i, a in enumerate(a_parts):
        for j, b in enumerate(b_parts):
            if i + j < terms:
                out = out + a @ b
    return out


def dft_last_axis(re, im, terms):
    c, s = dft_mats(re.shape[-1])
    cp, sp, rp = split(c, terms), split(s, terms), split(re, terms)
    if im is None:
        return mm_split(rp, cp, terms), mm_split(rp, sp, terms)
    ip = split(im, terms)
    return mm_split(rp, cp, terms) - mm_split(ip, sp, terms), mm_split(rp, sp, terms) + mm_split(ip, cp, terms)


def amplitude(x, terms):
    re, im = dft_last_axis(x, None, terms)
    for perm in ((0, 1, 3, 2), (0, 3, 2, 1)):            # bring the next axis last, transform, (order restored at the end)
        re, im = re.permute(perm).contiguous(), im.permute(perm).contiguous()
        re, im = dft_last_axis(re, im, terms)
    re, im = re.permute(0, 3, 1, 2), im.permute(0, 3, 1, 2)          # (b, h, w, d) back to (b, d, h, w)
    return torch.sqrt(re ** 2 + im ** 2)


if __name__ == "__main__":
    torch.manual_seed(0)
    x = torch.rand(1, 40, 56, 40)
    y = x + 0.1 * torch.randn_like(x)
    ref_x = torch.fft.fftn(x.double(), dim=(1, 2, 3), norm="ortho").abs()
    ref_y = torch.fft.fftn(y.double(), dim=(1, 2, 3), norm="ortho").abs()
    true = ((ref_x - ref_y) ** 2).mean()
    for terms in (1, 2, 3):
        ax, ay = amplitude(x, terms), amplitude(y, terms)
        err = (ax.double() - ref_x).abs().max()
        got = ((ax - ay) ** 2).mean()
        print(f"split order {terms} ({terms * (terms + 1) // 2} MMAs per product): amplitude max error {err:.3e} "
              f"({err / ref_x.pow(2).mean().sqrt():.2e} of rms), spectral loss rel. error {abs(got - true) / true:.2e}")
