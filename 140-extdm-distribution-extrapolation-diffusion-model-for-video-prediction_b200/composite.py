"""Algebra of the composite init_conv (extdm_b200/unet.py: UnetRunner._composite_init).

init_conv(init_noise_conv(x)) is two linear 7x7 convolutions in a row (model/BaseDM_adaptor/
DenoiseNet_STWAtt_w_w_ref_adaptor_cross_multi_traj_ada.py:916,1032-1042); with xn = b1 + w1 * pad3(x) zero padded before
the second window,

    w2 * pad3(xn)  =  w2 * xn_ext  -  w2 * ring,        xn_ext = b1 + w1 * pad6(x) on the image extended by 3 pixels,

`ring` = xn_ext outside the image.  The first term is ONE 13x13 convolution of the 3-channel x; the second is linear in x
as well and, summed over all ring positions of a side, position independent.  `compose` returns every weight block the
runner's GEMMs and the corner kernel use, in float64 and in the runner's K layouts (all over the x-direction im2col
tensor Xc[y, x, (dx + 6)*3 + c] = x[y, x + dx, c] (zero outside the image), Xc[y, x, 39] = 1):

    comp    (Cout, 13, 64)      out[y, x] += sum_dy comp[:, dy + 6] . Xc[y + dy, x]           (+ comp_bias = w2 . b1)
    top     (3, Cout, 3, 64)    out[p, x] += sum_s top[p][:, s] . Xc[s, x]                     p = 0, 1, 2 (already negated)
    bottom  (3, Cout, 3, 64)    out[H - 3 + p, x] += sum_s bottom[p][:, s] . Xc[H - 3 + s, x]
    left    (3, Cout, 13, 64)   out[y, p] += sum_dy left[p][:, dy + 6] . Xc[y + dy, p]
    right   (3, Cout, 13, 64)   out[y, W - 3 + p] += sum_dy right[p][:, dy + 6] . Xc[y + dy, W - 3 + p]
    corners (4, 9, 28, Cout)    out[y0 + py, x0 + px] += corners[cn, py*3 + px, :27] . x[y0:y0+3, x0:x0+3, :] + corners[.., 27]
                                (cn = top-left, top-right, bottom-left, bottom-right: the 3x3 corner blocks of the ring belong
                                to a row side and a column side, i.e. the sides subtract them twice)

tests/test_host_cpu.py::test_composite_init_conv_algebra checks the decomposition against the two-stage convolution on the CPU.
"""
import torch


def _outside(side, pr):
    """Kernel rows / columns of the second convolution that reach outside the image for output row / column `pr`."""
    return range(-3, -pr) if side in ("top", "left") else range(3 - pr, 4)


def compose(w1, b1, w2):
    """w1 (M, 3, 7, 7), b1 (M,): the first convolution; w2 (Cout, M, 7, 7): the second one's weights over its M channels."""
    w1, b1, w2 = w1.double(), b1.double(), w2.double()
    co, dev = w2.shape[0], w2.device
    z = lambda *shape: torch.zeros(*shape, dtype=torch.float64, device=dev)
    w12 = z(co, 3, 13, 13)
    for a in range(7):                       # out(p) = sum_a w2[a] xn(p + a), xn(q) = sum_a' w1[a'] x(q + a')
        for b in range(7):
            w12[:, :, a:a + 7, b:b + 7] += torch.einsum("om,mcij->ocij", w2[:, :, a, b], w1)
    comp = z(co, 13, 64)
    comp[:, :, :39] = w12.permute(0, 2, 3, 1).reshape(co, 13, 39)
    out = {"comp": comp, "comp_bias": torch.einsum("omab,m->o", w2, b1)}
    for side in ("top", "bottom"):           # source row s = p + ky + a in {0, 1, 2}
        mats = []
        for pr in range(3):
            cp = z(co, 3, 64)
            for ky in _outside(side, pr):
                for kx in range(-3, 4):
                    w2t = w2[:, :, ky + 3, kx + 3]
                    cp[:, 0, 39] += w2t @ b1                                           # constant carrier: a pixel of the first source row
                    for sr in range(3):
                        a = sr - pr - ky
                        if -3 <= a <= 3:
                            blk = torch.einsum("om,mcb->obc", w2t, w1[:, :, a + 3, :])  # (Cout, 7 (b), 3 (c))
                            lo = (kx + 3) * 3                                          # dx + 6 = kx + b + 6
                            cp[:, sr, lo:lo + 21] += blk.reshape(co, 21)
            mats.append(-cp)
        out[side] = torch.stack(mats)
    for side in ("left", "right"):           # all 13 row offsets dy = ky + a, read at the output column
        mats = []
        for pr in range(3):
            cp = z(co, 13, 64)
            for kx in _outside(side, pr):
                for ky in range(-3, 4):
                    w2t = w2[:, :, ky + 3, kx + 3]
                    cp[:, 6, 39] += w2t @ b1
                    blk = torch.einsum("om,mcab->oabc", w2t, w1)                       # (Cout, 7 (a), 7 (b), 3)
                    lo = (kx + 3) * 3
                    cp[:, ky + 3:ky + 10, lo:lo + 21] += blk.reshape(co, 7, 21)
            mats.append(-cp)
        out[side] = torch.stack(mats)
    tab = z(4, 9, 28, co)
    for cn, (vs, hs) in enumerate((("top", "left"), ("top", "right"), ("bottom", "left"), ("bottom", "right"))):
        for py in range(3):
            for px in range(3):
                for ky in _outside(vs, py):
                    for kx in _outside(hs, px):
                        w2t = w2[:, :, ky + 3, kx + 3]
                        tab[cn, py * 3 + px, 27] += w2t @ b1
                        for r in range(3):
                            for sc in range(3):
                                a, b = r - py - ky, sc - px - kx
                                if -3 <= a <= 3 and -3 <= b <= 3:
                                    k0 = (r * 3 + sc) * 3
                                    tab[cn, py * 3 + px, k0:k0 + 3] += (w2t @ w1[:, :, a + 3, b + 3]).t()
    out["corners"] = tab
    return out


def _bilinear2_taps(phase):
    """B[k, u]: weight of low-resolution index i + (u - 2) in the x2 bilinear (align_corners=False) up-sampling at the
    high-resolution position 2 i + phase + (k - 3), k = 0 .. 6 (ATen area_pixel_compute_source_index: s = (P + 0.5) / 2 - 0.5;
    the border clamping equals replicate padding of the low-resolution row)."""
    b = torch.zeros(7, 5, dtype=torch.float64)
    for k in range(7):
        e = phase + k - 3
        j = e // 2
        if e % 2 == 0:                       # s = i + j - 0.25
            b[k, j - 1 + 2] += 0.25
            b[k, j + 2] += 0.75
        else:                                # s = i + j + 0.25
            b[k, j + 2] += 0.75
            b[k, j + 1 + 2] += 0.25
    return b


def compose_upsampled(w2):
    """7x7 zero-padded convolution (weights w2 (Cout, M, 7, 7)) of a x2 bilinearly up-sampled tensor, evaluated on the
    LOW-resolution tensor f (TrajWarp branch of the u12 UNet, ..._traj_u12.py:1017-1042: F.interpolate then init_conv):

        w2 * pad0(up(f))  =  w2 * rep3(up(f))  -  w2 * (rep3(up(f)) outside the image)

    The first term is a 5x5 convolution of the replicate-padded f per output phase (py, px) = output pixel parity
    (49 -> 25 taps); outside the image rep3(up(f)) repeats the border rows / columns of up(f), so the correction of an output
    row / column next to the border is a 7-tap 1-D convolution of that border row / column.

        poly   (4, Cout, 5, 5, M)  out[2i + py, 2j + px] = sum_uv poly[py*2 + px][:, u, v] . fpad[i + u, j + v]   (fpad = rep2(f))
        top    (3, Cout, 7, M)     out[p, x] += sum_kx top[p][:, kx + 3] . T[x + kx + 3],  T[x'] = up(f)[0, clamp(x' - 3)]  (negated)
        bottom (3, Cout, 7, M)     out[H - 3 + p, x] += sum_kx bottom[p][:, kx + 3] . Bo[x + kx + 3],  Bo = last row likewise
        left   (3, Cout, 7, M)     out[y, p] += sum_ky left[p][:, ky + 3] . L[y + ky]  (L[y] = up(f)[y, 0], zero outside 0 .. H-1)
        right  (3, Cout, 7, M)     out[y, W - 3 + p] += sum_ky right[p][:, ky + 3] . R[y + ky]  (R[y] = up(f)[y, W - 1])
    (the corner positions of the ring are in the top / bottom rows only: the columns stop at the image's rows)"""
    w2 = w2.double()
    co, m = w2.shape[:2]
    poly = []
    for py in range(2):
        for px in range(2):
            by, bx = _bilinear2_taps(py).to(w2.device), _bilinear2_taps(px).to(w2.device)
            poly.append(torch.einsum("omyx,yu,xv->ouvm", w2, by, bx))
    out = {"poly": torch.stack(poly)}
    out["top"] = torch.stack([-w2[:, :, 0:3 - p, :].sum(2).permute(0, 2, 1) for p in range(3)])          # ky <= -1 - p
    out["bottom"] = torch.stack([-w2[:, :, 6 - p:7, :].sum(2).permute(0, 2, 1) for p in range(3)])       # ky >= 3 - p
    out["left"] = torch.stack([-w2[:, :, :, 0:3 - p].sum(3).permute(0, 2, 1) for p in range(3)])         # kx <= -1 - p
    out["right"] = torch.stack([-w2[:, :, :, 6 - p:7].sum(3).permute(0, 2, 1) for p in range(3)])        # kx >= 3 - p
    return out
