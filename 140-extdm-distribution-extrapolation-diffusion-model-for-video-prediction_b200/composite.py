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
