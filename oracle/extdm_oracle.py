"""ORACLE -- test infrastructure only, never the product.

A CPU (fp32, plain torch.nn.functional) restatement of the reference's sampling hot path:

    Unet3D.forward            model/BaseDM_adaptor/DenoiseNet_STWAtt_w_w_ref_adaptor_cross_multi_traj_ada.py:1020-1089
                              .../DenoiseNet_STWAtt_w_w_ref_adaptor_cross_multi_traj_u12.py:1017-1084
                              .../DenoiseNet_STWAtt_w_wo_ref_adaptor_cross_multi.py:906-966
    GaussianDiffusion.ddim_sample   model/BaseDM_adaptor/Diffusion.py:208-258
    Generator.forward_with_flow     model/LFAE/generator.py:152-206

Every function works on a *flat state_dict* with the reference's key names, so the same perturbed
weights drive the reference, this oracle and the CUDA path.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product package never does.

Pinning: the reference ships no tests or golden vectors for this path.  The oracle is
pinned against outputs of the reference itself, generated in the build container by
tests/golden/make_golden.py and committed under tests/golden/*.pt (tests/test_oracle_golden.py).
One third-party piece stays "parity unpinned": rotary-embedding-torch==0.8.3 is absent from the image,
its rotation is restated here (rotary()) from its published algorithm (interleaved pairs, theta=10000).
"""
import math
from functools import lru_cache

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- state-dict view
class SD:
    """Prefix view over a flat state_dict."""

    def __init__(self, sd, prefix=""):
        self.sd, self.prefix = sd, prefix

    def __getitem__(self, k):
        return self.sd[self.prefix + k]

    def has(self, k):
        return (self.prefix + k) in self.sd

    def sub(self, p):
        return SD(self.sd, self.prefix + p + ".")


# ----------------------------------------------------------------------------- UNet config
def unet_config(variant, tc, tp, dim=64, dim_mults=(1, 2, 4, 4)):
    """variant: 'ada' (KTH/UCF/City), 'u12' (BAIR), 'base' (SMMNIST), 'u22' (ada_u22, the shipped Cityscapes
    pairing).  App. A of SURVEY.md."""
    if variant == "ada":
        window, dim_head = (4, 4, 4), 16          # ..._traj_ada.py:872-877
    elif variant == "u22":
        window, dim_head = (4, 4, 4), 32          # ..._traj_ada_u22.py:1016-1021
    elif variant in ("u12", "base"):
        window, dim_head = (2, 4, 4), 32          # ..._traj_u12.py:871-876, ...cross_multi.py:762-767
    else:
        raise ValueError(variant)
    tm = tc - 1 if variant == "base" else tc      # ...cross_multi.py:699 vs ..._traj_ada.py:699
    return dict(variant=variant, tc=tc, tp=tp, tm=tm, T=tm + tp, dim=dim, dim_mults=tuple(dim_mults),
                window=window, heads=8, dim_head=dim_head, groups=8)


def adaptor_layers(tm, tp):
    """compute_layer, ..._traj_ada.py:644-649 (l=None)."""
    L = max(1, int(math.ceil(math.log2((tp + 1) / tm))))
    return L, (2 ** L - 1) * tm


# ----------------------------------------------------------------------------- small pieces
def rotary(t):
    """rotary-embedding-torch 0.8.3 rotate_queries_or_keys over dim -2 (see module docstring)."""
    n, d = t.shape[-2], t.shape[-1]
    freqs = 1.0 / (10000.0 ** (torch.arange(0, d, 2, dtype=torch.float32)[: d // 2] / d))
    ang = torch.arange(n, dtype=torch.float32)[:, None] * freqs[None, :]
    ang = ang.repeat_interleave(2, dim=-1)
    pair = t.reshape(*t.shape[:-1], d // 2, 2)
    rot = torch.stack((-pair[..., 1], pair[..., 0]), dim=-1).reshape(t.shape)
    return t * ang.cos() + rot * ang.sin()


def t5_bucket_bias(emb_weight, n, num_buckets=32, max_distance=32):
    """RelativePositionBias.forward, ...cross_multi.py:43-80 -> (heads, n, n)."""
    q = torch.arange(n)
    rel = q[None, :] - q[:, None]                  # k_pos - q_pos
    m = -rel
    nb = num_buckets // 2
    ret = (m < 0).long() * nb
    m = m.abs()
    max_exact = nb // 2
    small = m < max_exact
    large = max_exact + (torch.log(m.float() / max_exact) / math.log(max_distance / max_exact)
                         * (nb - max_exact)).long()
    large = torch.minimum(large, torch.full_like(large, nb - 1))
    bucket = ret + torch.where(small, m, large)
    return emb_weight[bucket].permute(2, 0, 1)


def chan_layernorm(x, gamma, eps=1e-5):
    """LayerNorm over dim 1 of a 5-D tensor, gamma only, biased variance. ...cross_multi.py:139-148."""
    var = x.var(dim=1, unbiased=False, keepdim=True)
    mean = x.mean(dim=1, keepdim=True)
    return (x - mean) / (var + eps).sqrt() * gamma


def block(x, sd, groups, scale_shift=None):
    """Block: conv(1,3,3) -> GroupNorm -> (scale+1, shift) -> SiLU.  ...cross_multi.py:163-178."""
    x = F.conv3d(x, sd["proj.weight"], sd["proj.bias"], padding=(0, 1, 1))
    x = F.group_norm(x, groups, sd["norm.weight"], sd["norm.bias"], eps=1e-5)
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    return F.silu(x)


def resnet_block(x, sd, groups, t_emb=None):
    """ResnetBlock, ...cross_multi.py:182-204."""
    ss = None
    if sd.has("mlp.1.weight"):
        e = F.linear(F.silu(t_emb), sd["mlp.1.weight"], sd["mlp.1.bias"])
        e = e[:, :, None, None, None]
        ss = e.chunk(2, dim=1)
    h = block(x, sd.sub("block1"), groups, ss)
    h = block(h, sd.sub("block2"), groups)
    res = F.conv3d(x, sd["res_conv.weight"], sd["res_conv.bias"]) if sd.has("res_conv.weight") else x
    return h + res


def temporal_attention(x, sd, pos_bias, heads, dim_head):
    """Residual(PreNorm(EinopsToAndFrom(AttentionLayer))): y = x + z + to_out(attn(LN(z))), z = chanLN(x).
    ...cross_multi.py:253-328 (App. B.4 of SURVEY.md)."""
    b, c, t, h, w = x.shape
    z = chan_layernorm(x, sd["fn.norm.gamma"])
    zt = z.permute(0, 3, 4, 2, 1).reshape(b, h * w, t, c)
    a = sd.sub("fn.fn.fn")
    u = F.layer_norm(zt, (c,), a["norm.weight"], a["norm.bias"], eps=1e-5)
    qkv = F.linear(u, a["attn.to_qkv.weight"])
    q, k, v = qkv.chunk(3, dim=-1)

    def split(y):
        return y.reshape(b * h * w, t, heads, dim_head).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    q = q * dim_head ** -0.5
    q, k = rotary(q), rotary(k)
    sim = q @ k.transpose(-1, -2) + pos_bias
    attn = sim.softmax(dim=-1)
    o = (attn @ v).permute(0, 2, 1, 3).reshape(b, h * w, t, heads * dim_head)
    o = F.linear(o, a["attn.to_out.weight"])
    y = zt + o
    y = y.reshape(b, h, w, t, c).permute(0, 4, 3, 1, 2)
    return y + x


def _window_geometry(dims, window, shift):
    ws, ss = list(window), list(shift)
    for i in range(3):
        if dims[i] <= window[i]:
            ws[i] = dims[i]
            ss[i] = 0
    return tuple(ws), tuple(ss)


@lru_cache(maxsize=None)
def _shift_mask(Dp, Hp, Wp, ws, ss):
    """compute_mask, ...cross_multi.py:376-389: (nW, N, N) with -100 between different regions."""
    ids = torch.zeros(Dp, Hp, Wp)
    cnt = 0
    for d in (slice(-ws[0]), slice(-ws[0], -ss[0]), slice(-ss[0], None)):
        for h in (slice(-ws[1]), slice(-ws[1], -ss[1]), slice(-ss[1], None)):
            for w in (slice(-ws[2]), slice(-ws[2], -ss[2]), slice(-ss[2], None)):
                ids[d, h, w] = cnt
                cnt += 1
    win = ids.reshape(Dp // ws[0], ws[0], Hp // ws[1], ws[1], Wp // ws[2], ws[2])
    win = win.permute(0, 2, 4, 1, 3, 5).reshape(-1, ws[0] * ws[1] * ws[2])
    diff = win[:, None, :] - win[:, :, None]
    return torch.where(diff != 0, torch.tensor(-100.0), torch.tensor(0.0))


def _rel_pos_index(window):
    """relative_position_index buffer, ...cross_multi.py:437-451 (built for the *configured* window)."""
    wd, wh, ww = window
    coords = torch.stack(torch.meshgrid(torch.arange(wd), torch.arange(wh), torch.arange(ww), indexing="ij"))
    cf = coords.flatten(1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += wd - 1
    rel[:, :, 1] += wh - 1
    rel[:, :, 2] += ww - 1
    rel[:, :, 0] *= (2 * wh - 1) * (2 * ww - 1)
    rel[:, :, 1] *= (2 * ww - 1)
    return rel.sum(-1)


def stw_attention(x, sd, window, shift, heads, dim_head):
    """Residual(PreNorm(STWAttentionLayer)), ...cross_multi.py:409-560 (App. B.5/B.6)."""
    B, C, D, H, W = x.shape
    ws, ss = _window_geometry((D, H, W), window, shift)
    z = chan_layernorm(x, sd["fn.norm.gamma"]).permute(0, 2, 3, 4, 1)      # b d h w c
    Dp = -(-D // ws[0]) * ws[0]
    Hp = -(-H // ws[1]) * ws[1]
    Wp = -(-W // ws[2]) * ws[2]
    z = F.pad(z, (0, 0, 0, Wp - W, 0, Hp - H, 0, Dp - D))
    shifted = any(s > 0 for s in ss)
    if shifted:
        z = torch.roll(z, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
    N = ws[0] * ws[1] * ws[2]
    win = z.reshape(B, Dp // ws[0], ws[0], Hp // ws[1], ws[1], Wp // ws[2], ws[2], C)
    win = win.permute(0, 1, 3, 5, 2, 4, 6, 7).reshape(-1, N, C)
    a = sd.sub("fn.fn.attn")
    qkv = F.linear(win, a["qkv.weight"]).reshape(-1, N, 3, heads, dim_head).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * dim_head ** -0.5
    q, k = rotary(q), rotary(k)
    attn = q @ k.transpose(-1, -2)
    idx = _rel_pos_index(window)[:N, :N].reshape(-1)
    bias = a["relative_position_bias_table"][idx].reshape(N, N, heads).permute(2, 0, 1)
    attn = attn + bias[None]
    if shifted:
        mask = _shift_mask(Dp, Hp, Wp, ws, ss)
        nW = mask.shape[0]
        attn = (attn.reshape(-1, nW, heads, N, N) + mask[None, :, None]).reshape(-1, heads, N, N)
    attn = attn.softmax(dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(-1, N, heads * dim_head)
    o = F.linear(o, a["proj.weight"], a["proj.bias"])                       # (B*nW, N, C)
    o = o.reshape(B, Dp // ws[0], Hp // ws[1], Wp // ws[2], ws[0], ws[1], ws[2], C)
    o = o.permute(0, 1, 4, 2, 5, 3, 6, 7).reshape(B, Dp, Hp, Wp, C)
    if shifted:
        o = torch.roll(o, shifts=ss, dims=(1, 2, 3))
    o = o[:, :D, :H, :W].permute(0, 4, 1, 2, 3)
    return o + x


def motion_adaptor(x, sd, tm, tp):
    """MotionAdaptor + adaptor, ..._traj_ada.py:659-718 (App. B.7)."""
    xm, xp = x[:, :, :tm], x[:, :, tm:]
    C = x.shape[1]
    ad = sd.sub("adaptors")
    y = xm + F.conv3d(chan_layernorm(xm, ad["predictor.fn.norm.gamma"]),
                      ad["predictor.fn.fn.weight"], ad["predictor.fn.fn.bias"])
    L, _ = adaptor_layers(tm, tp)
    cur = y
    for i in range(L):
        flat = cur.reshape(cur.shape[0], C, -1)
        std = (flat.var(dim=2) + 1e-5).sqrt()[:, :, None, None, None]       # unbiased
        mean = flat.mean(dim=2)[:, :, None, None, None]
        nh = (cur - mean) / std
        wx = ad[f"extrapolators.{i}.fn.weight"]   # (1,3,3) kernels; ada_u22 uses 3x3x3, padding 1 (..._ada_u22.py:793)
        nh = nh + F.conv3d(nh, wx, None, padding=(wx.shape[2] // 2, 1, 1))
        cur = torch.cat([cur, nh * std + mean], dim=2)
    ext = cur[:, :, tm:]                                                     # (2^L-1)*tm frames
    n, _, Te, h, w = ext.shape
    flat = ext.permute(0, 2, 1, 3, 4).reshape(n, Te * C, h, w)               # 'N C T H W -> N (T C) H W'
    mod = F.conv2d(flat, sd["Tmodulator.weight"], sd["Tmodulator.bias"])
    mod = mod.reshape(n, tp, C, h, w).permute(0, 2, 1, 3, 4)
    cat = torch.cat([mod, xp], dim=1)
    fused = F.conv3d(chan_layernorm(cat, sd["fuser.norm.gamma"]), sd["fuser.fn.weight"], sd["fuser.fn.bias"])
    return torch.cat([xm, fused + xp], dim=2)


def traj_warp(xp, f, sd, tm, tp, heads=8):
    """TrajWarp + MultiHeadAttentionOp, ..._traj_u12.py:719-827."""
    fm, fp = f[:, :, :tm], f[:, :, tm:]
    n, c = f.shape[:2]
    h, w = fp.shape[3:]
    xp = F.max_pool3d(xp, (1, 2, 2), (1, 2, 2))
    kv = fm.permute(0, 2, 3, 4, 1).reshape(n, -1, c)
    q = xp.permute(0, 2, 3, 4, 1).reshape(n, -1, c)
    ca = sd.sub("cross_att")
    q = F.relu(F.linear(q, ca["linear_q.weight"], ca["linear_q.bias"]))
    k = F.relu(F.linear(kv, ca["linear_k.weight"], ca["linear_k.bias"]))
    v = F.relu(F.linear(kv, ca["linear_v.weight"], ca["linear_v.bias"]))
    dh = c // heads

    def split(y):
        return y.reshape(n, -1, heads, dh).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    att = (q @ k.transpose(-1, -2) / math.sqrt(dh)).softmax(dim=-1)
    y = (att @ v).permute(0, 2, 1, 3).reshape(n, -1, c)
    y = F.relu(F.linear(y, ca["linear_o.weight"], ca["linear_o.bias"]))
    y = y.reshape(n, tp, h, w, c).permute(0, 4, 1, 2, 3)
    fp = F.conv3d(torch.cat([fp, y], dim=1), sd["fuser.weight"], sd["fuser.bias"])
    return torch.cat([fm, fp], dim=2)


def time_embedding(time, sd, dim):
    """SinusoidalPosEmb + time_mlp, ...cross_multi.py:110-122,812-817."""
    half = dim // 2
    f = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    e = time[:, None].float() * f[None, :]
    e = torch.cat((e.sin(), e.cos()), dim=-1)
    e = F.gelu(F.linear(e, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"]))
    return F.linear(e, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])


# ----------------------------------------------------------------------------- UNet forward
def unet_cond_features(sd, cfg, cond_fea, pos_bias, out_hw):
    """Step-invariant part of the 'ada' forward (..._traj_ada.py:1035-1041): cond_adaptor ->
    cond_temporal_attn -> bilinear resize to the flow resolution."""
    cf = motion_adaptor(cond_fea, sd.sub("cond_adaptor"), cfg["tm"], cfg["tp"])
    cf = temporal_attention(cf, sd.sub("cond_temporal_attn"), pos_bias, cfg["heads"], cfg["dim_head"])
    return _resize_frames(cf, out_hw)


def _resize_frames(cf, out_hw):
    n, c, t, h, w = cf.shape
    y = cf.permute(0, 2, 1, 3, 4).reshape(n * t, c, h, w)
    y = F.interpolate(y, size=out_hw, mode="bilinear")
    return y.reshape(n, t, c, *out_hw).permute(0, 2, 1, 3, 4)


def unet_forward(sd, cfg, x, time, cond_frames, cond_fea, taps=None):
    """Unet3D.forward for the working variants ('u22' = ..._traj_ada_u22.py:1172-1310 with path=0).  `sd` is an SD view at the UNet root
    (prefix 'denoise_fn.' inside a diffusion state_dict).  `taps`: optional dict that receives
    intermediate tensors for layer-by-layer parity tests."""
    v, tc, tp, tm = cfg["variant"], cfg["tc"], cfg["tp"], cfg["tm"]
    heads, dh, groups, window = cfg["heads"], cfg["dim_head"], cfg["groups"], cfg["window"]
    shift = tuple(i // 2 for i in window)
    assert cond_frames.shape[2] == tc and x.shape[2] == tp and cond_fea.shape[2] == tm + tp

    def tap(name, val):
        if taps is not None:
            taps[name] = val

    x = torch.cat([cond_frames[:, :, :tm], x], dim=2)
    pos_bias = t5_bucket_bias(sd["time_rel_pos_bias.relative_attention_bias.weight"], tm + tp)
    if v == "u22":
        # no init_noise_conv in this forward: [flow(3) | adapted cond_fea(256)] -> init_conv
        cf = unet_cond_features(sd, cfg, cond_fea, pos_bias, x.shape[-2:])
        tap("cond_up", cf)
        x = torch.cat([x, cf], dim=1)
    elif v != "base":
        x = F.conv3d(x, sd["init_noise_conv.weight"], sd["init_noise_conv.bias"], padding=(0, 3, 3))
        tap("init_noise_conv", x)
        if v == "ada":
            cf = unet_cond_features(sd, cfg, cond_fea, pos_bias, x.shape[-2:])
        else:
            cf = traj_warp(x[:, :, tc:], cond_fea, sd.sub("init_traj"), tm, tp)
            cf = _resize_frames(cf, x.shape[-2:])
        tap("cond_up", cf)
        x = torch.cat([x, cf], dim=1)
    else:
        x = torch.cat([x, cond_fea], dim=1)
    x = F.conv3d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=(0, 3, 3))
    tap("init_conv", x)
    r = x
    x = temporal_attention(x, sd.sub("init_temporal_attn"), pos_bias, heads, dh)
    tap("init_temporal_attn", x)
    t = time_embedding(time, sd, cfg["dim"])
    tap("time_emb", t)

    n_lvl = len(cfg["dim_mults"])
    skips = []
    rs = "6" if v == "u22" else "5"                 # Downsample / Upsample slot inside a stage's ModuleList

    def stage_u22(x, s, name):
        # ..._traj_ada_u22.py:1268-1279 / :1290-1301: both ResnetBlocks first, then both STW layers, the adaptor
        # (every down level, up levels > 1) and a per-level temporal attention
        x = resnet_block(x, s.sub("0"), groups, t)
        tap(name + ".0", x)
        x = resnet_block(x, s.sub("2"), groups, t)
        tap(name + ".2", x)
        x = stw_attention(x, s.sub("1"), window, shift, heads, dh)
        tap(name + ".1", x)
        x = stw_attention(x, s.sub("3"), window, (0, 0, 0), heads, dh)
        tap(name + ".3", x)
        if s.has("4.Tmodulator.weight"):
            x = motion_adaptor(x, s.sub("4"), tm, tp)
            tap(name + ".4", x)
        x = temporal_attention(x, s.sub("5"), pos_bias, heads, dh)
        tap(name + ".5", x)
        return x

    def stage(x, s, name):
        if v == "u22":
            return stage_u22(x, s, name)
        x = resnet_block(x, s.sub("0"), groups, t)
        tap(name + ".0", x)
        x = stw_attention(x, s.sub("1"), window, shift, heads, dh)
        tap(name + ".1", x)
        x = resnet_block(x, s.sub("2"), groups, t)
        tap(name + ".2", x)
        x = stw_attention(x, s.sub("3"), window, (0, 0, 0), heads, dh)
        tap(name + ".3", x)
        if s.has("4.Tmodulator.weight"):
            x = motion_adaptor(x, s.sub("4"), tm, tp)
            tap(name + ".4", x)
        return x

    for i in range(n_lvl):
        s = sd.sub(f"downs.{i}")
        x = stage(x, s, f"downs.{i}")
        skips.append(x)
        if s.has(rs + ".weight"):
            x = F.conv3d(x, s[rs + ".weight"], s[rs + ".bias"], stride=(1, 2, 2), padding=(0, 1, 1))
            tap(f"downs.{i}.{rs}", x)

    x = resnet_block(x, sd.sub("mid_block1"), groups, t)
    x = stw_attention(x, sd.sub("mid_attn1"), window, shift, heads, dh)
    if v == "u22":                                  # ..._traj_ada_u22.py:1283-1287: attn1, attn2, adaptor, block2
        x = stw_attention(x, sd.sub("mid_attn2"), window, (0, 0, 0), heads, dh)
        x = motion_adaptor(x, sd.sub("mid_adaptor"), tm, tp)
        x = resnet_block(x, sd.sub("mid_block2"), groups, t)
    else:
        x = resnet_block(x, sd.sub("mid_block2"), groups, t)
        x = stw_attention(x, sd.sub("mid_attn2"), window, (0, 0, 0), heads, dh)
        x = motion_adaptor(x, sd.sub("mid_adaptor"), tm, tp)
    tap("mid", x)

    for i in range(n_lvl):
        s = sd.sub(f"ups.{i}")
        x = torch.cat((x, skips.pop()), dim=1)
        x = stage(x, s, f"ups.{i}")
        if s.has(rs + ".weight"):
            x = F.conv_transpose3d(x, s[rs + ".weight"], s[rs + ".bias"], stride=(1, 2, 2), padding=(0, 1, 1))
            tap(f"ups.{i}.{rs}", x)

    x = torch.cat((x, r), dim=1)

    def head(s):
        y = resnet_block(x, s.sub("0"), groups, None)
        return F.conv3d(y, s["1.weight"], s["1.bias"])[:, :, tm:]

    return torch.cat((head(sd.sub("final_conv")), head(sd.sub("occlusion_map"))), dim=1)


# ----------------------------------------------------------------------------- DDIM sampler
def cosine_schedule_tables(timesteps=1000, s=0.008):
    """cosine_beta_schedule + ctor buffers, Diffusion.py:39-49,76-115 (fp64 then cast to fp32)."""
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps, dtype=torch.float64)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * torch.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.9999)
    alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
    prev = F.pad(alphas_cumprod[:-1], (1, 0), value=1.0)
    return dict(
        alphas_cumprod_prev=prev.float(),
        sqrt_recip_alphas_cumprod=torch.sqrt(1.0 / alphas_cumprod).float(),
        sqrt_recipm1_alphas_cumprod=torch.sqrt(1.0 / alphas_cumprod - 1).float(),
    )


def ddim_time_pairs(total=1000, sampling=10):
    """Diffusion.py:214-216."""
    times = torch.linspace(0.0, total, steps=sampling + 2)[:-1]
    times = list(reversed(times.int().tolist()))
    return list(zip(times[:-1], times[1:]))


def dynamic_threshold(x_start, q=0.9):
    """Diffusion.py:233-246; torch.quantile is the external ATen definition."""
    s = torch.quantile(x_start.flatten(1).abs(), q, dim=-1).clamp_(min=1.0)
    s = s.view(-1, *((1,) * (x_start.ndim - 1)))
    return x_start.clamp(-s, s) / s, s


def ddim_step(tab, img, pred_noise, time, time_next, noise, eta=1.0):
    """One iteration body of ddim_sample after the denoiser call, Diffusion.py:221-255."""
    alpha = tab["alphas_cumprod_prev"][time]
    alpha_next = tab["alphas_cumprod_prev"][time_next]
    x_start = tab["sqrt_recip_alphas_cumprod"][time] * img - tab["sqrt_recipm1_alphas_cumprod"][time] * pred_noise
    x_start, s = dynamic_threshold(x_start)
    sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
    c = ((1 - alpha_next) - sigma ** 2).sqrt()
    nz = noise if time_next > 0 else 0.0
    return x_start * alpha_next.sqrt() + c * pred_noise + sigma * nz, x_start, s


def ddim_sample(sd, cfg, x_cond, cond_fea, init_noise, step_noises, sampling=10, total=1000, eta=1.0,
                trace=None):
    """ddim_sample with *injected* noise: init_noise replaces torch.randn(shape) (Diffusion.py:218),
    step_noises[i] replaces randn_like at iteration i (Diffusion.py:251)."""
    tab = cosine_schedule_tables(total)
    img = init_noise
    unet = sd.sub("denoise_fn")
    for i, (time, time_next) in enumerate(ddim_time_pairs(total, sampling)):
        tcond = torch.full((img.shape[0],), time, dtype=torch.long)
        pred_noise = unet_forward(unet, cfg, img, tcond, x_cond, cond_fea)
        img, x_start, s = ddim_step(tab, img, pred_noise, time, time_next, step_noises[i], eta)
        if trace is not None:
            trace.append(dict(pred_noise=pred_noise, x_start=x_start, s=s.flatten(), img=img))
    return img


# ----------------------------------------------------------------------------- LFAE decode
def _bn(x, sd, name):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], training=False, eps=1e-5)


def _conv_bn_relu(x, sd, pad):
    y = F.conv2d(x, sd["conv.weight"], sd["conv.bias"], padding=pad)
    return F.relu(_bn(y, sd, "norm"))


def generator_encode(sd, src, num_down=2):
    """first + down blocks of forward_with_flow, generator.py:153-157; util.py:114-149."""
    out = _conv_bn_relu(src, sd.sub("first"), 3)
    skips = [out]
    for i in range(num_down):
        out = F.avg_pool2d(_conv_bn_relu(out, sd.sub(f"down_blocks.{i}"), 1), 2)
        skips.append(out)
    return skips


def warp(inp, flow):
    """Generator.deform_input, generator.py:63-71. flow (B,h,w,2) normalised (x,y)."""
    h, w = inp.shape[2:]
    if flow.shape[1] != h or flow.shape[2] != w:
        flow = F.interpolate(flow.permute(0, 3, 1, 2), size=(h, w), mode="bilinear").permute(0, 2, 3, 1)
    return F.grid_sample(inp, flow, align_corners=True)


def warp_index_math(flow, occ, H, W):
    """numpy fp32 restatement (no FMA contraction) of the index arithmetic under deform_input / apply_optical:
    F.interpolate(bilinear, align_corners=False) of flow (F,h,w,2) and occ (F,1,h,w) to (H,W)  [ATen
    UpSampleBilinear2d: src = max(0, scale*(dst+0.5)-0.5), h0*(w0*a + w1*b) + h1*(w0*c + w1*d)], then the
    F.grid_sample(align_corners=True) tap computation [ATen GridSampler: ix = ((x+1)/2)*(W-1), floor, weights
    nw=(x1-ix)(y1-iy) ...].  Returns (xy int32 (F,H,W,2), weights (F,H,W,4), gflow (F,H,W,3))."""
    import numpy as np
    f32 = np.float32
    fl = flow.detach().cpu().numpy().astype(f32)
    Fn, h, w = fl.shape[:3]
    oc = None if occ is None else occ.detach().cpu().numpy().astype(f32).reshape(Fn, h, w)

    def src_index(n_out, n_in):
        scale = f32(n_in) / f32(n_out)
        s = scale * (np.arange(n_out, dtype=f32) + f32(0.5)) - f32(0.5)
        s = np.maximum(s, f32(0))
        i0 = s.astype(np.int32)
        i1 = i0 + (i0 < n_in - 1)
        l1 = (s - i0.astype(f32)).astype(f32)
        return i0, i1, (f32(1) - l1).astype(f32), l1

    def resize(plane):                      # (F,h,w) -> (F,H,W)
        if h == H and w == W:
            return plane
        y0, y1, ly0, ly1 = src_index(H, h)
        x0, x1, lx0, lx1 = src_index(W, w)
        a, b = plane[:, y0][:, :, x0], plane[:, y0][:, :, x1]
        c, d = plane[:, y1][:, :, x0], plane[:, y1][:, :, x1]
        top = (lx0[None, None] * a + lx1[None, None] * b).astype(f32)
        bot = (lx0[None, None] * c + lx1[None, None] * d).astype(f32)
        return (ly0[None, :, None] * top + ly1[None, :, None] * bot).astype(f32)

    gx, gy = resize(fl[..., 0]), resize(fl[..., 1])
    go = np.ones_like(gx) if oc is None else resize(oc)
    ix = ((gx + f32(1)) * f32(0.5)) * f32(W - 1)
    iy = ((gy + f32(1)) * f32(0.5)) * f32(H - 1)
    fx, fy = np.floor(ix), np.floor(iy)
    ex, ey = (fx + f32(1)) - ix, (fy + f32(1)) - iy
    wx, wy = ix - fx, iy - fy
    wts = np.stack([ex * ey, wx * ey, ex * wy, wx * wy], -1).astype(f32)
    xy = np.stack([fx.astype(np.int32), fy.astype(np.int32)], -1)
    return torch.from_numpy(xy), torch.from_numpy(wts), torch.from_numpy(np.stack([gx, gy, go], -1).astype(f32))


def warp_from_taps(inp, xy, wts):
    """grid_sample(bilinear, zeros) evaluated from precomputed taps, ATen accumulation order nw, ne, sw, se."""
    Fn, Cc, H, W = inp.shape
    out = torch.zeros(Fn, Cc, H, W)
    x0, y0 = xy[..., 0].long(), xy[..., 1].long()
    fi = torch.arange(Fn)[:, None, None].expand(Fn, H, W)
    for k, (dx, dy) in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
        xs, ys = x0 + dx, y0 + dy
        ok = (xs >= 0) & (xs < W) & (ys >= 0) & (ys < H)
        v = inp[fi, :, ys.clamp(0, H - 1), xs.clamp(0, W - 1)]           # (F,H,W,C)
        out = out + (v * (wts[..., k] * ok)[..., None]).permute(0, 3, 1, 2)
    return out


def _blend(prev, skip, flow, occ):
    """apply_optical, generator.py:74-93."""
    skip = warp(skip, flow)
    if occ is not None:
        if occ.shape[2:] != skip.shape[2:]:
            occ = F.interpolate(occ, size=skip.shape[2:], mode="bilinear")
        skip = skip * occ + prev * (1 - occ) if prev is not None else skip * occ
    return skip


def generator_forward_with_flow(sd, src, flow, occ, num_down=2, num_bottleneck=6):
    """Generator.forward_with_flow (skips=True), generator.py:152-206."""
    skips = generator_encode(sd, src, num_down)
    deformed = warp(src, flow)
    out = _blend(None, skips[-1], flow, occ)
    for i in range(num_bottleneck):
        r = sd.sub(f"bottleneck.r{i}")
        y = F.conv2d(F.relu(_bn(out, r, "norm1")), r["conv1.weight"], r["conv1.bias"], padding=1)
        y = F.conv2d(F.relu(_bn(y, r, "norm2")), r["conv2.weight"], r["conv2.bias"], padding=1)
        out = out + y
    for i in range(num_down):
        out = _blend(out, skips[-(i + 1)], flow, occ)
        u = sd.sub(f"up_blocks.{i}")
        out = F.interpolate(out, scale_factor=2)
        out = F.relu(_bn(F.conv2d(out, u["conv.weight"], u["conv.bias"], padding=1), u, "norm"))
    out = _blend(out, skips[0], flow, occ)
    out = torch.sigmoid(F.conv2d(out, sd["final.weight"], sd["final.bias"], padding=3))
    out = _blend(out, src, flow, occ)
    return dict(prediction=out, deformed=deformed)


def decode_video(gen_sd, ref_img, grid, conf):
    """Decode loop of sample_one_video, VideoFlowDiffusion_multi_w_ref.py:292-308.
    grid (B,2,T,h,w), conf (B,1,T,h,w) or None -> (out (B,3,T,H,W), warped (B,3,T,H,W))."""
    outs, warps = [], []
    for i in range(grid.shape[2]):
        g = generator_forward_with_flow(gen_sd, ref_img, grid[:, :, i].permute(0, 2, 3, 1),
                                        None if conf is None else conf[:, :, i])
        outs.append(g["prediction"])
        warps.append(g["deformed"])
    return torch.stack(outs, dim=2), torch.stack(warps, dim=2)


# ----------------------------------------------------------------------------- fixtures helpers
def perturb_state_dict(sd, seed):
    """Deterministic 'perturbed random init' (SURVEY fact 8): every float tensor gets 0.02*randn added
    (all-zero tensors become 0.02*randn), BN running stats become non-trivial.  Integer buffers and
    rotary freqs / schedule tables are left alone.  Iterates keys in sorted order with one generator so
    the result depends only on (key set, shapes, seed)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    skip = ("rotary_emb.freqs", "relative_position_index", "num_batches_tracked")
    sched = ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
             "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
             "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
             "posterior_mean_coef1", "posterior_mean_coef2")
    for k in sorted(sd.keys()):
        v = sd[k]
        if (not v.is_floating_point()) or k in sched or any(k.endswith(s) for s in skip):
            out[k] = v.clone()
            continue
        r = torch.randn(v.shape, generator=g)
        if k.endswith("running_var"):
            out[k] = 1.0 + 0.1 * torch.rand(v.shape, generator=g)
        elif k.endswith("running_mean"):
            out[k] = 0.1 * r
        else:
            out[k] = v + 0.02 * r
    return out
