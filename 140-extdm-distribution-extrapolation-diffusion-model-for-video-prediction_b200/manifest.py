"""State-dict manifests (key -> shape) of the reference module trees for the hot path, derived from the
architecture hyper-parameters only.  They let the product build parameter containers whose
`state_dict()` / `load_state_dict()` key sets are identical to the reference's (SURVEY.md section 8b:
`model.diffusion.load_state_dict(ckpt['diffusion'])`, scripts/DM/valid.py:111-112) without copying
its module definitions.  tests/test_host_cpu.py::test_unet_manifest_matches_reference checks key sets and shapes of
every variant (ada, u12, base, ada_u22; mini and shipped tc / tp) against manifests captured from the reference
itself (tests/golden/unet_*.pt); test_wrapper_state_dicts_match_reference does the same for the wrappers' parts.
"""
import math

UNET_ARCHITECTURES = {
    "DenoiseNet_STWAtt_w_w_ref_adaptor_cross_multi_traj_ada": "ada",
    "DenoiseNet_STWAtt_w_w_ref_adaptor_cross_multi_traj_u12": "u12",
    "DenoiseNet_STWAtt_w_w_ref_adaptor_cross_multi_traj_u22": "u12",   # byte-identical file in the reference
    "DenoiseNet_STWAtt_w_wo_ref_adaptor_cross_multi": "base",
    "DenoiseNet_STWAtt_w_w_ref_adaptor_cross_multi_traj_ada_u22": "u22",
}

SCHEDULE_KEYS = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
    "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
    "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2",
)


def adaptor_layers(tm, tp):
    """compute_layer (..._traj_ada.py:644-649, l=None): (#extrapolation rounds, #extrapolated frames)."""
    L = max(1, int(math.ceil(math.log2((tp + 1) / tm))))
    return L, (2 ** L - 1) * tm


class UnetConfig:
    """Hyper-parameters of one Unet3D variant."""

    def __init__(self, variant, tc, tp, dim=64, dim_mults=(1, 2, 4, 4), channels=None, heads=8, groups=8):
        if variant not in ("ada", "u12", "base", "u22"):
            raise ValueError(f"unknown Unet3D variant {variant!r}")
        self.variant, self.tc, self.tp = variant, int(tc), int(tp)
        self.tm = self.tc - 1 if variant == "base" else self.tc
        self.T = self.tm + self.tp
        self.dim, self.dim_mults = int(dim), tuple(dim_mults)
        self.heads, self.groups = heads, groups
        # ..._traj_ada.py:872-877 (4,4,4)/16; ..._traj_ada_u22.py:1016-1021 (4,4,4)/32; the others (2,4,4)/32
        self.window = (4, 4, 4) if variant in ("ada", "u22") else (2, 4, 4)
        self.dim_head = 16 if variant == "ada" else 32
        self.hidden = self.heads * self.dim_head
        self.shift = tuple(w // 2 for w in self.window)
        # base and ada_u22 feed the 3-channel flow volume to init_conv directly (no init_noise_conv in their forward)
        self.channels = channels if channels is not None else (3 + 256 if variant in ("base", "u22") else 512)
        # ada_u22: a MotionAdaptor and a temporal attention at every level, 3x3x3 extrapolators, blocks re-ordered
        self.per_level_temporal = variant == "u22"
        self.extrap_kt = 3 if variant == "u22" else 1
        self.resample_slot = 6 if variant == "u22" else 5   # index of Downsample / Upsample in a stage's ModuleList
        self.levels = [self.dim * m for m in self.dim_mults]
        self.L, self.n_extra = adaptor_layers(self.tm, self.tp)

    def key(self):
        return (self.variant, self.tc, self.tp, self.dim, self.dim_mults, self.channels)


def unet_manifest(cfg):
    m = {}
    d, hid, heads, dh = cfg.dim, cfg.hidden, cfg.heads, cfg.dim_head
    wd, wh, ww = cfg.window
    n_tok = wd * wh * ww
    n_tbl = (2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1)
    rot = min(32, dh) // 2

    def temporal(p, C):
        m[f"{p}.fn.fn.fn.norm.weight"] = (C,)
        m[f"{p}.fn.fn.fn.norm.bias"] = (C,)
        m[f"{p}.fn.fn.fn.attn.rotary_emb.freqs"] = (rot,)
        m[f"{p}.fn.fn.fn.attn.to_qkv.weight"] = (3 * hid, C)
        m[f"{p}.fn.fn.fn.attn.to_out.weight"] = (C, hid)
        m[f"{p}.fn.norm.gamma"] = (1, C, 1, 1, 1)

    def stw(p, C):
        m[f"{p}.fn.fn.attn.relative_position_bias_table"] = (n_tbl, heads)
        m[f"{p}.fn.fn.attn.relative_position_index"] = (n_tok, n_tok)
        m[f"{p}.fn.fn.attn.rotary_emb.freqs"] = (rot,)
        m[f"{p}.fn.fn.attn.qkv.weight"] = (3 * hid, C)
        m[f"{p}.fn.fn.attn.proj.weight"] = (C, hid)
        m[f"{p}.fn.fn.attn.proj.bias"] = (C,)
        m[f"{p}.fn.norm.gamma"] = (1, C, 1, 1, 1)

    def res(p, cin, cout, time=True):
        if time:
            m[f"{p}.mlp.1.weight"] = (2 * cout, 4 * d)
            m[f"{p}.mlp.1.bias"] = (2 * cout,)
        for blk, ci in (("block1", cin), ("block2", cout)):
            m[f"{p}.{blk}.proj.weight"] = (cout, ci, 1, 3, 3)
            m[f"{p}.{blk}.proj.bias"] = (cout,)
            m[f"{p}.{blk}.norm.weight"] = (cout,)
            m[f"{p}.{blk}.norm.bias"] = (cout,)
        if cin != cout:
            m[f"{p}.res_conv.weight"] = (cout, cin, 1, 1, 1)
            m[f"{p}.res_conv.bias"] = (cout,)

    def adaptor(p, C):
        m[f"{p}.adaptors.predictor.fn.fn.weight"] = (C, C, 1, 1, 1)
        m[f"{p}.adaptors.predictor.fn.fn.bias"] = (C,)
        m[f"{p}.adaptors.predictor.fn.norm.gamma"] = (1, C, 1, 1, 1)
        for i in range(cfg.L):
            m[f"{p}.adaptors.extrapolators.{i}.fn.weight"] = (C, C, cfg.extrap_kt, 3, 3)
        m[f"{p}.Tmodulator.weight"] = (C * cfg.tp, C * cfg.n_extra, 1, 1)
        m[f"{p}.Tmodulator.bias"] = (C * cfg.tp,)
        m[f"{p}.fuser.fn.weight"] = (C, 2 * C, 1, 1, 1)
        m[f"{p}.fuser.fn.bias"] = (C,)
        m[f"{p}.fuser.norm.gamma"] = (1, 2 * C, 1, 1, 1)

    m["time_rel_pos_bias.relative_attention_bias.weight"] = (32, heads)
    m["init_conv.weight"] = (d, cfg.channels, 1, 7, 7)
    m["init_conv.bias"] = (d,)
    temporal("init_temporal_attn", d)
    if cfg.variant != "base":                             # constructed (and checkpointed) but unused by ada_u22
        m["init_noise_conv.weight"] = (256, 3, 1, 7, 7)
        m["init_noise_conv.bias"] = (256,)
    if cfg.variant in ("ada", "u22"):
        temporal("cond_temporal_attn", 256)
        adaptor("cond_adaptor", 256)
    if cfg.variant == "u22":
        # used by forward(path=1) only; the sampler calls path=0 (..._traj_ada_u22.py:1048,1123-1124,1181-1210)
        m["rel_pos_bias_thw.relative_attention_bias.weight"] = (32, heads)
        m["alpha"] = (heads,)
        m["beta"] = (heads,)
    if cfg.variant == "u12":
        adaptor("init_adaptor", 256)                      # constructed but unused by the reference forward
        for n in ("q", "k", "v", "o"):
            m[f"init_traj.cross_att.linear_{n}.weight"] = (256, 256)
            m[f"init_traj.cross_att.linear_{n}.bias"] = (256,)
        m["init_traj.fuser.weight"] = (256, 512, 1, 1, 1)
        m["init_traj.fuser.bias"] = (256,)
    m["time_mlp.1.weight"] = (4 * d, d)
    m["time_mlp.1.bias"] = (4 * d,)
    m["time_mlp.3.weight"] = (4 * d, 4 * d)
    m["time_mlp.3.bias"] = (4 * d,)

    dims = [d] + cfg.levels
    in_out = list(zip(dims[:-1], dims[1:]))
    nres = len(in_out)
    u22 = cfg.variant == "u22"
    rs = cfg.resample_slot
    for i, (ci, co) in enumerate(in_out):
        p = f"downs.{i}"
        res(f"{p}.0", ci, co)
        stw(f"{p}.1", co)
        res(f"{p}.2", co, co)
        stw(f"{p}.3", co)
        if i > 1 or u22:
            adaptor(f"{p}.4", co)
        if u22:
            temporal(f"{p}.5", co)
        if i < nres - 1:
            m[f"{p}.{rs}.weight"] = (co, co, 1, 4, 4)
            m[f"{p}.{rs}.bias"] = (co,)
    mid = dims[-1]
    res("mid_block1", mid, mid)
    stw("mid_attn1", mid)
    res("mid_block2", mid, mid)
    stw("mid_attn2", mid)
    adaptor("mid_adaptor", mid)
    for i, (ci, co) in enumerate(reversed(in_out)):
        p = f"ups.{i}"
        res(f"{p}.0", co * 2, ci)
        stw(f"{p}.1", ci)
        res(f"{p}.2", ci, ci)
        stw(f"{p}.3", ci)
        if i > 1:
            adaptor(f"{p}.4", ci)
        if u22:
            temporal(f"{p}.5", ci)
        if i < nres - 1:
            m[f"{p}.{rs}.weight"] = (ci, ci, 1, 4, 4)
            m[f"{p}.{rs}.bias"] = (ci,)
    for head, oc in (("final_conv", 2), ("occlusion_map", 1)):
        res(f"{head}.0", 2 * d, d, time=False)
        m[f"{head}.1.weight"] = (oc, d, 1, 1, 1)
        m[f"{head}.1.bias"] = (oc,)
    return m


INT_SUFFIXES = ("relative_position_index", "num_batches_tracked")


def is_int_key(k):
    return k.endswith(INT_SUFFIXES)
