"""Synthetic, fully deterministic weights for a given (key -> shape) manifest.

Used by bench.py (random-init weights of the named architecture; there are no checkpoints offline) and
by the parity tests, which feed the *same* tensors to the reference (fixture generation), the oracle and
the CUDA path.  Unlike the reference's default init, nothing is left at zero (the
zero-initialised adaptor convs / bg-predictor fc would otherwise hide whole kernels behind a multiply by 0).
"""
import math
import torch

_SCHEDULE_KEYS = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
    "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
    "posterior_mean_coef1", "posterior_mean_coef2",
)
_DERIVED_SUFFIXES = ("rotary_emb.freqs", "relative_position_index", "num_batches_tracked")


def is_derived_key(key):
    """Buffers whose values are constants of the architecture (never synthesised or perturbed)."""
    return key in _SCHEDULE_KEYS or key.endswith(_DERIVED_SUFFIXES) or key == "down.weight" \
        or key.endswith(".down.weight")


def synth_state_dict(shapes, seed, base=None):
    """shapes: {key: shape}.  Keys for derived buffers are copied from `base` (a state_dict holding the
    constants) when given, otherwise skipped.  Iteration is in sorted key order from one CPU generator."""
    g = torch.Generator().manual_seed(int(seed))
    out = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        if is_derived_key(k):
            if base is not None and k in base:
                out[k] = base[k].clone()
            continue
        r = torch.randn(shp, generator=g)
        leaf = k.rsplit(".", 1)[-1]
        if leaf == "running_var":
            v = 1.0 + 0.2 * torch.rand(shp, generator=g)
        elif leaf == "running_mean":
            v = 0.1 * r
        elif leaf == "gamma" or (leaf == "weight" and len(shp) == 1):
            v = 1.0 + 0.1 * r                      # norm scales
        elif leaf == "bias":
            v = 0.05 * r
        elif leaf == "relative_position_bias_table" or "relative_attention_bias" in k:
            v = 0.3 * r
        elif leaf == "weight" and len(shp) >= 2:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            if "ups." in k and k.endswith(".5.weight"):  # ConvTranspose3d: (in, out, 1, 4, 4)
                fan_in = shp[0] * 4
            v = r / math.sqrt(max(fan_in, 1))
        else:
            v = 0.1 * r
        out[k] = v.contiguous()
    return out
