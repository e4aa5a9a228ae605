"""Import shims that let the UNMODIFIED reference (/root/reference) import and run on CPU.

Only used by tests/golden/make_golden.py (fixture generation, run in the build container where
/root/reference exists).  Nothing in the product, the gpu tests, smoke() or bench.py imports this.

The reference depends on packages absent from this image:
  einops_exts.rearrange_many, timm.models.layers.{DropPath,trunc_normal_}, xformers.ops,
  matplotlib.pyplot, skimage.draw.disk, rotary_embedding_torch.RotaryEmbedding.
All but the last are import-time only.  RotaryEmbedding is real arithmetic: it is restated here from the
published behaviour of rotary-embedding-torch==0.8.3 (freqs_for='lang', theta=10000, interleaved pairs,
seq_dim=-2).  The reference has no test that pins it => "parity unpinned" for that one dependency.
"""
import sys
import types
import math
import torch
from torch import nn

REF_ROOT = "/root/reference"


class _RotaryEmbedding(nn.Module):
    def __init__(self, dim, theta=10000):
        super().__init__()
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        self.freqs = nn.Parameter(freqs, requires_grad=False)

    def rotate_queries_or_keys(self, t, seq_dim=-2):
        n = t.shape[seq_dim]
        pos = torch.arange(n, device=t.device, dtype=self.freqs.dtype)
        ang = pos[:, None] * self.freqs[None, :]            # (n, dim/2)
        ang = ang.repeat_interleave(2, dim=-1)              # (n, dim) pairs share an angle
        rot_dim = ang.shape[-1]
        t_rot, t_pass = t[..., :rot_dim], t[..., rot_dim:]
        x = t_rot.reshape(*t_rot.shape[:-1], rot_dim // 2, 2)
        x1, x2 = x.unbind(-1)
        half = torch.stack((-x2, x1), dim=-1).reshape(t_rot.shape)
        out = t_rot * ang.cos() + half * ang.sin()
        return torch.cat((out, t_pass), dim=-1)


def install():
    """Register shim modules and CPU patches; idempotent."""
    if getattr(install, "_done", False):
        return
    sys.dont_write_bytecode = True
    from einops import rearrange

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mod("einops_exts", rearrange_many=lambda ts, pattern, **kw: tuple(rearrange(t, pattern, **kw) for t in ts))
    mod("rotary_embedding_torch", RotaryEmbedding=_RotaryEmbedding)

    class DropPath(nn.Identity):
        def __init__(self, *a, **k):
            super().__init__()

    def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b)

    timm = mod("timm")
    timm.models = mod("timm.models")
    timm.models.layers = mod("timm.models.layers", DropPath=DropPath, trunc_normal_=trunc_normal_)
    xf = mod("xformers")
    xf.ops = mod("xformers.ops")
    mpl = mod("matplotlib")
    mpl.pyplot = mod("matplotlib.pyplot", get_cmap=lambda *a, **k: None)
    sk = mod("skimage")
    sk.draw = mod("skimage.draw", disk=lambda *a, **k: None)

    if not torch.cuda.is_available():
        nn.Module.cuda = lambda self, *a, **k: self
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    install._done = True
