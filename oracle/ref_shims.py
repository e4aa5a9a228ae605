"""Import shims that let the UNMODIFIED reference import and run (on CPU, or on a GPU as it is).

Used by tests/golden/make_golden.py (fixture generation in the build container, from /root/reference) and by
bench.py's reference legs (`--impl reference`, `cpu_baseline`, `gpu_eager_reference`), which run the copy staged under
oracle/_ref/ by oracle/make_ref.py -- /root/reference does not exist on the GPU box.  Test infrastructure: nothing in
the product package, the gpu tests or smoke() imports this.

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

import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# the build container reads the reference where it lies; elsewhere the staged copy (oracle/make_ref.py)
REF_ROOT = os.environ.get("EXTDM_REFERENCE") or \
    ("/root/reference" if os.path.isdir("/root/reference/model") else os.path.join(_HERE, "_ref"))


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


_REAL_CUDA = (nn.Module.cuda, torch.Tensor.cuda)


def cuda_identity(on):
    """The reference hard-codes `.cuda()` at VideoFlowDiffusion_multi_w_ref.py:49,58 (and the other wrappers): patch it
    to the identity to run the unmodified code on the host cores; restore it to run the same code on a GPU."""
    if on:
        nn.Module.cuda = lambda self, *a, **k: self
        torch.Tensor.cuda = lambda self, *a, **k: self
    else:
        nn.Module.cuda, torch.Tensor.cuda = _REAL_CUDA


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "model", "BaseDM_adaptor"))


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
        cuda_identity(True)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    install._done = True
