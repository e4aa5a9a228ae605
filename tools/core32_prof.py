"""Per-role cycle counters of the tcgen05 window-attention core (csrc/attn_core32.cu) at the BAIR level-1/2/3 shapes.
Needs the profiling build:  EXTDM_BUILD_TAG=prof EXTDM_NVCC_DEFS=-DEXTDM_CORE32_PROF python <pkg>/build.py ;
EXTDM_LIB=<pkg>/libextdm_b200_prof.so python tools/core32_prof.py.  Also times the product / legacy kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import extdm_b200  # noqa: E402,F401
from extdm_b200 import ops  # noqa: E402
from extdm_b200.unet import _rope_tables  # noqa: E402

BF, dev = torch.bfloat16, "cuda"
R = ops.IMMEDIATE
heads, dh = 8, 32
hid = heads * dh
for B, T, H in [(32, 12, 16), (32, 12, 8), (32, 12, 4), (32, 14, 16)]:
    qkv = torch.randn(B, T, H, H, 3 * hid, device=dev).to(BF)
    out = torch.zeros(B, T, H, H, hid, device=dev, dtype=BF)
    tbl = torch.randn(147, heads, device=dev) * 0.5
    rc, rs = _rope_tables(32, dh, dev)
    shift = (1, 2 if H > 4 else 0, 2 if H > 4 else 0)
    fn = lambda: ops.window_attention(R, qkv, out, tbl, rc, rs, heads, dh, (2, 4, 4), shift)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn()
    b.record()
    torch.cuda.synchronize()
    print(f"window core B={B} T={T} H={H}: {a.elapsed_time(b) / 10 * 1e3:7.1f} us", flush=True)
