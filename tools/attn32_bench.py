"""Time the dim_head-32 attention layers at the BAIR / SMMNIST level-0 shapes (CUDA events, 20 launches each).
   python tools/attn32_bench.py            # tcgen05 kernel (attn_ws32.cu); EXTDM_ATTN32_PROF=1 prints per-role cycles
   EXTDM_ATTN32_LEGACY=1 python tools/attn32_bench.py   # the mma.sync kernels"""
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
heads, dh, C = 8, 32, 64
hid = heads * dh


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for B, T, H in [(32, 12, 32), (32, 14, 32), (32, 12, 16), (1, 14, 32)]:
    x = torch.randn(B, T, H, H, C, device=dev).to(BF)
    y = torch.zeros_like(x)
    gamma = torch.ones(C, device=dev)
    wqkv = (torch.randn(3 * hid, C, device=dev) * C ** -0.5).to(BF)
    wproj = (torch.randn(C, hid, device=dev) * hid ** -0.5).to(BF)
    pb = torch.zeros(C, device=dev)
    tbl = torch.randn(147, heads, device=dev) * 0.5
    rc, rs = _rope_tables(32, dh, dev)
    for shift in [(0, 0, 0), (1, 2, 2)]:
        us = timeit(lambda: ops.stw_fused(R, x, y, gamma, wqkv, wproj, pb, tbl, rc, rs, heads, dh, (2, 4, 4), shift))
        print(f"stw (2,4,4) dh32 B={B} T={T} H={H} shift={shift}: {us:8.1f} us   {B*T*H*H*4*C/us/1e3:7.0f} GB/s algorithmic")
    lnw, lnb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    rel = torch.randn(heads, 2 * T - 1, device=dev) * 0.5
    us = timeit(lambda: ops.temporal_fused(R, x, y, gamma, lnw, lnb, wqkv, wproj, rel, rc, rs, heads, dh))
    print(f"temporal dh32 B={B} T={T} H={H}: {us:8.1f} us")
