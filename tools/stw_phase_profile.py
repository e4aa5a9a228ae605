"""Per-phase cycle counts of the tcgen05 window-attention kernel (csrc/stw_tc.cu) at the KTH level-0 shape:
   EXTDM_STW_PROF=1 python tools/stw_phase_profile.py
The launcher prints, per launch, the average cycles CTA 0 spends per window pair in each phase of the software
pipeline.  GPU box only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("EXTDM_STW_PROF", "1")
import torch  # noqa: E402
import extdm_b200  # noqa: E402,F401
from extdm_b200 import ops  # noqa: E402
from extdm_b200.unet import _rope_tables  # noqa: E402

torch.manual_seed(0)
B, T, H, W, C = 32, 30, 32, 32, 64
dev = "cuda"
x = torch.randn(B, T, H, W, C, device=dev).bfloat16()
y = torch.empty_like(x)
g = torch.ones(C, device=dev)
wqkv = (torch.randn(384, C, device=dev) * 0.1).bfloat16()
wp = (torch.randn(C, 128, device=dev) * 0.1).bfloat16()
pb = torch.zeros(C, device=dev)
tbl = torch.randn(343, 8, device=dev) * 0.1
rc, rs = _rope_tables(64, 16, dev)
for shift in ((0, 0, 0), (2, 2, 2)):
    for _ in range(2):
        ops.stw_fused(ops.IMMEDIATE, x, y, g, wqkv, wp, pb, tbl, rc, rs, 8, 16, (4, 4, 4), shift)
torch.cuda.synchronize()
