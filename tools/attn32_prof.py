"""Per-role cycle counts of attn_ws32 (EXTDM_ATTN32_PROF=1 is set here) at the BAIR level-0 shape: one launch each of the
window and the temporal layer; the launcher prints the counters of CTA 0 to stderr."""
import os
import sys

os.environ["EXTDM_ATTN32_PROF"] = "1"
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
B, T, H = 32, 12, 32
x = torch.randn(B, T, H, H, C, device=dev).to(BF)
y = torch.zeros_like(x)
gamma = torch.ones(C, device=dev)
wqkv = (torch.randn(3 * hid, C, device=dev) * C ** -0.5).to(BF)
wproj = (torch.randn(C, hid, device=dev) * hid ** -0.5).to(BF)
pb = torch.zeros(C, device=dev)
tbl = torch.randn(147, heads, device=dev) * 0.5
rc, rs = _rope_tables(32, dh, dev)
lnw, lnb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
rel = torch.randn(heads, 2 * T - 1, device=dev) * 0.5
for _ in range(2):
    ops.stw_fused(R, x, y, gamma, wqkv, wproj, pb, tbl, rc, rs, heads, dh, (2, 4, 4), (1, 2, 2))
    ops.temporal_fused(R, x, y, gamma, lnw, lnb, wqkv, wproj, rel, rc, rs, heads, dh)
torch.cuda.synchronize()
