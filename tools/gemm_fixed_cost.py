"""Fixed cost of one conv_gemm launch inside a CUDA graph: a chain of N dependent tiny GEMMs (one 128-row tile, K = 64 ..
2304), microseconds per graph node.  EXTDM_GEMM_DBG isolates the parts (4 = no TMA loads, 2 = no MMA issue, 8 = no
tcgen05.ld, 1 = no stores).  GPU box only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import extdm_b200  # noqa: E402,F401
from extdm_b200 import ops  # noqa: E402

dev, BF = "cuda", torch.bfloat16
N = 400


def chain(rows, cin, cout, k, label):
    H = 4
    B = max(1, rows // (H * H * 8))
    x = torch.randn(B, 8, H, H, cin, device=dev).to(BF)
    ys = [torch.zeros(B, 8, H, H, cout, device=dev, dtype=BF) for _ in range(2)]
    w = (torch.randn(cout, k * k * cin, device=dev) * 0.05).to(BF)
    bias = torch.zeros(cout, device=dev)
    rec = ops.Recorder(record=True)
    for i in range(N):
        src = x if (i == 0 or cin != cout) else ys[(i + 1) % 2]
        ops.conv_cl(rec, src, w, cout, k, ys[i % 2], bias=bias)
    rec.run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        rec.run()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    print(f"{label:40s} rows={B * 8 * H * H:6d} cin={cin:4d} cout={cout:4d} k={k}: {a.elapsed_time(b) / 5 / N * 1e3:7.2f} us per node",
          flush=True)


tag = "dbg=" + os.environ.get("EXTDM_GEMM_DBG", "0")
chain(128, 64, 64, 1, tag + " 1 tile 1x1")
chain(128, 256, 256, 1, tag + " 1 tile 1x1 K=256")
chain(6144, 256, 256, 1, tag + " level-3 1x1")
chain(6144, 256, 256, 3, tag + " level-3 3x3")
chain(6144, 256, 768, 1, tag + " level-3 qkv")
