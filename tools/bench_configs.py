"""Predicted frames/s of every dataset configuration (SURVEY.md section 8d, configs 1-5 + the ada_u22 pairing) on one
GPU: on-device autoregressive rollout, batch 32 videos, synthetic clips, random-init weights.  These are the
parity-test configurations, measured for reference; bench.py's line stays on KTH.  GPU box only.

    python tools/bench_configs.py [name ...]      -> markdown table on stdout
"""
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import extdm_b200  # noqa: E402,F401
from extdm_b200 import configs  # noqa: E402

names = sys.argv[1:] or ["kth", "bair", "smmnist", "ucf", "cityscapes", "cityscapes64", "cityscapes_u22"]
B = int(os.environ.get("EXTDM_BENCH_BATCH", "32"))
print("| config | UNet | tc -> total | rounds | batch | ms / rollout | predicted frames/s |")
print("|---|---|---|---|---|---|---|")
for name in names:
    model, cfg = configs.build_model(name, device="cuda")
    tc = model.cond_frame_num
    total = cfg["dataset_params"]["valid_params"]["pred_frames"]
    hw = cfg["dataset_params"]["frame_shape"]
    clip = torch.rand(B, 1, tc, hw, hw, device="cuda").expand(B, 3, tc, hw, hw).contiguous()
    t_warm, n_warm = time.perf_counter(), 0              # until the SM clock has ramped after the idle model build
    while n_warm < 2 or time.perf_counter() - t_warm < 1.0:
        configs.rollout(model, clip, total)
        torch.cuda.synchronize()
        n_warm += 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n):
        out = configs.rollout(model, clip, total)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    assert out.shape[2] == total and torch.isfinite(out).all()
    rounds = math.ceil(total / model.pred_frame_num)
    print(f"| {name} {hw}x{hw} | {model.unet.cfg.variant} | {tc} -> {total} | {rounds} | {B} | {ms:.1f} | "
          f"{B * total / ms * 1e3:.0f} |", flush=True)
    del model
    torch.cuda.empty_cache()
