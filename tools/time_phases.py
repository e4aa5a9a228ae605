"""Wall/device time of the three phases of one sample_one_video round (KTH, batch 32) + top torch ops of the
conditioning stage.  GPU box only."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import extdm_b200  # noqa: E402
from extdm_b200 import configs  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "kth"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
model, cfg = configs.build_model(name, device="cuda")
tc, tp = model.cond_frame_num, model.pred_frame_num
hw = cfg["dataset_params"]["frame_shape"]
clip = torch.rand(B, 3, tc, hw, hw, device="cuda")
t_warm = time.perf_counter()
while time.perf_counter() - t_warm < 1.5:            # until the SM clock has ramped after the idle model build
    model.sample_one_video(1.0, clip)
    torch.cuda.synchronize()


def timed(fn, n=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return r, e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n


(ret, x_cond, fea, ref), dev_ms, wall_ms = timed(lambda: model.condition(clip))
print(f"condition   device {dev_ms:8.2f} ms  wall {wall_ms:8.2f} ms")
pred, dev_ms, wall_ms = timed(lambda: model.diffusion.sample(x_cond, cond_fea=fea, batch_size=1, cond_scale=1.0))
print(f"ddim sample device {dev_ms:8.2f} ms  wall {wall_ms:8.2f} ms")
grid = torch.cat([ret["real_vid_grid"][:, :, :tc], pred[:, :2]], dim=2)
conf = torch.cat([ret["real_vid_conf"][:, :, :tc], (pred[:, 2:3] + 1) * 0.5], dim=2)
_, dev_ms, wall_ms = timed(lambda: model.generator.decode_video(ref, grid, conf))
print(f"decode      device {dev_ms:8.2f} ms  wall {wall_ms:8.2f} ms")
_, dev_ms, wall_ms = timed(lambda: model.sample_one_video(1.0, clip))
print(f"full round  device {dev_ms:8.2f} ms  wall {wall_ms:8.2f} ms")

from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    model.condition(clip)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
