"""Replay selected launches of one KTH DDIM step / decode for an `ncu --set full --profile-from-start off` capture.
   python tools/ncu_target.py [--batch 32] key [key ...]
A key is a substring of the launch label printed by tools/kernel_table.py, e.g. "gemm rows=983040 n=64 k=576" or
"groupnorm_stats"; the first matching launch of (prologue, step, decode) is replayed once between
cudaProfilerStart/Stop after a full un-profiled warm-up round.  GPU box only."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import extdm_b200  # noqa: E402
from extdm_b200 import configs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--dataset", default="kth")
ap.add_argument("keys", nargs="+")
args = ap.parse_args()
B = args.batch
model, cfg = configs.build_model(args.dataset, device="cuda")
tc, tp = model.cond_frame_num, model.pred_frame_num
clip = torch.rand(B, 3, tc, 64, 64, device="cuda")
model.sample_one_video(1.0, clip)
torch.cuda.synchronize()
runner = model.unet.runner(B, 32, 32, 16)
dec = model.generator.decoder(B, tc + tp, 64, 64, 32, 32, True)
launches = []
for label, rec in (("prologue", runner.prologue), ("step", runner.step), ("decode", dec.rec)):
    for (fn, a, name), meta in zip(rec.steps, rec.meta):
        if name == "extdm_conv_gemm":
            key = f"{label} gemm rows={meta['rows']} n={meta['n']} k={meta['k']} taps={meta['taps']}"
        else:
            key = f"{label} {name.replace('extdm_', '')}" + (" " + meta["tag"] if "tag" in meta else "")
        launches.append((key, fn, a))
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for want in args.keys:
    hit = [l for l in launches if want in l[0]]
    if not hit:
        print("no launch matches", want)
        continue
    key, fn, a = hit[0]
    fn(*a, stream)                       # warm (un-profiled)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    rc = fn(*a, stream)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled:", key, "rc", rc)
