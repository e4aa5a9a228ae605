"""Per-shape kernel time table of one round (CUDA events per launch, eager replay).  GPU box only.
   python tools/kernel_table.py [batch] [config name, default kth]  -> markdown on stdout"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import extdm_b200  # noqa: E402
from extdm_b200 import configs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
NAME = sys.argv[2] if len(sys.argv) > 2 else "kth"
model, cfg = configs.build_model(NAME, device="cuda")
tc, tp = model.cond_frame_num, model.pred_frame_num
HW = cfg["dataset_params"]["frame_shape"]
clip = torch.rand(B, 3, tc, HW, HW, device="cuda")
for _ in range(2):
    model.sample_one_video(1.0, clip)
torch.cuda.synchronize()
runner = next(iter(model.unet._runners.values()))
dec = next(r for k, r in model.generator._runners.items() if k[0] != "enc")
steps = cfg["diffusion_params"]["model_params"]["sampling_timesteps"]
rows = {}
cond = [("cond", r_, 1) for cr in model._cond_runners.values() for r_ in cr.recorders()]
cond += [("cond", r_.rec, 1) for k_, r_ in model.generator._runners.items() if k_[0] == "enc"]
for label, rec, mult in [("prologue", runner.prologue, 1), ("step", runner.step, steps), ("decode", dec.rec, 1)] + cond:
    rec.run()
    torch.cuda.synchronize()
    best = None
    for rep in range(3):
        cur = rec.run_timed()
        best = cur if best is None else [(n, m, min(t, t2)) for (n, m, t), (_, _, t2) in zip(best, cur)]
    for name, meta, ms in best:
        if name == "extdm_conv_gemm":
            key = (label, f"gemm{'[tf32]' if meta.get('tf32') else ''} rows={meta['rows']} n={meta['n']} k={meta['k']} "
                          f"taps={meta['taps']}")
        else:
            key = (label, name.replace("extdm_", "") + (" " + meta["tag"] if "tag" in meta else ""))
        r = rows.setdefault(key, dict(ms=0.0, n=0, flops=0.0, bytes=0.0))
        r["ms"] += ms * mult
        r["n"] += mult
        r["flops"] += meta.get("flops", 0.0) * mult
        r["bytes"] += meta.get("bytes", 0.0) * mult
total = sum(r["ms"] for r in rows.values())
print(f"| phase | kernel / shape | launches/round | ms/round | share | TFLOP/s | GB/s (alg.) |")
print("|---|---|---|---|---|---|---|")
for (label, key), r in sorted(rows.items(), key=lambda kv: -kv[1]["ms"]):
    tf = r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["flops"] else 0
    gb = r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["bytes"] else 0
    print(f"| {label} | {key} | {r['n']} | {r['ms']:.3f} | {100 * r['ms'] / total:.1f}% | {tf:.0f} | {gb:.0f} |")
print(f"\ntotal kernel time per round: {total:.1f} ms (batch {B})")
