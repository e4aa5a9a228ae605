#!/usr/bin/env python
"""bench.py -- predicted frames/s of the ExtDM sampling hot path (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W]           # this repo's CUDA path
  python bench.py --impl reference [...]                        # the reference algorithm on the host CPU cores

A "step" is one full autoregressive rollout (KTH: 10 -> 40 frames = 2 rounds of conditioning + 10 DDIM
steps + 30-frame decode) of one batch of 32 synthetic clips per GPU.  Videos shard over ranks (weak
scaling, no data-path collective); the predicted frames are all-gathered over NCCL at the end of a step.
One JSON line is printed by rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = {"dataset": "kth", "batch_per_gpu": 32, "total_pred": 40}
METRIC = "predicted frames/sec (DDIM, KTH 64x64, 10->40 rollout)"


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": reasons, "samples": len(sm)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_baseline_port(threads=None):
    """The oracle (a CPU restatement of the reference algorithm, kind='port') on the host cores, bounded sample:
    KTH tc=10/tp=20, batch 1: ONE UNet forward (of the 10 DDIM steps) and a 6-frame decode (of 30) are timed
    and extrapolated to one round = 10 UNet forwards + 30 decodes -> 20 predicted frames."""
    import torch
    from oracle import extdm_oracle as O
    import extdm_b200  # noqa: F401
    from extdm_b200.manifest import UnetConfig, unet_manifest
    from extdm_b200.weights import synth_state_dict
    from extdm_b200 import configs
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    tc, tp = 10, 20
    sd = synth_state_dict(unet_manifest(UnetConfig("ada", tc, tp)), seed=1)
    ocfg = O.unet_config("ada", tc, tp)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 3, tp, 32, 32, generator=g)
    cf = torch.randn(1, 3, tc, 32, 32, generator=g)
    fea = torch.randn(1, 256, tc + tp, 16, 16, generator=g)
    t = torch.full((1,), 500, dtype=torch.long)
    cfg, _, _ = configs.dataset("kth")
    from extdm_b200.lfae import Generator
    fp = cfg["flow_params"]["model_params"]
    gen = Generator(num_regions=fp["num_regions"], num_channels=3, revert_axis_swap=True, **fp["generator_params"])
    gsd = {k: v for k, v in gen.state_dict().items()}
    src = torch.rand(1, 3, 64, 64, generator=g)
    flow = torch.rand(1, 32, 32, 2, generator=g) * 2 - 1
    occ = torch.rand(1, 1, 32, 32, generator=g)
    with torch.no_grad():
        O.unet_forward(O.SD(sd), ocfg, x, t, cf, fea)                      # warm-up (oneDNN primitive creation)
        t0 = time.perf_counter()
        O.unet_forward(O.SD(sd), ocfg, x, t, cf, fea)
        t_unet = time.perf_counter() - t0
        O.generator_forward_with_flow(O.SD(gsd), src, flow, occ)
        t0 = time.perf_counter()
        for _ in range(6):
            O.generator_forward_with_flow(O.SD(gsd), src, flow, occ)
        t_dec = (time.perf_counter() - t0) / 6
    t_round = 10 * t_unet + (tc + tp) * t_dec
    return {"value": tp / t_round, "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": f"KTH tc=10 tp=20 batch 1: 1 UNet forward ({t_unet:.2f} s) + 6 decodes ({t_dec:.3f} s each) "
                      f"timed, extrapolated to 10 forwards + 30 decodes per 20 predicted frames "
                      f"(conditioning stage excluded)"}, t_round


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for _ in range(max(1, min(args.steps, 2))):
        cb, t_round = cpu_baseline_port()
        vals.append(cb["value"])
    cb["value"] = sum(vals) / len(vals)
    line = {"metric": METRIC, "value": cb["value"], "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * WORKLOAD["total_pred"] / cb["value"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference", "config": {"workload": "KTH 64x64 ch1->3, 10->40 autoregressive rollout, batch 1 "
                                                        "on host CPU cores (bounded sample, see cpu_baseline.sample)"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=WORKLOAD["batch_per_gpu"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-kernels", action="store_true", help="also print the per-kernel time table")
    ap.add_argument("--ncu-range", action="store_true",
                    help="for `ncu --profile-from-start off`: after the warm-up, bracket ONE step with "
                         "cudaProfilerStart/Stop and exit without printing a bench line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import extdm_b200  # noqa: F401
    from extdm_b200 import configs, lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()                                                # fails loudly when the CUDA library is missing

    B, total_pred = args.batch, WORKLOAD["total_pred"]
    model, cfg = configs.build_model(WORKLOAD["dataset"], seed=1234, device=dev)
    tc, tp = model.cond_frame_num, model.pred_frame_num
    gen = torch.Generator(device="cpu").manual_seed(1000 + rank)
    clip_host = torch.rand(B, 1, tc, 64, 64, generator=gen).expand(B, 3, tc, 64, 64).contiguous()   # gray -> 3ch
    clip_dev = clip_host.to(dev)
    pin_in = clip_host.pin_memory()                                    # the step's input, in pinned host memory
    pin_out = torch.empty(B, 3, total_pred, 64, 64).pin_memory()       # the step's result
    gathered = torch.empty(world * B, 3, total_pred, 64, 64, device=dev) if world > 1 else None

    def step(host):
        """One step = one full autoregressive rollout of the batch.  host=True is the end-to-end call a user makes with
        HOST buffers: H2D copy of the conditioning clips from pinned memory, the rollout (intermediate rounds stay on
        the device), D2H copy of the predicted frames into pinned memory, stream synchronised."""
        clip = pin_in.to(dev, non_blocking=True) if host else clip_dev
        pred = configs.rollout(model, clip, total_pred)
        if world > 1:
            dist.all_gather_into_tensor(gathered, pred.contiguous())
        if host:
            pin_out.copy_(pred, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return pred

    def timed(host, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.launches()
        e0.record()
        for _ in range(n):
            step(host)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / n, (lib.launches() - l0) // max(n, 1)

    for _ in range(max(args.warmup, 3)):
        step(False)
    if args.ncu_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        sys.stderr.write("ncu range done (no bench value is reported from a profiled run)\n")
        return
    step(True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    torch.cuda.nvtx.range_push("timed")
    ms_dev, _ = timed(False, args.steps)
    torch.cuda.nvtx.range_pop()
    ms_e2e, _ = timed(True, args.steps)
    clock_info = clocks.stop() if rank == 0 else None

    # ---- kernel accounting: launches per step and the per-kernel roofline, measured live with CUDA events
    unet = model.unet
    runner = unet.runner(B, 32, 32, 16)
    dec = model.generator.decoder(B, tc + tp, 64, 64, 32, 32, True)
    n_rounds = math.ceil(total_pred / tp)
    n_ddim = model.diffusion.sampling_timesteps
    # conditioning stage (region / background / flow predictors + the cond_fea encoder), once per round
    cond_recs = [r_ for cr in model._cond_runners.values() for r_ in (cr.recA, cr.recB)]
    cond_recs += [r_.rec for k_, r_ in model.generator._runners.items() if k_[0] == "enc"]
    launches_per_round = len(runner.prologue) + n_ddim * (len(runner.step) + 2) + len(dec.rec) + \
        sum(len(r_) for r_ in cond_recs)
    gpu_launches = launches_per_round * n_rounds * args.steps
    peaks = load_peaks()
    table = {}
    top = None
    for rec_, mult in [(runner.prologue, 1), (runner.step, n_ddim), (dec.rec, 1)] + [(r_, 1) for r_ in cond_recs]:
        rec_.run()
        torch.cuda.synchronize()
        for name, meta, ms in rec_.run_timed():
            if meta.get("tf32"):
                name += "[tf32]"                          # the conditioning convolutions: fp32 operands, kind::tf32
            t = table.setdefault(name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
            t["ms"] += ms * mult
            t["flops"] += meta.get("flops", 0.0) * mult
            t["bytes"] += meta.get("bytes", 0.0) * mult
            t["n"] += mult
            if name == "extdm_conv_gemm" and (top is None or ms * mult > top[2] * top[3]):
                top = (name, meta, ms, mult)
    gemm = table["extdm_conv_gemm"]
    total_kernel_ms = sum(t["ms"] for t in table.values())
    peak_tf, peak_bw = peaks["bf16_tflops"], peaks["hbm_gbs"]
    traffic = None                                    # dram bytes per launch of the top GEMM shape, from the ncu capture
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    top_key = f"rows={top[1]['rows']} n={top[1]['n']} k={top[1]['k']} taps={top[1]['taps']}"
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(top_key)
    # The dominant kernel of a round is the tcgen05 implicit-GEMM (conv_gemm_kernel / conv_halo_kernel templates,
    # one C-ABI entry point): `achieved` is algorithmic FLOPs of ALL its launches in a round over their summed
    # CUDA-event durations; `top_launch` is its single most expensive shape; `other_kernels` are the HBM-bound ones.
    others = {}
    if "extdm_conv_gemm[tf32]" in table:
        t = table["extdm_conv_gemm[tf32]"]
        tf = t["flops"] / (t["ms"] * 1e-3) / 1e12
        others["extdm_conv_gemm[tf32]"] = {"bound": "tensor", "achieved": tf, "peak": peak_tf / 2, "unit": "TFLOP/s",
                                           "frac": tf / (peak_tf / 2), "peak_source": "half the measured bf16 peak "
                                           "(tf32 runs at half the bf16 rate)",
                                           "share_of_kernel_time": t["ms"] / total_kernel_ms}
    for name, t in table.items():
        if not name.startswith("extdm_conv_gemm") and t["bytes"] > 0:
            gbs = t["bytes"] / (t["ms"] * 1e-3) / 1e9
            others[name] = {"bound": "hbm", "achieved": gbs, "peak": peak_bw, "unit": "GB/s", "frac": gbs / peak_bw,
                            "share_of_kernel_time": t["ms"] / total_kernel_ms}
            if name in ("extdm_stw_fused", "extdm_temporal_fused", "extdm_window_attention"):
                # fused attention layers: the HBM figure is their algorithmic traffic (read x, write y) over time; what
                # limits them is instruction issue + MUFU in the score / softmax phase
                others[name]["limited_by"] = "instruction issue / MUFU (softmax), not HBM"
    roofline = {
        "bound": "tensor", "kernel": "extdm_conv_gemm (tcgen05 implicit GEMM: conv_gemm_kernel / conv_halo_kernel), all "
                                     f"{gemm['n']} launches of one round",
        "achieved": gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12 / peak_tf, "peak_source": peaks["source"] + " burst",
        "share_of_kernel_time": gemm["ms"] / total_kernel_ms,
        "traffic": traffic,
        "top_launch": {"shape": top_key, "achieved": top[1]["flops"] / (top[2] * 1e-3) / 1e12,
                       "frac": top[1]["flops"] / (top[2] * 1e-3) / 1e12 / peak_tf,
                       "algorithmic_bytes": top[1]["bytes"], "ms": top[2], "launches_per_round": top[3]},
        "other_kernels": others,
    }
    if args.profile_kernels and rank == 0:
        for name, t in sorted(table.items(), key=lambda kv: -kv[1]["ms"]):
            sys.stderr.write(f"{name:28s} {t['ms']:9.3f} ms/round  {t['n']:5d} launches  "
                             f"{t['flops'] / max(t['ms'], 1e-9) / 1e9:8.1f} TFLOP/s\n")

    frames = world * B * total_pred
    line = {
        "metric": METRIC, "value": frames / (ms_dev * 1e-3), "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"KTH 64x64 ch1->3, 10->40 autoregressive rollout (2 rounds x 10 DDIM steps, eta=1, "
                               f"dynamic threshold), batch {B} per GPU, random-init LFAE + DM (ada UNet3D)",
                   "cache": "working set per step (>10 GB of activations) exceeds the 126 MB L2; no explicit flush",
                   "parallelism": f"dp{world} (videos sharded, one all_gather of predicted frames per step)"},
        "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": "frames/s",
                "h2d_bytes_per_step": pin_in.numel() * 4, "d2h_bytes_per_step": pin_out.numel() * 4,
                "ms_per_step": ms_e2e},
        "gpu_launches": gpu_launches,
        "roofline": roofline,
        "clocks": clock_info,
        "kernel_time_ms_per_round": {k: round(v["ms"], 3) for k, v in table.items()},
    }
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"], _ = cpu_baseline_port()
            except Exception as e:                           # the baseline is reported-only: never fail the bench on it
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {e}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
