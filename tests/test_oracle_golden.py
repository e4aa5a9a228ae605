"""CPU: the oracle (oracle/extdm_oracle.py) against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  fp32 vs fp32, tolerance 1e-4 abs on O(1..10) values (op re-association only)."""
import os
import sys

import pytest
import torch

import extdm_b200  # noqa: F401
from extdm_b200.weights import synth_state_dict
from oracle import extdm_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rnd(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def unet_inputs(variant, tc, tp, B, seed):
    tm = tc - 1 if variant == "base" else tc
    fea_hw = 32 if variant == "base" else 16
    return dict(x=rnd((B, 3, tp, 32, 32), seed + 1), cond_frames=rnd((B, 3, tc, 32, 32), seed + 2, 0.5),
                cond_fea=rnd((B, 256, tm + tp, fea_hw, fea_hw), seed + 3, 0.5).abs(),
                time=torch.full((B,), 545, dtype=torch.long))


@pytest.mark.parametrize("name", ["unet_ada_c2p5", "unet_u12_c2p3", "unet_base_c3p2", "unet_u22_c2p5", "unet_ada_c10p20"])
def test_unet_forward(name):
    torch.set_num_threads(8)
    fx = torch.load(os.path.join(GOLD, name + ".pt"))
    sd = synth_state_dict(fx["manifest"], fx["weight_seed"])
    cfg = O.unet_config(fx["variant"], fx["tc"], fx["tp"], dim_mults=fx["dim_mults"])
    inp = unet_inputs(fx["variant"], fx["tc"], fx["tp"], fx["B"], fx["input_seed"])
    assert abs(float(sum(v.double().sum() for v in inp.values())) - fx["input_checksum"]) < 1e-3, "RNG drift"
    with torch.no_grad():
        out = O.unet_forward(O.SD(sd), cfg, inp["x"], inp["time"], inp["cond_frames"], inp["cond_fea"])
    assert (out - fx["out"]).abs().max().item() < 1e-4


def test_ddim_sample():
    torch.set_num_threads(8)
    fx = torch.load(os.path.join(GOLD, "ddim_ada_c2p5.pt"))
    sd = {"denoise_fn." + k: v for k, v in synth_state_dict(fx["manifest"], fx["weight_seed"]).items()}
    cfg = O.unet_config("ada", fx["tc"], fx["tp"])
    inp = unet_inputs("ada", fx["tc"], fx["tp"], fx["B"], fx["input_seed"])
    noises = [rnd((fx["B"], 3, fx["tp"], 32, 32), fx["noise_seed"] + i) for i in range(fx["sampling"])]
    tab = O.cosine_schedule_tables()
    for k, v in fx["tables"].items():
        assert torch.equal(tab[k], v), k
    with torch.no_grad():
        out = O.ddim_sample(O.SD(sd), cfg, inp["cond_frames"], inp["cond_fea"], noises[0], noises[1:] + [None],
                            sampling=fx["sampling"])
    assert (out - fx["out"]).abs().max().item() < 2e-4


def test_generator_forward_with_flow():
    fx = torch.load(os.path.join(GOLD, "generator_fwf.pt"))
    sd = synth_state_dict(fx["manifest"], fx["weight_seed"])
    B = fx["B"]
    src = torch.rand((B, 3, 64, 64), generator=torch.Generator().manual_seed(400))
    ident = torch.stack(torch.meshgrid(torch.linspace(-1, 1, 32), torch.linspace(-1, 1, 32), indexing="xy"), -1)
    flow = ident[None] + rnd((B, 32, 32, 2), 401, 0.15)
    occ = torch.rand((B, 1, 32, 32), generator=torch.Generator().manual_seed(402))
    with torch.no_grad():
        a = O.generator_forward_with_flow(O.SD(sd), src, flow, occ)
        b = O.generator_forward_with_flow(O.SD(sd), src, flow, None)
    assert (a["prediction"] - fx["prediction"]).abs().max().item() < 1e-5
    assert (a["deformed"] - fx["deformed"]).abs().max().item() < 1e-6
    assert (b["prediction"] - fx["prediction_noocc"]).abs().max().item() < 1e-6


def test_quantile_rank_is_fp32():
    """the interpolation weight comes from an fp32 rank (0.0996.., not 0.1)."""
    x = torch.rand(2, 15360, generator=torch.Generator().manual_seed(3))
    _, s = O.dynamic_threshold(x * 3 - 1.5)
    srt = (x * 3 - 1.5).abs().sort(dim=1).values
    pos = torch.tensor(0.9, dtype=torch.float32) * torch.tensor(15359, dtype=torch.float32)
    lo = int(pos.floor())
    w = pos - lo
    ref = torch.lerp(srt[:, lo], srt[:, lo + 1], w).clamp(min=1.0)
    assert torch.equal(s.flatten(), ref)


@pytest.mark.parametrize("h,H", [(32, 64), (32, 32), (32, 16)])
def test_warp_index_oracle_vs_aten(h, H):
    """The numpy fp32 index-math oracle (what the CUDA warp kernels are compared with bit-for-bit) against ATen's
    F.interpolate + F.grid_sample on CPU: identical up to the 1-ulp FMA-contraction difference of the flow resize."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    flow = (torch.rand(2, h, h, 2, generator=g) * 2 - 1) * 1.2
    occ = torch.rand(2, 1, h, h, generator=g)
    src = torch.rand(2, 3, H, H, generator=g)
    xy, wts, gf = O.warp_index_math(flow, occ, H, H)
    fl = F.interpolate(flow.permute(0, 3, 1, 2), size=(H, H), mode="bilinear").permute(0, 2, 3, 1) if h != H else flow
    oc = F.interpolate(occ, size=(H, H), mode="bilinear") if h != H else occ
    assert (gf[..., :2] - fl).abs().max().item() <= 2.5e-7
    assert (gf[..., 2] - oc[:, 0]).abs().max().item() <= 2.5e-7
    ref = F.grid_sample(src, fl, align_corners=True)
    got = O.warp_from_taps(src, xy, wts)
    assert (got - ref).abs().max().item() <= (2e-5 if h != H else 5e-7)
