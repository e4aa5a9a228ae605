"""smoke(): one small invocation of the hot path on cuda:0, checked against the oracle (CPU fp32)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def smoke():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import extdm_oracle as O            # checker only
    from .manifest import UnetConfig, unet_manifest
    from .unet import Unet3D
    from .diffusion import GaussianDiffusion
    from .weights import synth_state_dict

    tc, tp, B = 2, 5, 1
    cfg = UnetConfig("ada", tc, tp)
    sd = synth_state_dict(unet_manifest(cfg), seed=11)
    unet = Unet3D(dim=64, channels=512, dim_mults=(1, 2, 4, 4), cond_num=tc, pred_num=tp).cuda()
    unet.load_state_dict(sd, strict=False)
    diff = GaussianDiffusion(unet, image_size=32, num_frames=tc + tp, sampling_timesteps=2, timesteps=1000,
                             loss_type="l2", null_cond_prob=0.0).cuda()
    g = torch.Generator().manual_seed(7)
    x_cond = torch.randn(B, 3, tc, 32, 32, generator=g) * 0.5
    cond_fea = (torch.randn(B, 256, tc + tp, 16, 16, generator=g) * 0.5).abs()
    noise = torch.randn(2, B, 3, tp, 32, 32, generator=g)
    out = diff.sample(x_cond.cuda(), cond_fea=cond_fea.cuda(), noise=noise.cuda()).cpu()
    full = {"denoise_fn." + k: v for k, v in sd.items()}
    ocfg = O.unet_config("ada", tc, tp)
    with torch.no_grad():
        ref = O.ddim_sample(O.SD(full), ocfg, x_cond, cond_fea, noise[0], [noise[1], None], sampling=2)
    rel = ((out - ref).norm() / ref.norm()).item()
    print(f"[smoke] DDIM(2 steps) UNet3D 'ada' tc={tc} tp={tp}: rel-L2 vs oracle = {rel:.3e}")
    assert rel < 3e-2, rel

    # ---- decode: Generator.forward_with_flow (warp / occlusion blend / conv decoder) on the same latent flow
    from . import configs
    from .lfae import Generator
    fp = configs.dataset("kth")[0]["flow_params"]["model_params"]
    gen = Generator(num_regions=fp["num_regions"], num_channels=3, revert_axis_swap=True, **fp["generator_params"]).eval()
    base = gen.state_dict()
    gsd = synth_state_dict({k: tuple(v.shape) for k, v in base.items()}, seed=21, base=base)
    gen.load_state_dict(gsd)
    gen = gen.cuda()
    src = torch.rand(B, 3, 64, 64, generator=g)
    ident = torch.stack(torch.meshgrid(torch.linspace(-1, 1, 32), torch.linspace(-1, 1, 32), indexing="xy"), -1)
    flow = ident[None] + 0.1 * out[:, :2, 0].permute(0, 2, 3, 1).clamp(-1, 1)
    occ = ((out[:, 2:3, 0] + 1) * 0.5).clamp(0, 1)
    got = gen.forward_with_flow(src.cuda(), flow.cuda(), occ.cuda())
    with torch.no_grad():
        want = O.generator_forward_with_flow(O.SD(gsd), src, flow, occ)
    mse = ((got["prediction"].cpu() - want["prediction"]) ** 2).mean().item()
    warp_err = (got["deformed"].cpu() - want["deformed"]).abs().max().item()
    import math
    psnr = 99.0 if mse == 0 else 10 * math.log10(1.0 / mse)
    print(f"[smoke] forward_with_flow 64x64: decoded frame PSNR vs oracle = {psnr:.1f} dB, warped image max-abs {warp_err:.1e}")
    assert psnr >= 35.0 and warp_err <= 2e-5, (psnr, warp_err)
    return rel
