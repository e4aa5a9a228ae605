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
    return rel
