"""Shared helpers of the end-to-end rollout parity tests (tests/golden/rollout_*.pt, written by
tests/golden/make_golden.py::gen_rollout from the UNMODIFIED reference)."""
import math
import os

import torch
import torch.nn.functional as F

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

ROLLOUTS = ["rollout_smmnist", "rollout_bair", "rollout_ucf", "rollout_cityscapes", "rollout_cityscapes_u22"]

UNET_ARCH = {
    "ada": "DenoiseNet_STWAtt_w_w_ref_adaptor_cross_multi_traj_ada",
    "u12": "DenoiseNet_STWAtt_w_w_ref_adaptor_cross_multi_traj_u12",
    "base": "DenoiseNet_STWAtt_w_wo_ref_adaptor_cross_multi",
    "u22": "DenoiseNet_STWAtt_w_w_ref_adaptor_cross_multi_traj_ada_u22",
}


def rnd(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def load(name):
    return torch.load(os.path.join(GOLD, name + ".pt"))


def smooth_clip(B, tc, hw, seed, gray):
    """= tests/golden/make_golden.py::smooth_clip (seeded on the CPU generator)."""
    c = 1 if gray else 3
    coarse = torch.rand((B, c, 4, 8, 8), generator=torch.Generator().manual_seed(seed))
    vid = F.interpolate(coarse, size=(tc, hw, hw), mode="trilinear", align_corners=True).clamp(0, 1)
    return vid.expand(B, 3, tc, hw, hw).contiguous()


def round_noise(fx, r):
    """(steps, B, 3, tp, 32, 32): the Gaussian tensors round r of the fixture consumed, in draw order."""
    return torch.stack([rnd((fx["B"], 3, fx["tp"], 32, 32), fx["noise_seed"] + 100 * r + i)
                        for i in range(fx["steps"])])


def build_model(fx, device):
    """This repo's FlowDiffusion for the fixture's (config, wrapper, UNet) with the fixture's synthetic weights."""
    import extdm_b200  # noqa: F401
    from extdm_b200.flow_diffusion import flow_diffusion_class
    from extdm_b200.weights import synth_state_dict
    fd = flow_diffusion_class(fx["dm_arch"])(config=fx["cfg"], pretrained_pth="", is_train=False,
                                             Unet3D_architecture=UNET_ARCH[fx["variant"]]).eval()
    for part, seed in fx["weight_seeds"].items():
        m = getattr(fd, part)
        base = m.state_dict()
        man = fx["manifests"].get(part) or {k: tuple(v.shape) for k, v in base.items()}
        m.load_state_dict(synth_state_dict(man, seed, base=base), strict=True)
    return fd.to(device)


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def psnr(a, b):
    mse = ((a.float() - b.float()) ** 2).mean().item()
    return 99.0 if mse == 0 else 10 * math.log10(1.0 / mse)
