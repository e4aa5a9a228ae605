"""FlowDiffusion: the reference's pipeline wrapper (model/BaseDM_adaptor/VideoFlowDiffusion_multi_w_ref.py,
VideoFlowDiffusion_multi1248.py, VideoFlowDiffusion_multi.py) with the same constructor, attributes and
`sample_one_video(cond_scale, real_vid) -> dict`.

  (A) conditioning  : torch (RegionPredictor / BGMotionPredictor / Generator.forward), batched over the tc frames
  (B) denoise       : GaussianDiffusion.sample  -> CUDA-graphed UNet3D + DDIM kernels
  (C) decode        : Generator.decode_video    -> CUDA warp / blend / conv kernels
"""
import torch
import torch.nn.functional as F
from torch import nn

from .diffusion import GaussianDiffusion
from .lfae import BGMotionPredictor, CondRunner, Generator, RegionPredictor
from .unet import Unet3D

_DEFAULT_UNET = {
    "w_ref": "DenoiseNet_STWAtt_w_w_ref_adaptor_cross_multi_traj_ada",
    "multi1248": "DenoiseNet_STWAtt_w_wo_ref_adaptor_cross_multi",
    "multi": "DenoiseNet_STWAtt_w_wo_ref_adaptor_cross_multi",
}


class FlowDiffusion(nn.Module):
    WRAPPER = "w_ref"            # "w_ref" | "multi1248" | "multi"
    native_conditioning = True   # class-level switch: False runs the conditioning stage as torch modules (tests A/B)

    def __init__(self, config="", pretrained_pth="", is_train=True, ddim_sampling_eta=1.0, timesteps=1000,
                 dim_mults=None, learn_null_cond=False, use_deconv=True, padding_mode="zeros", withFea=True,
                 Unet3D_architecture=None):
        super().__init__()
        kind = self.WRAPPER
        if dim_mults is None:
            dim_mults = (1, 2, 4, 8) if kind == "multi1248" else (1, 2, 4, 4)
        if Unet3D_architecture is None or kind == "multi1248":      # multi1248 imports the base UNet unconditionally
            Unet3D_architecture = _DEFAULT_UNET[kind]
        flow_params = config["flow_params"]["model_params"]
        diffusion_params = config["diffusion_params"]["model_params"]
        dataset_params = config["dataset_params"]
        self.estimate_occlusion_map = \
            flow_params["generator_params"]["pixelwise_flow_predictor_params"]["estimate_occlusion_map"]
        self.use_residual_flow = diffusion_params["use_residual_flow"]
        self.only_use_flow = diffusion_params["only_use_flow"]
        self.withFea = withFea
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        ckpt = torch.load(pretrained_pth, map_location=dev) if pretrained_pth != "" else None

        self.generator = Generator(num_regions=flow_params["num_regions"], num_channels=flow_params["num_channels"],
                                   revert_axis_swap=flow_params["revert_axis_swap"],
                                   **flow_params["generator_params"]).to(dev)
        self.region_predictor = RegionPredictor(num_regions=flow_params["num_regions"],
                                                num_channels=flow_params["num_channels"],
                                                estimate_affine=flow_params["estimate_affine"],
                                                **flow_params["region_predictor_params"]).to(dev)
        self.bg_predictor = BGMotionPredictor(num_channels=flow_params["num_channels"],
                                              **flow_params["bg_predictor_params"]).to(dev)
        if ckpt is not None:
            # VideoFlowDiffusion_multi_w_ref.py:51 (and _u22) load the generator with strict=False; multi1248 / multi
            # (VideoFlowDiffusion_multi1248.py:50) load it strictly
            self.generator.load_state_dict(ckpt["generator"], strict=kind != "w_ref")
            self.region_predictor.load_state_dict(ckpt["region_predictor"])
            self.bg_predictor.load_state_dict(ckpt["bg_predictor"])
        for m in (self.generator, self.region_predictor, self.bg_predictor):
            m.eval()
            for p in m.parameters():
                p.requires_grad = False

        tc = dataset_params["train_params"]["cond_frames"]
        tp = dataset_params["train_params"]["pred_frames"]
        from .manifest import UNET_ARCHITECTURES
        if Unet3D_architecture not in UNET_ARCHITECTURES:
            # the reference silently constructs NotImplementedError() and then dies with a NameError
            # (VideoFlowDiffusion_multi_w_ref.py:71-80); fail with a message instead
            raise NotImplementedError(f"unknown Unet3D architecture {Unet3D_architecture!r}")
        # base and ada_u22 feed the 3-channel flow volume to init_conv directly (VideoFlowDiffusion_multi1248.py:72,
        # VideoFlowDiffusion_multi_w_ref_u22.py:199-201); the others go through init_noise_conv (3 -> 256) first
        base = UNET_ARCHITECTURES[Unet3D_architecture] in ("base", "u22")
        self.unet = Unet3D(dim=64, channels=3 + 256 if base else 256 + 256, out_grid_dim=2, out_conf_dim=1,
                           dim_mults=dim_mults, use_bert_text_cond=False, learn_null_cond=learn_null_cond,
                           use_final_activation=False, use_deconv=use_deconv, padding_mode=padding_mode,
                           cond_num=tc, pred_num=tp, architecture=Unet3D_architecture).to(dev)
        self.diffusion = GaussianDiffusion(
            self.unet, image_size=dataset_params["frame_shape"] // 2, num_frames=tc + tp,
            sampling_timesteps=diffusion_params["sampling_timesteps"], timesteps=timesteps,
            loss_type=diffusion_params["loss_type"], use_dynamic_thres=True,
            null_cond_prob=diffusion_params["null_cond_prob"], ddim_sampling_eta=ddim_sampling_eta).to(dev)
        self.cond_frame_num, self.pred_frame_num, self.frame_num = tc, tp, tc + tp
        self._cond_runners = {}
        for m in (self.generator, self.region_predictor, self.bg_predictor):
            m.register_load_state_dict_post_hook(lambda mod, k: self._cond_runners.clear())
        self.is_train = is_train
        if is_train:
            raise NotImplementedError("training (FlowDiffusion.forward / p_losses) is outside the sampling hot path")

    # ------------------------------------------------------------------ (A) conditioning, torch
    @torch.no_grad()
    def condition(self, real_vid, with_decode=False):
        """real_vid (B,3,tc,H,W) in [0,1] -> dict with x_cond, cond_fea and the 'real_*' entries of the reference
        result dict (VideoFlowDiffusion_multi_w_ref.py:231-278 / VideoFlowDiffusion_multi1248.py:221-264).
        The tc per-frame passes of the reference are batched into one pass over B*tc images."""
        B, _, tc, H, W = real_vid.shape
        assert tc == self.cond_frame_num
        tp = self.pred_frame_num
        ref = real_vid[:, :, tc - 1]
        frames = real_vid.permute(0, 2, 1, 3, 4).reshape(B * tc, 3, H, W).contiguous(memory_format=torch.channels_last)
        pfp = self.generator.pixelwise_flow_predictor
        if real_vid.is_cuda and self.native_conditioning and not with_decode and \
                CondRunner.supported(self.region_predictor, self.bg_predictor, pfp):
            # (A) on the CUDA kernels: tf32 tcgen05 convolutions + fp32 element-wise kernels (lfae.CondRunner)
            key = (real_vid.device, B, tc, H, W)
            if key not in self._cond_runners:
                self._cond_runners[key] = CondRunner(self.region_predictor, self.bg_predictor, pfp, real_vid.device,
                                                     B, tc, H, W)
            grid, conf = self._cond_runners[key].run(real_vid.float())
            ret = {"real_vid_grid": grid.clone()}
            if self.estimate_occlusion_map:
                ret["real_vid_conf"] = conf.clone()
            elif self.WRAPPER != "w_ref":
                raise KeyError("occlusion_map")
        else:
            ret = self._condition_torch(real_vid, ref, frames, with_decode)
        # bottleneck features: encoder of frames 0..tc-2, then the reference frame's repeated
        enc_frames = self.generator.forward_bottle(frames).reshape(B, tc, 256, H // 4, W // 4)
        ref_fea = enc_frames[:, tc - 1]
        n_rep = (1 + tp) if self.WRAPPER == "w_ref" else tp
        fea = torch.cat([enc_frames[:, :tc - 1], ref_fea[:, None].expand(B, n_rep, *ref_fea.shape[1:])], dim=1)
        fea = fea.transpose(1, 2).contiguous()                                       # (B, 256, T', h, w)
        if self.WRAPPER != "w_ref" and not fea.is_cuda:
            # VideoFlowDiffusion_multi1248.py:243-245 resizes cond_fea to the flow resolution before the UNet.  On the
            # CUDA path the UNet prologue does that resize itself (extdm_bilinear_resize_cl, align_corners=False like
            # F.interpolate), so the features stay at H/4 here; this branch serves the CPU parity fixtures only.
            n, c, t, h, w = fea.shape
            hw = ret["real_vid_grid"].shape[-2:]
            fea = F.interpolate(fea.transpose(1, 2).reshape(n * t, c, h, w), size=hw, mode="bilinear")
            fea = fea.reshape(n, t, c, *hw).transpose(1, 2).contiguous()
        grid = ret["real_vid_grid"]
        if self.estimate_occlusion_map:
            x_cond = torch.cat((grid, ret["real_vid_conf"] * 2 - 1), dim=1)
        else:
            x_cond = torch.cat((grid, torch.zeros_like(grid)[:, 0:1]), dim=1)
        return ret, x_cond, fea, ref

    def _condition_torch(self, real_vid, ref, frames, with_decode):
        """The same stage as ordinary torch modules (cuDNN on a GPU): CPU parity fixtures, unsupported predictor
        hyper-parameters, and the A/B partner of the native path in the tests."""
        B, _, tc, H, W = real_vid.shape
        ref_rep = ref.repeat_interleave(tc, dim=0)
        src_params = self.region_predictor(ref)
        src_rep = {k: v.repeat_interleave(tc, dim=0) for k, v in src_params.items()}
        drv_params = self.region_predictor(frames)
        bg = self.bg_predictor(ref_rep, frames)
        # Generator.forward = flow predictor + the same warp/blend decode as forward_with_flow (generator.py:105-144 vs
        # :152-206): only the flow predictor runs here; the decoded conditioning frames ('real_out_vid',
        # 'real_warped_vid') are the first tc frames of the decode in sample_one_video, which the reference
        # computes a second time from identical inputs (VideoFlowDiffusion_multi_w_ref.py:295-306).
        gen = self.generator.pixelwise_flow_predictor(source_image=ref_rep, driving_region_params=drv_params,
                                                      source_region_params=src_rep, bg_params=bg)
        per_frame = lambda t: t.reshape(B, tc, *t.shape[1:]).transpose(1, 2)       # (B*tc, C, ..) -> (B, C, tc, ..)
        ret = {"real_vid_grid": per_frame(gen["optical_flow"].permute(0, 3, 1, 2)).contiguous()}
        if self.estimate_occlusion_map:
            ret["real_vid_conf"] = per_frame(gen["occlusion_map"]).contiguous()
        elif self.WRAPPER != "w_ref":
            raise KeyError("occlusion_map")         # VideoFlowDiffusion_multi1248.py:236 without estimate_occlusion_map
        if with_decode:
            # the reference's own torch decode of the conditioning frames (used by the CPU parity test only)
            full = self.generator(ref_rep, source_region_params=src_rep, driving_region_params=drv_params, bg_params=bg)
            ret["real_out_vid"] = per_frame(full["prediction"]).contiguous()
            ret["real_warped_vid"] = per_frame(full["deformed"]).contiguous()
        return ret

    # ------------------------------------------------------------------ full round
    @torch.no_grad()
    def sample_one_video(self, cond_scale, real_vid, noise=None):
        ret, x_cond, cond_fea, ref = self.condition(real_vid)
        pred = self.diffusion.sample(x_cond, cond_fea=cond_fea, batch_size=1, cond_scale=cond_scale, noise=noise)
        tc = self.cond_frame_num
        grid = pred[:, :2]
        if self.use_residual_flow:
            grid = grid + self.get_grid(grid.shape[0], 1, grid.shape[3], grid.shape[4]).to(grid.device)
        sample_vid_grid = torch.cat([ret["real_vid_grid"][:, :, :tc], grid], dim=2)
        sample_vid_conf = None
        if self.estimate_occlusion_map:
            sample_vid_conf = torch.cat([ret["real_vid_conf"][:, :, :tc], (pred[:, 2:3] + 1) * 0.5], dim=2)
        out, warped = self.generator.decode_video(ref, sample_vid_grid, sample_vid_conf)
        ret["sample_vid_grid"] = sample_vid_grid
        if sample_vid_conf is not None:
            ret["sample_vid_conf"] = sample_vid_conf
        ret["sample_out_vid"] = out
        ret["sample_warped_vid"] = warped
        ret["real_out_vid"] = out[:, :, :tc]
        ret["real_warped_vid"] = warped[:, :, :tc]
        return ret

    def forward(self, real_vid):
        raise NotImplementedError("training forward is outside the sampling hot path (SURVEY.md section 2, row 3a)")

    @staticmethod
    def get_grid(b, nf, H, W, normalize=True):
        if normalize:
            hr, wr = torch.linspace(-1, 1, H), torch.linspace(-1, 1, W)
        else:
            hr, wr = torch.arange(0, H), torch.arange(0, W)
        grid = torch.stack(torch.meshgrid(hr, wr, indexing="xy"), -1).repeat(b, 1, 1, 1).flip(3).float()
        return grid.permute(0, 3, 1, 2).unsqueeze(dim=2).repeat(1, 1, nf, 1, 1)


class FlowDiffusionMulti1248(FlowDiffusion):
    WRAPPER = "multi1248"


class FlowDiffusionU22(FlowDiffusion):
    """VideoFlowDiffusion_multi_w_ref_u22.py:142-236,415-510: the w_ref pipeline around the ada_u22 UNet.  The
    reference spreads LFAE and DM over `device_ids` (model parallel, one video batch at a time); here the whole
    round runs on one GPU and videos shard across ranks (sharding.py), so `device_ids` is accepted and ignored."""

    def __init__(self, config="", pretrained_pth="", is_train=True, ddim_sampling_eta=1.0, timesteps=1000,
                 dim_mults=(1, 2, 4, 4), learn_null_cond=False, use_deconv=True, padding_mode="zeros", withFea=True,
                 Unet3D_architecture="DenoiseNet_STWAtt_w_w_ref_adaptor_cross_multi_traj_ada_u22", device_ids=None):
        super().__init__(config, pretrained_pth, is_train, ddim_sampling_eta, timesteps, dim_mults, learn_null_cond,
                         use_deconv, padding_mode, withFea, Unet3D_architecture)


class FlowDiffusionMulti(FlowDiffusion):
    WRAPPER = "multi"


WRAPPERS = {
    "VideoFlowDiffusion_multi_w_ref": FlowDiffusion,
    "VideoFlowDiffusion_multi1248": FlowDiffusionMulti1248,
    "VideoFlowDiffusion_multi": FlowDiffusionMulti,
    "VideoFlowDiffusion_multi_w_ref_u22": FlowDiffusionU22,
}


def flow_diffusion_class(dm_arch):
    """`--DM_arch` string (scripts/DM/valid.py:83-92) -> class."""
    if dm_arch not in WRAPPERS:
        raise NotImplementedError(f"unknown DM architecture {dm_arch!r}")
    return WRAPPERS[dm_arch]
