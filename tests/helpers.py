"""Shared test helpers: build the product models with the seeded synthetic weights used by the golden fixtures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from aniportrait_b200.synthetic import randomize_state_dict  # noqa: E402

MOTION_KWARGS = dict(num_attention_heads=8, num_transformer_block=1,
                     attention_block_types=["Temporal_Self", "Temporal_Self"], temporal_position_encoding=True,
                     temporal_position_encoding_max_len=32, temporal_attention_dim_div=1)


def rel_l2(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def build_unet3d(chans, seed, device=None):
    from aniportrait_b200.models import UNet3DConditionModel
    m = UNet3DConditionModel(sample_size=64, block_out_channels=tuple(chans), cross_attention_dim=768,
                             attention_head_dim=8, use_inflated_groupnorm=True, unet_use_cross_frame_attention=False,
                             unet_use_temporal_attention=False, use_motion_module=True,
                             motion_module_resolutions=[1, 2, 4, 8], motion_module_mid_block=True,
                             motion_module_type="Vanilla", motion_module_kwargs=dict(MOTION_KWARGS))
    sd = randomize_state_dict(m.state_dict(), seed=seed)
    m.load_state_dict(sd)
    if device is not None:
        m = m.to(device=device, dtype=torch.float16)
    return m, sd


def build_unet2d(chans, seed, device=None):
    from aniportrait_b200.models import UNet2DConditionModel
    m = UNet2DConditionModel(sample_size=64, block_out_channels=tuple(chans), cross_attention_dim=768,
                             attention_head_dim=8)
    sd = randomize_state_dict(m.state_dict(), seed=seed)
    m.load_state_dict(sd)
    if device is not None:
        m = m.to(device=device, dtype=torch.float16)
    return m, sd


def seeded_inputs_unet3d(B, Fr, h, w, chans, seed):
    """Must stay identical to oracle/make_golden.py::seeded_inputs_unet3d."""
    g = torch.Generator().manual_seed(seed)
    sample = torch.randn(1, 4, Fr, h, w, generator=g).repeat(B, 1, 1, 1, 1)
    clip = torch.randn(1, 768, generator=g)
    ehs = torch.cat([torch.zeros_like(clip), clip], 0).unsqueeze(1) if B == 2 else clip.unsqueeze(1)
    ref_lat = torch.randn(1, 4, h, w, generator=g)
    sizes = [(chans[0], h), (chans[0], h // 2), (chans[1], h // 4), (chans[2], h // 8), (chans[3], h // 8)]
    pose = [0.5 * torch.randn(1, c, Fr, s, s * w // h, generator=g).repeat(B, 1, 1, 1, 1) for c, s in sizes]
    return sample, ehs, ref_lat, pose


# ---------------------------------------------------------------------------------------------------------------
# whole-pipeline fixtures (must stay identical to oracle/make_golden.py)
# ---------------------------------------------------------------------------------------------------------------
SCHED_KWARGS = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="linear", clip_sample=False, steps_offset=1,
                    prediction_type="v_prediction", rescale_betas_zero_snr=True, timestep_spacing="trailing")


def small_clip_encoder(seed):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    cfg = CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                           image_size=224, patch_size=32, projection_dim=768)
    torch.manual_seed(seed)
    m = CLIPVisionModelWithProjection(cfg)
    m.load_state_dict(randomize_state_dict(m.state_dict(), seed=seed))
    return m.eval()


def full_clip_encoder(seed):
    """Must stay identical to oracle/make_golden.py::full_clip_encoder (CLIP ViT-L/14 vision tower, seeded weights)."""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    cfg = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                           image_size=224, patch_size=14, projection_dim=768)
    torch.manual_seed(seed)
    m = CLIPVisionModelWithProjection(cfg)
    m.load_state_dict(randomize_state_dict(m.state_dict(), seed=seed))
    return m.eval()


def pipeline_inputs(size, L, seed):
    import numpy as np
    import PIL.Image
    rng = np.random.RandomState(seed)
    ref_image = PIL.Image.fromarray(rng.randint(0, 256, (size + 40, size + 24, 3), dtype=np.uint8))

    def pose_map(r):
        img = np.zeros((size, size, 3), dtype=np.uint8)
        for _ in range(24):
            x0, y0 = r.randint(0, size, 2)
            ln = r.randint(4, size // 3)
            col = r.randint(64, 256, 3)
            if r.rand() < 0.5:
                img[y0:y0 + 2, x0:min(size, x0 + ln)] = col
            else:
                img[y0:min(size, y0 + ln), x0:x0 + 2] = col
        return img

    poses = [pose_map(np.random.RandomState(seed + 1 + f)) for f in range(L)]
    ref_pose = pose_map(np.random.RandomState(seed + 1000))
    return ref_image, poses, ref_pose


def build_pipeline(P, device):
    """Product Pose2VideoPipeline with the seeded weights of golden case `P` (params dict of the fixture)."""
    from aniportrait_b200.models.pose_guider import PoseGuider
    from aniportrait_b200.models.vae import AutoencoderKL
    from aniportrait_b200.pipelines.pipeline_pose2vid_long import Pose2VideoPipeline
    from aniportrait_b200.pipelines.scheduler import DDIMScheduler
    sd = P["seeds"]
    unet3d, _ = build_unet3d(P["chans"], sd["unet3d"])
    unet2d, _ = build_unet2d(P["chans"], sd["unet2d"])
    pose = PoseGuider(P["chans"][0])
    pose.load_state_dict(randomize_state_dict(pose.state_dict(), seed=sd["pose"]))
    vae = AutoencoderKL(block_out_channels=tuple(P["vae_chans"]))
    vae.load_state_dict(randomize_state_dict(vae.state_dict(), seed=sd["vae"]))
    clip = full_clip_encoder(sd["clip"]) if P.get("clip") == "vit_l_14" else small_clip_encoder(sd["clip"])
    pipe = Pose2VideoPipeline(vae=vae, image_encoder=clip, reference_unet=unet2d, denoising_unet=unet3d,
                              pose_guider=pose, scheduler=DDIMScheduler(**SCHED_KWARGS))
    return pipe.to(device, dtype=torch.float16)
