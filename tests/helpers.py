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
