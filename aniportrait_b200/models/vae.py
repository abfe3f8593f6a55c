"""AutoencoderKL (sd-vae-ft-mse layout) with diffusers' parameter names so published checkpoints load.

Role on the hot path: `decode` of the denoised latents (reference pipeline_pose2vid_long.py:113-126, one frame per
call) and a single `encode` of the reference image. Round-1 status: both run as batched fp16 torch/cuDNN library ops
(decode is batched over frames instead of frame-at-a-time); the sm_100a conv / GroupNorm kernels already cover every
decoder layer shape and replacing this module's forward with them is the first "next" row.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch import nn

from .modeling import ModelBase


class _Resnet(nn.Module):
    def __init__(self, cin, cout, groups=32, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (self.conv_shortcut(x) if self.conv_shortcut is not None else x) + h


class _Attn(nn.Module):
    def __init__(self, c, groups=32, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(t)[:, None], self.to_k(t)[:, None], self.to_v(t)[:, None]
        a = F.scaled_dot_product_attention(q, k, v)[:, 0]
        return x + self.to_out[0](a).transpose(1, 2).reshape(b, c, h, w)


class _Mid(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.attentions = nn.ModuleList([_Attn(c, groups)])
        self.resnets = nn.ModuleList([_Resnet(c, c, groups), _Resnet(c, c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _Sampler(nn.Module):
    def __init__(self, c, down):
        super().__init__()
        self.down = down
        self.conv = nn.Conv2d(c, c, 3, stride=2 if down else 1, padding=0 if down else 1)

    def forward(self, x):
        if self.down:
            return self.conv(F.pad(x, (0, 1, 0, 1)))
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Stage(nn.Module):
    def __init__(self, cin, cout, layers, groups, sampler):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        if sampler == "down":
            self.downsamplers = nn.ModuleList([_Sampler(cout, True)])
        elif sampler == "up":
            self.upsamplers = nn.ModuleList([_Sampler(cout, False)])
        self._sampler = sampler

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self._sampler == "down":
            x = self.downsamplers[0](x)
        elif self._sampler == "up":
            x = self.upsamplers[0](x)
        return x


class Encoder(nn.Module):
    def __init__(self, cin, latent, boc, layers, groups):
        super().__init__()
        self.conv_in = nn.Conv2d(cin, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        oc = boc[0]
        for i, c in enumerate(boc):
            ic, oc = oc, c
            self.down_blocks.append(_Stage(ic, oc, layers, groups, "down" if i != len(boc) - 1 else None))
        self.mid_block = _Mid(boc[-1], groups)
        self.conv_norm_out = nn.GroupNorm(groups, boc[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[-1], 2 * latent, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(self.mid_block(x))))


class Decoder(nn.Module):
    def __init__(self, latent, cout, boc, layers, groups):
        super().__init__()
        self.conv_in = nn.Conv2d(latent, boc[-1], 3, padding=1)
        self.mid_block = _Mid(boc[-1], groups)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        oc = rev[0]
        for i, c in enumerate(rev):
            ic, oc = oc, c
            self.up_blocks.append(_Stage(ic, oc, layers + 1, groups, "up" if i != len(rev) - 1 else None))
        self.conv_norm_out = nn.GroupNorm(groups, boc[0], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], cout, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class _LatentDist:
    def __init__(self, moments):
        self.mean, self.logvar = torch.chunk(moments, 2, dim=1)

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        std = torch.exp(0.5 * self.logvar.clamp(-30.0, 20.0))
        return self.mean + std * torch.randn(self.mean.shape, generator=generator, device=self.mean.device,
                                             dtype=self.mean.dtype)


@dataclass
class AutoencoderKLOutput:
    latent_dist: _LatentDist


@dataclass
class DecoderOutput:
    sample: torch.Tensor


class AutoencoderKL(ModelBase):
    def __init__(self, in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
                 up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=(128, 256, 512, 512),
                 layers_per_block=2, act_fn="silu", latent_channels=4, norm_num_groups=32, sample_size=512,
                 scaling_factor=0.18215, force_upcast=True):
        super().__init__()
        self.register_to_config(in_channels=in_channels, out_channels=out_channels,
                                down_block_types=tuple(down_block_types), up_block_types=tuple(up_block_types),
                                block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                                act_fn=act_fn, latent_channels=latent_channels, norm_num_groups=norm_num_groups,
                                sample_size=sample_size, scaling_factor=scaling_factor, force_upcast=force_upcast)
        boc = list(block_out_channels)
        self.encoder = Encoder(in_channels, latent_channels, boc, layers_per_block, norm_num_groups)
        self.decoder = Decoder(latent_channels, out_channels, boc, layers_per_block, norm_num_groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)

    @torch.no_grad()
    def encode(self, x, return_dict=True):
        return AutoencoderKLOutput(latent_dist=_LatentDist(self.quant_conv(self.encoder(x))))

    @torch.no_grad()
    def decode(self, z, return_dict=True, generator=None):
        return DecoderOutput(sample=self.decoder(self.post_quant_conv(z)))
