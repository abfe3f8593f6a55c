"""AutoencoderKL restated from diffusers 0.24.0 (models/autoencoder_kl.py, models/vae.py, unet_2d_blocks.py
DownEncoderBlock2D / UpDecoderBlock2D / UNetMidBlock2D with one single-head attention) for the sd-vae-ft-mse config the
reference loads (README.md:92; scripts/pose2vid.py:59). Parameter names match the published checkpoint."""
from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch import nn

from ..configuration_utils import ConfigMixin, register_to_config
from ..utils import BaseOutput
from .modeling_utils import ModelMixin
from .resnet import Downsample2D, ResnetBlock2D, Upsample2D


class VaeAttention(nn.Module):
    """diffusers Attention(..., heads=1, residual_connection=True, bias=True, norm_num_groups=32) on 4-D input as
    run by AttnProcessor2_0."""

    def __init__(self, channels, groups=32, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps, affine=True)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])

    def forward(self, hidden_states, temb=None):
        residual = hidden_states
        b, c, h, w = hidden_states.shape
        x = hidden_states.view(b, c, h * w).transpose(1, 2)
        x = self.group_norm(x.transpose(1, 2)).transpose(1, 2)
        q = self.to_q(x).view(b, -1, 1, c).transpose(1, 2)
        k = self.to_k(x).view(b, -1, 1, c).transpose(1, 2)
        v = self.to_v(x).view(b, -1, 1, c).transpose(1, 2)
        x = F.scaled_dot_product_attention(q, k, v)
        x = x.transpose(1, 2).reshape(b, -1, c).to(q.dtype)
        x = self.to_out[0](x)
        x = x.transpose(-1, -2).reshape(b, c, h, w)
        return x + residual


class MidBlock(nn.Module):
    def __init__(self, channels, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels=channels, out_channels=channels, temb_channels=None, eps=eps, groups=groups),
            ResnetBlock2D(in_channels=channels, out_channels=channels, temb_channels=None, eps=eps, groups=groups)])
        self.attentions = nn.ModuleList([VaeAttention(channels, groups, eps)])

    def forward(self, x):
        x = self.resnets[0](x, None)
        x = self.attentions[0](x)
        return self.resnets[1](x, None)


class DownEncoderBlock(nn.Module):
    def __init__(self, cin, cout, layers, add_down, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels=cin if i == 0 else cout, out_channels=cout,
                                                    temb_channels=None, eps=eps, groups=groups)
                                      for i in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, use_conv=True, out_channels=cout, padding=0,
                                                        name="op")]) if add_down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x, None)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UpDecoderBlock(nn.Module):
    def __init__(self, cin, cout, layers, add_up, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels=cin if i == 0 else cout, out_channels=cout,
                                                    temb_channels=None, eps=eps, groups=groups)
                                      for i in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout, use_conv=True, out_channels=cout)]) if add_up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x, None)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Encoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, groups, double_z=True):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        oc = block_out_channels[0]
        for i, c in enumerate(block_out_channels):
            ic, oc = oc, c
            self.down_blocks.append(DownEncoderBlock(ic, oc, layers_per_block, i != len(block_out_channels) - 1,
                                                     groups, 1e-6))
        self.mid_block = MidBlock(block_out_channels[-1], groups, 1e-6)
        self.conv_norm_out = nn.GroupNorm(groups, block_out_channels[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[-1], 2 * out_channels if double_z else out_channels, 3,
                                  padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, groups):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[-1], 3, padding=1)
        self.mid_block = MidBlock(block_out_channels[-1], groups, 1e-6)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(block_out_channels))
        oc = rev[0]
        for i, c in enumerate(rev):
            ic, oc = oc, c
            self.up_blocks.append(UpDecoderBlock(ic, oc, layers_per_block + 1, i != len(rev) - 1, groups, 1e-6))
        self.conv_norm_out = nn.GroupNorm(groups, block_out_channels[0], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, 3, padding=1)

    def forward(self, z):
        x = self.conv_in(z)
        x = self.mid_block(x)
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    def __init__(self, parameters):
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None):
        return self.mean + self.std * torch.randn(self.mean.shape, generator=generator, dtype=self.mean.dtype)

    def mode(self):
        return self.mean


@dataclass
class AutoencoderKLOutput(BaseOutput):
    latent_dist: DiagonalGaussianDistribution = None


@dataclass
class DecoderOutput(BaseOutput):
    sample: torch.FloatTensor = None


class AutoencoderKL(ModelMixin, ConfigMixin):
    @register_to_config
    def __init__(self, in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
                 up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=(128, 256, 512, 512),
                 layers_per_block=2, act_fn="silu", latent_channels=4, norm_num_groups=32, sample_size=512,
                 scaling_factor=0.18215, force_upcast=True):
        super().__init__()
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)

    def encode(self, x, return_dict=True):
        moments = self.quant_conv(self.encoder(x))
        return AutoencoderKLOutput(latent_dist=DiagonalGaussianDistribution(moments))

    def decode(self, z, return_dict=True, generator=None):
        return DecoderOutput(sample=self.decoder(self.post_quant_conv(z)))
