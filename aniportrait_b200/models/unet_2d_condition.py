"""ReferenceNet (mirror of the reference's src/models/unet_2d_condition.py + unet_2d_blocks.py + transformer_2d.py):
an SD1.5 UNet2DConditionModel WITHOUT the output head (conv_norm_out / conv_out removed, reference
unet_2d_condition.py:645-653,1295-1299), run once per video at t=0 so that every BasicTransformerBlock (write mode)
records norm1(hidden_states) into its bank. Same kernels as the denoising UNet, frames = batch.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple, Union

import torch
from torch import nn

from .. import ops
from .blocks import Downsample, ResnetBlock, RunCtx, Transformer2DModel, Upsample
from .modeling import ModelBase, PackedCache, f32
from .unet_3d import TimestepEmbedding


@dataclass
class UNet2DConditionOutput:
    sample: torch.Tensor = None


class _Block2D(nn.Module):
    def __init__(self, resnet_io, out_channels, heads, cross_attention_dim, has_attn, groups, eps, sampler,
                 temb_channels):
        super().__init__()
        self.has_cross_attention = has_attn
        self.resnets = nn.ModuleList([ResnetBlock(i, o, temb_channels, groups, eps) for i, o in resnet_io])
        if has_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(heads, out_channels // heads, out_channels,
                                                                cross_attention_dim, groups) for _ in resnet_io])
        self.downsamplers = nn.ModuleList([Downsample(out_channels)]) if sampler == "down" else None
        self.upsamplers = nn.ModuleList([Upsample(out_channels)]) if sampler == "up" else None

    def layer(self, j, x, ctx, skip=None):
        x = self.resnets[j].run(x, ctx, skip)
        if self.has_cross_attention:
            x = self.attentions[j].run(x, ctx)
        return x


class CrossAttnDownBlock2D(_Block2D):
    pass


class DownBlock2D(_Block2D):
    pass


class CrossAttnUpBlock2D(_Block2D):
    pass


class UpBlock2D(_Block2D):
    pass


class UNetMidBlock2DCrossAttn(nn.Module):
    def __init__(self, channels, heads, cross_attention_dim, groups, eps, temb_channels):
        super().__init__()
        self.has_cross_attention = True
        self.resnets = nn.ModuleList([ResnetBlock(channels, channels, temb_channels, groups, eps) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, channels // heads, channels, cross_attention_dim,
                                                            groups)])

    def run(self, x, ctx):
        x = self.resnets[0].run(x, ctx)
        x = self.attentions[0].run(x, ctx)
        return self.resnets[1].run(x, ctx)


class UNet2DConditionModel(ModelBase):
    # the SD1.5 checkpoint the scripts load first (scripts/pose2vid.py:62-65) still carries the output head the
    # ReferenceNet does not have (reference unet_2d_condition.py:645-655)
    _allow_unexpected_keys = ("conv_out.", "conv_norm_out.")

    def __init__(self, sample_size: Optional[int] = None, in_channels: int = 4, out_channels: int = 4,
                 center_input_sample: bool = False, flip_sin_to_cos: bool = True, freq_shift: int = 0,
                 down_block_types: Tuple[str] = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D",
                                                 "CrossAttnDownBlock2D", "DownBlock2D"),
                 mid_block_type: Optional[str] = "UNetMidBlock2DCrossAttn",
                 up_block_types: Tuple[str] = ("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D",
                                               "CrossAttnUpBlock2D"),
                 only_cross_attention=False, block_out_channels: Tuple[int] = (320, 640, 1280, 1280),
                 layers_per_block: int = 2, downsample_padding: int = 1, mid_block_scale_factor: float = 1,
                 act_fn: str = "silu", norm_num_groups: Optional[int] = 32, norm_eps: float = 1e-5,
                 cross_attention_dim: int = 1280, attention_head_dim: Union[int, Tuple[int]] = 8,
                 use_linear_projection: bool = False, **unused):
        super().__init__()
        cfg = {k: v for k, v in locals().items() if k not in ("self", "__class__", "unused")}
        cfg.update(unused)
        self.register_to_config(**cfg)
        if list(down_block_types) != ["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"] or use_linear_projection \
                or center_input_sample or not flip_sin_to_cos or freq_shift != 0:
            raise NotImplementedError("only the SD1.5 UNet2D layout used as AniPortrait's ReferenceNet is supported")
        heads = attention_head_dim if isinstance(attention_head_dim, int) else attention_head_dim[0]
        boc = list(block_out_channels)
        ted = boc[0] * 4
        self.groups, self.eps = norm_num_groups, norm_eps
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], ted)
        self.down_blocks = nn.ModuleList()
        oc = boc[0]
        for i in range(4):
            ic, oc = oc, boc[i]
            cls = CrossAttnDownBlock2D if i < 3 else DownBlock2D
            io = [(ic if j == 0 else oc, oc) for j in range(layers_per_block)]
            self.down_blocks.append(cls(io, oc, heads, cross_attention_dim, i < 3, norm_num_groups, norm_eps,
                                        "down" if i < 3 else None, ted))
        self.mid_block = UNetMidBlock2DCrossAttn(boc[-1], heads, cross_attention_dim, norm_num_groups, norm_eps, ted)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        oc = rev[0]
        for i in range(4):
            prev, oc = oc, rev[i]
            ic = rev[min(i + 1, 3)]
            io = []
            for j in range(layers_per_block + 1):
                res_skip = ic if j == layers_per_block else oc
                res_in = prev if j == 0 else oc
                io.append((res_in + res_skip, oc))
            cls = UpBlock2D if i == 0 else CrossAttnUpBlock2D
            self.up_blocks.append(cls(io, oc, heads, cross_attention_dim, i > 0, norm_num_groups, norm_eps,
                                      "up" if i < 3 else None, ted))
        self._pk = PackedCache()

    def forward(self, sample, timestep, encoder_hidden_states, return_dict: bool = True, **unused):
        """sample [B, 4, h, w]; returns the (unused by AniPortrait) last hidden state, reference layout [B, C, h, w]."""
        if not sample.is_cuda:
            raise RuntimeError("aniportrait_b200.UNet2DConditionModel runs on CUDA (sm_100a) only: no CPU fallback")
        pk = self._pk.get(self, lambda: dict(wi=ops.pack_conv3x3_weight(self.conv_in.weight.detach()),
                                             bi=f32(self.conv_in.bias)))
        B, C, H, W = sample.shape
        x = ops.ncfhw_to_nhwc(sample.to(torch.float16).reshape(B, C, 1, H, W).contiguous(), 64)
        dev = sample.device
        t = timestep if torch.is_tensor(timestep) else torch.tensor([float(timestep)], device=dev)
        t = t.to(device=dev, dtype=torch.float32).reshape(-1)
        if t.numel() == 1:
            t = t.expand(B)
        emb = self.time_embedding.run(ops.timestep_embedding(t.contiguous(), self.conv_in.out_channels))
        ehs = encoder_hidden_states.to(torch.float16).contiguous()
        ctx = RunCtx(B, 1, ops.silu(emb), ehs)
        x = ops.conv3x3(x, pk["wi"], self.conv_in.out_channels, bias=pk["bi"])
        skips = [x]
        for blk in self.down_blocks:
            for j in range(len(blk.resnets)):
                x = blk.layer(j, x, ctx)
                skips.append(x)
            if blk.downsamplers is not None:
                x = blk.downsamplers[0].run(x)
                skips.append(x)
        x = self.mid_block.run(x, ctx)
        for blk in self.up_blocks:
            for j in range(len(blk.resnets)):
                x = blk.layer(j, x, ctx, skip=skips.pop())
            if blk.upsamplers is not None:
                x = blk.upsamplers[0].run(x)
        out = x.permute(0, 3, 1, 2)
        if not return_dict:
            return (out,)
        return UNet2DConditionOutput(sample=out)
