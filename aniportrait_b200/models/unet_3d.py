"""Denoising UNet (mirror of the reference's src/models/unet_3d.py + unet_3d_blocks.py).

Call surface kept: `UNet3DConditionModel(**sd15_config, **unet_additional_kwargs)`, `.from_pretrained_2d(...)`,
`.forward(sample, timestep, encoder_hidden_states, class_labels=None, pose_cond_fea=None, attention_mask=None,
down_block_additional_residuals=None, mid_block_additional_residual=None, return_dict=True)`, `.in_channels`,
`.config`, `.dtype`, `.device`, `.down_blocks / .mid_block / .up_blocks`, reference state-dict keys.
The forward pass is a fixed sequence of sm_100a kernels over channels-last fp16 activations.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Tuple, Union

import torch
from torch import nn

from .. import ops
from .blocks import (Downsample, ResnetBlock, RunCtx, Transformer3DModel, Upsample, VanillaTemporalModule)
from .modeling import ModelBase, PackedCache, f16, f32, load_checked


@dataclass
class UNet3DConditionOutput:
    sample: torch.Tensor


def get_motion_module(in_channels, motion_module_type, motion_module_kwargs):
    if motion_module_type == "Vanilla":
        return VanillaTemporalModule(in_channels=in_channels, **(motion_module_kwargs or {}))
    raise ValueError(motion_module_type)


class _Block3D(nn.Module):
    """One resolution level: [ResnetBlock -> (Transformer3DModel) -> (motion module)] x layers (+ down/upsampler).
    Mirrors CrossAttnDownBlock3D / DownBlock3D / CrossAttnUpBlock3D / UpBlock3D (reference unet_3d_blocks.py)."""

    def __init__(self, resnet_io, out_channels, heads, cross_attention_dim, has_attn, use_motion, motion_type,
                 motion_kwargs, groups, eps, sampler, temb_channels=1280):
        super().__init__()
        self.has_cross_attention = has_attn
        self.resnets = nn.ModuleList([ResnetBlock(i, o, temb_channels, groups, eps) for i, o in resnet_io])
        if has_attn:
            self.attentions = nn.ModuleList([Transformer3DModel(heads, out_channels // heads, out_channels,
                                                                cross_attention_dim, groups) for _ in resnet_io])
        self.motion_modules = nn.ModuleList([get_motion_module(out_channels, motion_type, motion_kwargs)
                                             if use_motion else None for _ in resnet_io])
        self.downsamplers = nn.ModuleList([Downsample(out_channels)]) if sampler == "down" else None
        self.upsamplers = nn.ModuleList([Upsample(out_channels)]) if sampler == "up" else None

    def layer(self, j, x, ctx, skip=None):
        x = self.resnets[j].run(x, ctx, skip)
        if self.has_cross_attention:
            x = self.attentions[j].run(x, ctx)
        if self.motion_modules[j] is not None:
            x = self.motion_modules[j].run(x, ctx)
        return x


class CrossAttnDownBlock3D(_Block3D):
    pass


class DownBlock3D(_Block3D):
    pass


class CrossAttnUpBlock3D(_Block3D):
    pass


class UpBlock3D(_Block3D):
    pass


class UNetMidBlock3DCrossAttn(nn.Module):
    """resnet -> Transformer3DModel -> motion module -> resnet (reference unet_3d_blocks.py:171-293)."""

    def __init__(self, channels, heads, cross_attention_dim, use_motion, motion_type, motion_kwargs, groups, eps,
                 temb_channels=1280):
        super().__init__()
        self.has_cross_attention = True
        self.resnets = nn.ModuleList([ResnetBlock(channels, channels, temb_channels, groups, eps) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer3DModel(heads, channels // heads, channels, cross_attention_dim,
                                                            groups)])
        self.motion_modules = nn.ModuleList([get_motion_module(channels, motion_type, motion_kwargs)
                                             if use_motion else None])

    def run(self, x, ctx):
        x = self.resnets[0].run(x, ctx)
        x = self.attentions[0].run(x, ctx)
        if self.motion_modules[0] is not None:
            x = self.motion_modules[0].run(x, ctx)
        return self.resnets[1].run(x, ctx)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)
        self._pk = PackedCache()

    def run(self, t_emb):
        pk = self._pk.get(self, lambda: dict(w1=f16(self.linear_1.weight), b1=f32(self.linear_1.bias),
                                             w2=f16(self.linear_2.weight), b2=f32(self.linear_2.bias)))
        h = ops.silu(ops.gemm(t_emb, pk["w1"], bias=pk["b1"]))
        return ops.gemm(h, pk["w2"], bias=pk["b2"])


class UNet3DConditionModel(ModelBase):
    _supports_gradient_checkpointing = False

    def __init__(self, sample_size: Optional[int] = None, in_channels: int = 4, out_channels: int = 4,
                 center_input_sample: bool = False, flip_sin_to_cos: bool = True, freq_shift: int = 0,
                 down_block_types: Tuple[str] = ("CrossAttnDownBlock3D", "CrossAttnDownBlock3D",
                                                 "CrossAttnDownBlock3D", "DownBlock3D"),
                 mid_block_type: str = "UNetMidBlock3DCrossAttn",
                 up_block_types: Tuple[str] = ("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D",
                                               "CrossAttnUpBlock3D"),
                 only_cross_attention=False, block_out_channels: Tuple[int] = (320, 640, 1280, 1280),
                 layers_per_block: int = 2, downsample_padding: int = 1, mid_block_scale_factor: float = 1,
                 act_fn: str = "silu", norm_num_groups: int = 32, norm_eps: float = 1e-5,
                 cross_attention_dim: int = 1280, attention_head_dim: Union[int, Tuple[int]] = 8,
                 dual_cross_attention: bool = False, use_linear_projection: bool = False,
                 class_embed_type: Optional[str] = None, num_class_embeds: Optional[int] = None,
                 upcast_attention: bool = False, resnet_time_scale_shift: str = "default",
                 use_inflated_groupnorm=False, use_motion_module=False, motion_module_resolutions=(1, 2, 4, 8),
                 motion_module_mid_block=False, motion_module_decoder_only=False, motion_module_type=None,
                 motion_module_kwargs=None, unet_use_cross_frame_attention=None, unet_use_temporal_attention=None):
        super().__init__()
        cfg = {k: v for k, v in locals().items() if k not in ("self", "__class__")}
        self.register_to_config(**cfg)
        # the fused path implements exactly the configuration the reference's inference uses
        unsupported = dict(center_input_sample=center_input_sample, dual_cross_attention=dual_cross_attention,
                           use_linear_projection=use_linear_projection, upcast_attention=upcast_attention,
                           unet_use_cross_frame_attention=bool(unet_use_cross_frame_attention),
                           unet_use_temporal_attention=bool(unet_use_temporal_attention),
                           class_embed=class_embed_type is not None or num_class_embeds is not None,
                           only_cross_attention=bool(only_cross_attention) if isinstance(only_cross_attention, bool)
                           else any(only_cross_attention))
        bad = [k for k, v in unsupported.items() if v]
        if bad or resnet_time_scale_shift != "default" or act_fn not in ("silu", "swish") or downsample_padding != 1 \
                or not flip_sin_to_cos or freq_shift != 0:
            raise NotImplementedError(f"UNet3DConditionModel options outside the AniPortrait inference config: {bad}")
        if list(down_block_types) != ["CrossAttnDownBlock3D"] * 3 + ["DownBlock3D"] or \
                list(up_block_types) != ["UpBlock3D"] + ["CrossAttnUpBlock3D"] * 3:
            raise NotImplementedError("only the SD1.5 block layout is supported")
        if motion_module_decoder_only:
            raise NotImplementedError("motion_module_decoder_only")
        heads = attention_head_dim if isinstance(attention_head_dim, int) else attention_head_dim[0]
        self.sample_size = sample_size
        self.heads = heads
        self.groups, self.eps = norm_num_groups, norm_eps
        boc = list(block_out_channels)
        time_embed_dim = boc[0] * 4
        mk = dict(motion_type=motion_module_type, motion_kwargs=motion_module_kwargs, groups=norm_num_groups,
                  eps=norm_eps, temb_channels=time_embed_dim)

        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], time_embed_dim)
        self.down_blocks = nn.ModuleList()
        oc = boc[0]
        for i in range(4):
            ic, oc = oc, boc[i]
            cls = CrossAttnDownBlock3D if i < 3 else DownBlock3D
            use_mm = use_motion_module and (2 ** i in motion_module_resolutions)
            io = [(ic if j == 0 else oc, oc) for j in range(layers_per_block)]
            self.down_blocks.append(cls(io, oc, heads, cross_attention_dim, i < 3, use_mm,
                                        sampler="down" if i < 3 else None, **mk))
        self.mid_block = UNetMidBlock3DCrossAttn(boc[-1], heads, cross_attention_dim,
                                                 use_motion_module and motion_module_mid_block, motion_module_type,
                                                 motion_module_kwargs, norm_num_groups, norm_eps, time_embed_dim)
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
            cls = UpBlock3D if i == 0 else CrossAttnUpBlock3D
            use_mm = use_motion_module and (2 ** (3 - i) in motion_module_resolutions)
            self.up_blocks.append(cls(io, oc, heads, cross_attention_dim, i > 0, use_mm,
                                      sampler="up" if i < 3 else None, **mk))
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, boc[0], eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)
        self._pk = PackedCache()
        self._time_dim = time_embed_dim

    # ------------------------------------------------------------------------------------------------ packing
    def packed(self):
        def build():
            cout = self.conv_out.out_channels
            wo = ops.pack_conv3x3_weight(self.conv_out.weight.detach())
            bo = torch.zeros(wo.shape[0], dtype=torch.float32, device=wo.device)
            bo[:cout] = f32(self.conv_out.bias)
            # every resnet's time-embedding projection in ONE GEMM per call (reference resnet.py:226-230 runs 22 small
            # Linears): rows of all time_emb_proj weights concatenated; resnet r reads columns [off_r, off_r + cout_r)
            from .blocks import ResnetBlock
            res = [m for m in self.modules() if isinstance(m, ResnetBlock) and m.time_emb_proj is not None]
            temb = None
            if res:
                pks = [m.packed() for m in res]
                offs, o = [], 0
                for m in res:
                    offs.append(o)
                    o += m.out_channels
                if o % 32 == 0:
                    temb = dict(w=torch.cat([q["wt"] for q in pks], 0).contiguous(),
                                b=torch.cat([q["bt"] for q in pks], 0).contiguous(),
                                slices=[(id(m), off, m.out_channels) for m, off in zip(res, offs)])
            return dict(wi=ops.pack_conv3x3_weight(self.conv_in.weight.detach()), bi=f32(self.conv_in.bias),
                        gn=f32(self.conv_norm_out.weight), bn=f32(self.conv_norm_out.bias), wo=wo, bo=bo, temb=temb)
        return self._pk.get(self, build)

    # ------------------------------------------------------------------------------------------------ forward
    def forward_nhwc(self, x: torch.Tensor, batch: int, frames: int, timestep, encoder_hidden_states,
                     pose_nhwc=None, ref_branch=None, ehs_key=None, group=None) -> torch.Tensor:
        """x: [(batch frames), H, W, 64] fp16 (4 latent channels zero padded to 64). pose_nhwc: 5 maps, each
        [(batch frames) | frames, h, w, C] (a [frames,...] map is shared by all CFG branches).
        Returns [(batch frames), H, W, out_channels] fp16."""
        if not x.is_cuda:
            raise RuntimeError("aniportrait_b200.UNet3DConditionModel runs on CUDA (sm_100a) only: no CPU fallback")
        pk = self.packed()
        dev = x.device
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([float(t)], dtype=torch.float32, device=dev)
        t = t.to(device=dev, dtype=torch.float32).reshape(-1)
        if t.numel() == 1:
            t = t.expand(batch)
        t_emb = ops.timestep_embedding(t.contiguous(), self.conv_in.out_channels)
        emb = self.time_embedding.run(t_emb)
        ehs = encoder_hidden_states.to(torch.float16).contiguous() if encoder_hidden_states is not None else None
        temb_act = ops.silu(emb)
        temb_bias = None
        if pk["temb"] is not None:
            tb = ops.gemm(temb_act, pk["temb"]["w"], bias=pk["temb"]["b"], out_f32=True)            # [B, sum of couts] fp32
            temb_bias = {rid: tb[:, off:off + n] for rid, off, n in pk["temb"]["slices"]}
        ctx = RunCtx(batch, frames, temb_act, ehs, ehs_key=ehs_key, ref_branch=ref_branch, temb_bias=temb_bias,
                     group=group)

        def add_pose(x, k):
            if pose_nhwc is None:
                return x
            p = pose_nhwc[k]
            return ops.add(x, p) if p.shape[0] == x.shape[0] else ops.add_bcast(x, p)

        x = ops.conv3x3(x, pk["wi"], self.conv_in.out_channels, bias=pk["bi"])
        x = add_pose(x, 0)
        skips = [x]
        for i, blk in enumerate(self.down_blocks):
            for j in range(len(blk.resnets)):
                x = blk.layer(j, x, ctx)
                skips.append(x)
            if blk.downsamplers is not None:
                x = blk.downsamplers[0].run(x)
                skips.append(x)
            x = add_pose(x, i + 1)
        x = self.mid_block.run(x, ctx)
        for blk in self.up_blocks:
            for j in range(len(blk.resnets)):
                x = blk.layer(j, x, ctx, skip=skips.pop())
            if blk.upsamplers is not None:
                x = blk.upsamplers[0].run(x)
        from .blocks import _cs
        hn = ops.group_norm(x, pk["gn"], pk["bn"], self.groups, self.eps, True, stats=_cs(x))
        return ops.conv3x3(hn, pk["wo"], self.conv_out.out_channels, bias=pk["bo"])

    def prepare_reference(self, batch: int, frames: int, encoder_hidden_states, ehs_key=None, ref_branch=None, group=None):
        """Once per video, before the denoising loop: every reader block projects its ReferenceNet bank to K/V and
        evaluates its (query-independent) attn2 constant, so that the per-step forward — typically replayed from a CUDA
        graph — contains none of this step-invariant work."""
        from .blocks import BasicTransformerBlock
        ehs = encoder_hidden_states.to(torch.float16).contiguous()
        if ehs_key is None:
            raise ValueError("prepare_reference needs an explicit ehs_key (the identity of this video's conditioning)")
        ctx = RunCtx(batch, frames, None, ehs, ehs_key=ehs_key, ref_branch=ref_branch, group=group)
        from .blocks import TemporalTransformer3DModel
        for m in self.modules():
            if isinstance(m, BasicTransformerBlock):
                m.prepare(ctx)
            elif isinstance(m, TemporalTransformer3DModel):
                m.prepare(batch, frames)       # positional-encoding bias tables of the folded q/k/v projections

    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, pose_cond_fea=None,
                attention_mask=None, down_block_additional_residuals=None, mid_block_additional_residual=None,
                return_dict: bool = True):
        """Reference-layout entry point: sample [B, 4, F, h, w] -> [B, 4, F, h, w] (reference unet_3d.py:399-580)."""
        if class_labels is not None or attention_mask is not None or down_block_additional_residuals is not None \
                or mid_block_additional_residual is not None:
            raise NotImplementedError("class_labels / attention_mask / additional residuals are unused by AniPortrait")
        if sample.dim() != 5:
            raise ValueError(f"Expected sample to have ndim=5, but got ndim={sample.dim()}.")
        B, C, F, H, W = sample.shape
        if H % 8 or W % 8:
            raise ValueError("latent height/width must be multiples of 8")
        x = ops.ncfhw_to_nhwc(sample.to(torch.float16).contiguous(), 64)
        pose = None
        if pose_cond_fea is not None:
            pose = []
            for p in pose_cond_fea:
                if p.dim() == 5:   # reference layout [B, C, F, h, w]
                    p = p.to(torch.float16).permute(0, 2, 3, 4, 1).reshape(B * F, p.shape[3], p.shape[4], p.shape[1])
                pose.append(p.contiguous())
        out = self.forward_nhwc(x, B, F, timestep, encoder_hidden_states, pose)
        out = ops.nhwc_to_ncfhw(out, B, self.conv_out.out_channels, F)
        if not return_dict:
            return (out,)
        return UNet3DConditionOutput(sample=out)

    # ------------------------------------------------------------------------------------------------ loading
    @classmethod
    def from_pretrained_2d(cls, pretrained_model_path, motion_module_path, subfolder=None,
                           unet_additional_kwargs=None, mm_zero_proj_out=False):
        """Same contract as the reference (unet_3d.py:582-673): SD1.5 2-D weights + motion-module weights."""
        pretrained_model_path = Path(pretrained_model_path)
        motion_module_path = Path(motion_module_path)
        if subfolder is not None:
            pretrained_model_path = pretrained_model_path.joinpath(subfolder)
        config_file = pretrained_model_path / "config.json"
        if not (config_file.exists() and config_file.is_file()):
            raise RuntimeError(f"{config_file} does not exist or is not a file")
        unet_config = cls.load_config(config_file)
        unet_config["_class_name"] = cls.__name__
        unet_config["down_block_types"] = ["CrossAttnDownBlock3D"] * 3 + ["DownBlock3D"]
        unet_config["up_block_types"] = ["UpBlock3D"] + ["CrossAttnUpBlock3D"] * 3
        unet_config["mid_block_type"] = "UNetMidBlock3DCrossAttn"
        model = cls.from_config(unet_config, **(unet_additional_kwargs or {}))
        st = pretrained_model_path / "diffusion_pytorch_model.safetensors"
        pt = pretrained_model_path / "diffusion_pytorch_model.bin"
        if st.exists():
            from safetensors.torch import load_file
            state_dict = load_file(st, device="cpu")
        elif pt.exists():
            state_dict = torch.load(pt, map_location="cpu", weights_only=True)
        else:
            raise FileNotFoundError(f"no weights file found in {pretrained_model_path}")
        if motion_module_path.exists() and motion_module_path.is_file():
            if motion_module_path.suffix.lower() in [".pth", ".pt", ".ckpt"]:
                motion_state_dict = torch.load(motion_module_path, map_location="cpu", weights_only=True)
            elif motion_module_path.suffix.lower() == ".safetensors":
                from safetensors.torch import load_file
                motion_state_dict = load_file(motion_module_path, device="cpu")
            else:
                raise RuntimeError(f"unknown file format for motion module weights: {motion_module_path.suffix}")
            if mm_zero_proj_out:
                motion_state_dict = OrderedDict((k, v) for k, v in motion_state_dict.items() if "proj_out" not in k)
            state_dict.update(motion_state_dict)
        # the reference loads with strict=False and prints the counts (unet_3d.py:663-673); here everything except the
        # motion-module keys (absent when no motion file is given, or proj_out dropped by mm_zero_proj_out) and the
        # position-encoding buffers must be present, and nothing unknown may be in the files
        load_checked(model, state_dict, allow_missing=("motion_modules.", ".pos_encoder.pe"),
                     allow_unexpected=(".pos_encoder.pe",))
        return model
