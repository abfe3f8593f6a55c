"""AutoencoderKL (sd-vae-ft-mse layout) with diffusers' parameter names so published checkpoints load.

Role on the hot path: `decode` of the denoised latents (reference pipeline_pose2vid_long.py:113-126, one frame per
call) and a single `encode` of the reference image (once per video, :430-431).
  * `decode` on a CUDA fp16 model runs on the sm_100a kernels (channels-last, batched over frames): implicit-GEMM 3x3
    convs, fused GroupNorm+SiLU, tcgen05 GEMMs for the 1x1 shortcuts and the mid-block attention, which (single head,
    d = 512: too wide for the fused attention kernel's TMEM budget) is evaluated per frame as GEMM -> row-softmax -> GEMM.
  * `encode` on a CUDA fp16 model takes the same kernels (the stride-2 downsamplers as stride-1 convolutions whose odd
    outputs are gathered); any non-fp16 / non-CUDA use runs as torch library ops.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .modeling import ModelBase, PackedCache, f16, f32


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


# ------------------------------------------------------------------------------------------------ kernel-path helpers
def _pack_conv(c: nn.Conv2d):
    w = ops.pack_conv3x3_weight(c.weight.detach())
    b = torch.zeros(w.shape[0], dtype=torch.float32, device=w.device)
    b[:c.out_channels] = f32(c.bias)
    return w, b


def _pack_res(r: "_Resnet"):
    d = dict(g1=f32(r.norm1.weight), b1=f32(r.norm1.bias), g2=f32(r.norm2.weight), b2=f32(r.norm2.bias),
             c1=_pack_conv(r.conv1), c2=_pack_conv(r.conv2), cout=r.conv1.out_channels)
    if r.conv_shortcut is not None:
        d["ws"] = f16(r.conv_shortcut.weight.reshape(r.conv_shortcut.out_channels, -1))
        d["bs"] = f32(r.conv_shortcut.bias)
    return d


def _pack_attn(a: "_Attn"):
    c = a.to_q.weight.shape[0]
    scale = c ** -0.5
    # softmax scale folded into the q projection; v bias folded into the output bias (softmax rows sum to 1):
    # out = P.(xWv^T) Wo^T + (Wo bv + bo)
    return dict(g=f32(a.group_norm.weight), b=f32(a.group_norm.bias), wq=f16(a.to_q.weight * scale),
                bq=f32(a.to_q.bias * scale), wk=f16(a.to_k.weight), bk=f32(a.to_k.bias), wv=f16(a.to_v.weight),
                wo=f16(a.to_out[0].weight), bo=f32(a.to_out[0].bias) + f32(a.to_out[0].weight) @ f32(a.to_v.bias), c=c)


def _res_run(x, d, groups):
    nf, h, w, cin = x.shape
    hn = ops.group_norm(x, d["g1"], d["b1"], groups, 1e-6, True)
    hc = ops.conv3x3(hn, d["c1"][0], d["cout"], bias=d["c1"][1])
    hn2 = ops.group_norm(hc, d["g2"], d["b2"], groups, 1e-6, True)
    res = ops.gemm(x.view(-1, cin), d["ws"], bias=d["bs"]).view(nf, h, w, d["cout"]) if "ws" in d else x
    return ops.conv3x3(hn2, d["c2"][0], d["cout"], bias=d["c2"][1], residual=res)


def _mid_attn_run(x, a, groups):
    """Single-head mid-block attention (d = C = 512), one frame at a time: GEMM -> row softmax -> GEMM (the [tokens, tokens]
    score matrix is 32 MB per frame at 512x512)."""
    nf, h, w, c = x.shape
    n = h * w
    hn = ops.group_norm(x, a["g"], a["b"], groups, 1e-6, False).view(nf, n, c)
    q = ops.gemm(hn.view(-1, c), a["wq"], bias=a["bq"]).view(nf, n, c)
    k = ops.gemm(hn.view(-1, c), a["wk"], bias=a["bk"]).view(nf, n, c)
    att = torch.empty(nf, n, c, dtype=torch.float16, device=x.device)
    npad = (n + 31) // 32 * 32
    for f in range(nf):
        vt = ops.gemm(a["wv"], hn[f].contiguous())                   # V^T = Wv . X^T  [c, n]
        sc = ops.gemm(q[f], k[f].contiguous() if npad == n else
                      torch.cat([k[f], k[f].new_zeros(npad - n, c)]), n_valid=n)   # [n, n] scaled scores
        ops.softmax_rows(sc)
        ops.gemm(sc, vt, out=att[f])                                  # P . V   (K = n keys)
    return ops.gemm(att.view(-1, c), a["wo"], bias=a["bo"], residual=x.view(-1, c)).view(nf, h, w, c)


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

    # ------------------------------------------------------------------------------------------ kernel path
    def _packed(self):
        if not hasattr(self, "_pk"):
            self._pk = PackedCache()

        def build():
            return dict(conv_in=_pack_conv(self.conv_in),
                        down=[dict(res=[_pack_res(r) for r in b.resnets],
                                   down=_pack_conv(b.downsamplers[0].conv) if b._sampler == "down" else None)
                              for b in self.down_blocks],
                        mid=[_pack_res(r) for r in self.mid_block.resnets], attn=_pack_attn(self.mid_block.attentions[0]),
                        gn=(f32(self.conv_norm_out.weight), f32(self.conv_norm_out.bias)), conv_out=_pack_conv(self.conv_out))
        return self._pk.get(self, build)

    def run_nhwc(self, x: torch.Tensor, groups: int = 32) -> torch.Tensor:
        """x: [Nf, H, W, 64] fp16 (3 image channels zero padded) -> moments [Nf, H/8, W/8, 2*latent] fp16.
        diffusers' Downsample2D pads right/bottom by one and convolves with stride 2 and no padding:
        out[o] = sum_k w[k] x[2o + k]. That is every other output of the ordinary stride-1, pad-1 convolution
        (y[p] = sum_k w[k] x[p + k - 1], p = 2o + 1), which the implicit-GEMM kernel computes; the odd rows / columns are
        then gathered (4x the FLOPs of three small layers instead of a dedicated asymmetric-padding mode)."""
        pk = self._packed()
        x = ops.conv3x3(x, pk["conv_in"][0], self.conv_in.out_channels, bias=pk["conv_in"][1])
        for blk in pk["down"]:
            for d in blk["res"]:
                x = _res_run(x, d, groups)
            if blk["down"] is not None:
                y = ops.conv3x3(x, blk["down"][0], x.shape[-1], bias=blk["down"][1])
                x = y[:, 1::2, 1::2, :].contiguous()
        x = _res_run(x, pk["mid"][0], groups)
        x = _mid_attn_run(x, pk["attn"], groups)
        x = _res_run(x, pk["mid"][1], groups)
        hn = ops.group_norm(x, pk["gn"][0], pk["gn"][1], groups, 1e-6, True)
        return ops.conv3x3(hn, pk["conv_out"][0], self.conv_out.out_channels, bias=pk["conv_out"][1])


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

    # ------------------------------------------------------------------------------------------ kernel path
    def _packed(self):
        if not hasattr(self, "_pk"):
            self._pk = PackedCache()

        def build():
            return dict(conv_in=_pack_conv(self.conv_in), mid=[_pack_res(r) for r in self.mid_block.resnets],
                        attn=_pack_attn(self.mid_block.attentions[0]),
                        up=[dict(res=[_pack_res(r) for r in b.resnets],
                                 up=_pack_conv(b.upsamplers[0].conv) if b._sampler == "up" else None)
                            for b in self.up_blocks],
                        gn=(f32(self.conv_norm_out.weight), f32(self.conv_norm_out.bias)),
                        conv_out=_pack_conv(self.conv_out))
        return self._pk.get(self, build)

    def run_nhwc(self, z: torch.Tensor, groups: int = 32) -> torch.Tensor:
        """z: [Nf, h, w, 64] fp16 (4 latent channels zero padded) -> [Nf, 8h, 8w, 3] fp16."""
        pk = self._packed()
        x = ops.conv3x3(z, pk["conv_in"][0], self.conv_in.out_channels, bias=pk["conv_in"][1])
        x = _res_run(x, pk["mid"][0], groups)
        x = _mid_attn_run(x, pk["attn"], groups)
        x = _res_run(x, pk["mid"][1], groups)
        for blk in pk["up"]:
            for d in blk["res"]:
                x = _res_run(x, d, groups)
            if blk["up"] is not None:
                x = ops.conv3x3(ops.upsample2x(x), blk["up"][0], x.shape[-1], bias=blk["up"][1])
        hn = ops.group_norm(x, pk["gn"][0], pk["gn"][1], groups, 1e-6, True)
        return ops.conv3x3(hn, pk["conv_out"][0], self.conv_out.out_channels, bias=pk["conv_out"][1])


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
        """x [n, 3, H, W] -> latent distribution. fp16 CUDA models with 64-multiple widths take the sm_100a kernel path
        (set `kernel_encode = False` to force the torch modules)."""
        if getattr(self, "kernel_encode", True) and self._kernel_decode_ok(x) and x.shape[1] <= 64 and x.shape[-1] % 64 == 0 \
                and x.shape[-2] % 64 == 0:
            n, c, h, w = x.shape
            xp = torch.zeros(n, h, w, 64, dtype=torch.float16, device=x.device)
            xp[..., :c] = x.permute(0, 2, 3, 1).to(torch.float16)
            m = self.encoder.run_nhwc(xp, self.config.norm_num_groups).permute(0, 3, 1, 2)
            return AutoencoderKLOutput(latent_dist=_LatentDist(self.quant_conv(m.to(self.quant_conv.weight.dtype))))
        return AutoencoderKLOutput(latent_dist=_LatentDist(self.quant_conv(self.encoder(x))))

    def _kernel_decode_ok(self, z):
        boc = self.config.block_out_channels
        return (z.is_cuda and self.dtype == torch.float16 and all(c % 64 == 0 for c in boc)
                and z.shape[-1] % 8 == 0 and z.shape[-2] % 8 == 0 and (z.shape[-1] * z.shape[-2]) % 64 == 0)

    @torch.no_grad()
    def decode(self, z, return_dict=True, generator=None):
        """z [n, 4, h, w] -> sample [n, 3, 8h, 8w]. fp16 CUDA models with 64-multiple widths (sd-vae-ft-mse) take the
        sm_100a kernel path; the 4->4 post_quant 1x1 conv is a per-pixel 4x4 matmul done while converting layouts."""
        if not self._kernel_decode_ok(z):
            return DecoderOutput(sample=self.decoder(self.post_quant_conv(z)))
        n, c, h, w = z.shape
        wq = self.post_quant_conv.weight.reshape(c, c).to(torch.float32)
        zq = torch.einsum("nchw,oc->nhwo", z.to(torch.float32), wq) + self.post_quant_conv.bias.to(torch.float32)
        zp = torch.zeros(n, h, w, 64, dtype=torch.float16, device=z.device)
        zp[..., :c] = zq.to(torch.float16)
        out = self.decoder.run_nhwc(zp, self.config.norm_num_groups)
        return DecoderOutput(sample=out.permute(0, 3, 1, 2))
