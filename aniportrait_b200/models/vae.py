"""AutoencoderKL (sd-vae-ft-mse layout) with diffusers' parameter names so published checkpoints load.

Role on the hot path: `decode` of the denoised latents (reference pipeline_pose2vid_long.py:113-126, one frame per
call) and a single `encode` of the reference image (once per video, :430-431).
  * `decode` runs on the sm_100a kernels (channels-last, batched over frames): implicit-GEMM 3x3 convs, fused
    GroupNorm+SiLU, tcgen05 GEMMs for the 1x1 shortcuts, the 4x4 / 8x8 (post_)quant 1x1 convs and the mid-block attention,
    which (single head, d = 512: too wide for the fused attention kernel's TMEM budget) is evaluated per frame as
    GEMM -> row-softmax -> GEMM.
  * `encode` takes the same kernels (the stride-2 downsamplers as stride-1 convolutions whose odd outputs are gathered).
There is no torch-op / CPU / fp32 path: the modules below are parameter holders (state-dict surface) with a kernel `run`;
calling encode / decode on anything but a CUDA fp16 model with 64-multiple block widths raises. (The fp32 torch evaluation
used by the tests lives in oracle/functional.py: vae_decode / vae_encode.)
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
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


class _Attn(nn.Module):
    def __init__(self, c, groups=32, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])


class _Mid(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.attentions = nn.ModuleList([_Attn(c, groups)])
        self.resnets = nn.ModuleList([_Resnet(c, c, groups), _Resnet(c, c, groups)])


class _Sampler(nn.Module):
    def __init__(self, c, down):
        super().__init__()
        self.down = down
        self.conv = nn.Conv2d(c, c, 3, stride=2 if down else 1, padding=0 if down else 1)


class _Stage(nn.Module):
    def __init__(self, cin, cout, layers, groups, sampler):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        if sampler == "down":
            self.downsamplers = nn.ModuleList([_Sampler(cout, True)])
        elif sampler == "up":
            self.upsamplers = nn.ModuleList([_Sampler(cout, False)])
        self._sampler = sampler


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

    # ------------------------------------------------------------------------------------------ kernel path
    def _packed(self, quant_conv: nn.Conv2d):
        if not hasattr(self, "_pk"):
            self._pk = PackedCache()
            self._pk_quant = None

        def build():
            # quant_conv (1x1, 8 -> 8) directly follows conv_out (3x3): their composition is ONE 3x3 convolution with
            # W'[o, c, ky, kx] = sum_k Wq[o, k] W_out[k, c, ky, kx],  b' = Wq b_out + bq   (exact; folded in fp32)
            wq = quant_conv.weight.detach().to(torch.float32).reshape(quant_conv.out_channels, -1)
            w_out = torch.einsum("ok,kcyx->ocyx", wq, self.conv_out.weight.detach().to(torch.float32))
            b_out = wq @ self.conv_out.bias.detach().to(torch.float32) + quant_conv.bias.detach().to(torch.float32)
            wo = ops.pack_conv3x3_weight(w_out)
            bo = torch.zeros(wo.shape[0], dtype=torch.float32, device=wo.device)
            bo[:b_out.numel()] = b_out
            return dict(conv_in=_pack_conv(self.conv_in),
                        down=[dict(res=[_pack_res(r) for r in b.resnets],
                                   down=_pack_conv(b.downsamplers[0].conv) if b._sampler == "down" else None)
                              for b in self.down_blocks],
                        mid=[_pack_res(r) for r in self.mid_block.resnets], attn=_pack_attn(self.mid_block.attentions[0]),
                        gn=(f32(self.conv_norm_out.weight), f32(self.conv_norm_out.bias)), conv_out=(wo, bo))
        qkey = tuple((p.data_ptr(), p._version) for p in quant_conv.parameters())
        if qkey != self._pk_quant:      # the cache below is keyed on this module's parameters only
            self._pk._key = None
            self._pk_quant = qkey
        return self._pk.get(self, build)

    def run_nhwc(self, x: torch.Tensor, quant_conv: nn.Conv2d, groups: int = 32) -> torch.Tensor:
        """x: [Nf, H, W, 64] fp16 (3 image channels zero padded) -> moments AFTER quant_conv [Nf, H/8, W/8, 2*latent] fp16.
        diffusers' Downsample2D pads right/bottom by one and convolves with stride 2 and no padding:
        out[o] = sum_k w[k] x[2o + k]. That is every other output of the ordinary stride-1, pad-1 convolution
        (y[p] = sum_k w[k] x[p + k - 1], p = 2o + 1), which the implicit-GEMM kernel computes; the odd rows / columns are
        then gathered (4x the FLOPs of three small layers instead of a dedicated asymmetric-padding mode)."""
        pk = self._packed(quant_conv)
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

    # ------------------------------------------------------------------------------------------ kernel path
    def _packed(self, post_quant_conv: nn.Conv2d):
        if not hasattr(self, "_pk"):
            self._pk = PackedCache()
            self._pk_quant = None

        def build():
            # post_quant_conv (1x1, 4 -> 4) precedes conv_in (3x3, zero padding): folded into conv_in over the input
            # [z (4 channels), 1 (a constant-one channel)]: W'[o, c<4] = sum_k W_in[o, k] Wq[k, c], W'[o, 4] = sum_k W_in[o, k]
            # bq[k]. The ones channel is zero outside the image like every padded tap, so the bias term vanishes at the
            # border exactly as in conv_in(post_quant_conv(z)).
            lat = post_quant_conv.in_channels
            wq = post_quant_conv.weight.detach().to(torch.float32).reshape(post_quant_conv.out_channels, lat)
            w_in = self.conv_in.weight.detach().to(torch.float32)                       # [C, 4, 3, 3]
            w_fold = torch.zeros(w_in.shape[0], lat + 1, 3, 3, dtype=torch.float32, device=w_in.device)
            w_fold[:, :lat] = torch.einsum("okyx,kc->ocyx", w_in, wq)
            w_fold[:, lat] = torch.einsum("okyx,k->oyx", w_in, post_quant_conv.bias.detach().to(torch.float32))
            wi = ops.pack_conv3x3_weight(w_fold)
            bi = torch.zeros(wi.shape[0], dtype=torch.float32, device=wi.device)
            bi[:self.conv_in.out_channels] = f32(self.conv_in.bias)
            return dict(conv_in=(wi, bi), mid=[_pack_res(r) for r in self.mid_block.resnets],
                        attn=_pack_attn(self.mid_block.attentions[0]),
                        up=[dict(res=[_pack_res(r) for r in b.resnets],
                                 up=_pack_conv(b.upsamplers[0].conv) if b._sampler == "up" else None)
                            for b in self.up_blocks],
                        gn=(f32(self.conv_norm_out.weight), f32(self.conv_norm_out.bias)),
                        conv_out=_pack_conv(self.conv_out))
        qkey = tuple((p.data_ptr(), p._version) for p in post_quant_conv.parameters())
        if qkey != self._pk_quant:
            self._pk._key = None
            self._pk_quant = qkey
        return self._pk.get(self, build)

    def run_nhwc(self, z: torch.Tensor, post_quant_conv: nn.Conv2d, groups: int = 32) -> torch.Tensor:
        """z: [Nf, h, w, 64] fp16 (channels: 4 latent, then a constant 1, then zeros) -> [Nf, 8h, 8w, 3] fp16."""
        pk = self._packed(post_quant_conv)
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

    # ------------------------------------------------------------------------------------------------ kernel path only
    def _require_kernel_path(self, t, what):
        boc = self.config.block_out_channels
        if not (t.is_cuda and self.dtype == torch.float16 and all(c % 64 == 0 for c in boc)
                and t.shape[-1] % 8 == 0 and t.shape[-2] % 8 == 0):
            raise RuntimeError(
                f"aniportrait_b200.AutoencoderKL.{what} runs on the sm_100a kernels only: needs a CUDA fp16 model with block "
                f"widths that are multiples of 64 and H, W multiples of 8 (got device={t.device}, dtype={self.dtype}, "
                f"block_out_channels={tuple(boc)}, input {tuple(t.shape)}); there is no torch-op fallback")

    @torch.no_grad()
    def encode(self, x, return_dict=True):
        """x [n, 3, H, W] -> latent distribution (reference pipeline_pose2vid_long.py:430-431 takes .mean)."""
        self._require_kernel_path(x, "encode")
        if x.shape[-1] % 64 or x.shape[-2] % 64:
            raise RuntimeError("AutoencoderKL.encode: image height / width must be multiples of 64")
        n, c, h, w = x.shape
        xp = ops.ncfhw_to_nhwc(x.to(torch.float16).contiguous().view(n, c, 1, h, w), 64)        # [n, H, W, 64]
        m = self.encoder.run_nhwc(xp, self.quant_conv, self.config.norm_num_groups)              # [n, h, w, 2*latent]
        nl = m.shape[-1]
        moments = ops.nhwc_to_ncfhw(m, n, nl, 1).view(n, nl, m.shape[1], m.shape[2])
        return AutoencoderKLOutput(latent_dist=_LatentDist(moments))

    @torch.no_grad()
    def decode(self, z, return_dict=True, generator=None):
        """z [n, 4, h, w] -> sample [n, 3, 8h, 8w] (reference pipeline_pose2vid_long.py:118-121), batched over frames."""
        self._require_kernel_path(z, "decode")
        n, c, h, w = z.shape
        zp = ops.ncfhw_to_nhwc(z.to(torch.float16).contiguous().view(n, c, 1, h, w), 64)         # [n, h, w, 64]
        zp[..., c] = 1.0                                                # the constant-one channel (see Decoder._packed)
        out = self.decoder.run_nhwc(zp, self.post_quant_conv, self.config.norm_num_groups)       # [n, 8h, 8w, 3]
        co = out.shape[-1]
        sample = ops.nhwc_to_ncfhw(out, n, co, 1).view(n, co, out.shape[1], out.shape[2])
        return DecoderOutput(sample=sample)
