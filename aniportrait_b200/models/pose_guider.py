"""PoseGuider (mirror of the reference's src/models/pose_guider.py; same parameter names: conv_layers.N, final_proj,
conv_layers_{1..4}, cross_attn{1..4}.*, scale).

Semantics preserved on purpose:
  * the BatchNorm2d layers run with BATCH statistics (the reference never calls .eval(); scripts/pose2vid.py:102-110),
    so a frame's features depend on the window it is evaluated with;
  * the `ref_x` branch is dead code in the reference (its Transformer2DModel blocks are built with
    cross_attention_dim=None, pose_guider.py:86-89, so `ref_x` is never read): it is accepted and ignored.
Everything runs on this library's sm_100a kernels (aniportrait_b200/csrc): the 3/16/32-channel k3 / k4 convolutions of the
stem on a direct-convolution kernel (ap_conv2d_direct_nhwc_f16), every 64..1280-channel 3x3 convolution (stride 1 and 2) on
the tcgen05 implicit-GEMM kernel, train-mode BatchNorm + ReLU on a two-stage order-fixed reduction + apply
(ap_batchnorm_train_nhwc_f16), the 1x1 final projection (with `scale` folded in) on the GEMM kernel, and the four
self-attention Transformer2DModel blocks (16 heads x 88) on the GEMM / attention / norm kernels.
It does not depend on the timestep, so the pipeline evaluates it once per window instead of once per step.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from .blocks import RunCtx, Transformer2DModel as _KernelTransformer2D
from .modeling import ModelBase, PackedCache, f16, f32


class Transformer2DModel(_KernelTransformer2D):
    """pose_guider.py:181-308: 16 heads x 88, no cross attention."""

    def __init__(self, in_channels, num_attention_heads=16, attention_head_dim=88):
        super().__init__(num_attention_heads, attention_head_dim, in_channels, cross_attention_dim=None)


def _stage(cin, cout, strides):
    layers = []
    c = cin
    for i, s in enumerate(strides):
        co = cout if i == len(strides) - 1 else cin
        layers += [nn.Conv2d(c, co, 3, stride=s, padding=1), nn.BatchNorm2d(co), nn.ReLU()]
        c = co
    return nn.Sequential(*layers)


class PoseGuider(ModelBase):
    def __init__(self, noise_latent_channels=320, use_ca=True):
        super().__init__()
        c = noise_latent_channels
        self.use_ca = use_ca
        spec = [(3, 3, 3, 1), (3, 16, 4, 2), (16, 16, 3, 1), (16, 32, 4, 2), (32, 32, 3, 1), (32, 64, 4, 2),
                (64, 64, 3, 1), (64, 128, 3, 1)]
        layers = []
        for cin, cout, k, s in spec:
            layers += [nn.Conv2d(cin, cout, k, stride=s, padding=1), nn.BatchNorm2d(cout), nn.ReLU()]
        self.conv_layers = nn.Sequential(*layers)
        self.final_proj = nn.Conv2d(128, c, 1)
        self.conv_layers_1 = _stage(c, c, (1, 2))
        self.conv_layers_2 = _stage(c, 2 * c, (1, 2))
        self.conv_layers_3 = _stage(2 * c, 4 * c, (1, 2))
        self.conv_layers_4 = _stage(4 * c, 4 * c, (1,))
        if use_ca:
            self.cross_attn1 = Transformer2DModel(c)
            self.cross_attn2 = Transformer2DModel(2 * c)
            self.cross_attn3 = Transformer2DModel(4 * c)
            self.cross_attn4 = Transformer2DModel(4 * c)
        self._initialize_weights()
        self.scale = nn.Parameter(torch.ones(1) * 2)
        self.train()  # BatchNorm uses batch statistics, exactly like the (never .eval()'d) reference module

    def _initialize_weights(self):
        for block in (self.conv_layers, self.conv_layers_1, self.conv_layers_2, self.conv_layers_3, self.conv_layers_4):
            for m in block:
                if isinstance(m, nn.Conv2d):
                    n = m.kernel_size[0] * m.kernel_size[1] * m.in_channels
                    nn.init.normal_(m.weight, mean=0.0, std=(2.0 / n) ** 0.5)
                    nn.init.zeros_(m.bias)
        nn.init.zeros_(self.final_proj.weight)
        nn.init.zeros_(self.final_proj.bias)

    # ------------------------------------------------------------------------------------------------ packing
    @staticmethod
    def _pairs(seq):
        mods = list(seq)
        return [(mods[i], mods[i + 1]) for i in range(0, len(mods), 3)]     # (Conv2d, BatchNorm2d), ReLU follows

    def packed(self):
        if not hasattr(self, "_pk"):
            self._pk = PackedCache()

        def pad8(v, n, fill=0.0):
            out = torch.full((n,), fill, dtype=torch.float32, device=v.device)
            out[:v.numel()] = f32(v)
            return out

        def build():
            def layer(conv, bn, cin_have):
                """cin_have: channel count of the incoming channels-last buffer (>= conv.in_channels, zero padded)."""
                cin, cout, k, st = conv.in_channels, conv.out_channels, conv.kernel_size[0], conv.stride[0]
                if cin < 64:      # small-channel direct convolution; channels padded to multiples of 8 / 16
                    cout_p = 8 if cout <= 8 else (cout + 15) // 16 * 16
                    d = dict(kind="direct", w=ops.pack_conv_direct_weight(conv.weight.detach(), cin_have, cout_p),
                             b=pad8(conv.bias, cout_p), stride=st, cout=cout_p)
                else:             # tcgen05 implicit-GEMM 3x3
                    assert k == 3 and cin == cin_have
                    w = ops.pack_conv3x3_weight(conv.weight.detach())
                    cout_p = cout
                    d = dict(kind="igemm", w=w, b=pad8(conv.bias, w.shape[0]), stride=st, cout=cout)
                # padded channels: gamma = beta = 0 keeps them exactly zero through BatchNorm + ReLU
                d.update(g=pad8(bn.weight, cout_p), be=pad8(bn.bias, cout_p), eps=bn.eps)
                return d, cout_p
            stem, c = [], 8      # the pose map arrives as [frames, H, W, 8] (3 channels zero padded)
            for conv, bn in self._pairs(self.conv_layers):
                d, c = layer(conv, bn, c)
                stem.append(d)
            cfin = self.final_proj.out_channels
            sc = self.scale.detach().to(torch.float32)
            pk = dict(stem=stem,
                      # (W x + b) * scale == (scale W) x + scale b  (reference pose_guider.py:129-131)
                      wf=f16(self.final_proj.weight.detach().to(torch.float32).reshape(cfin, -1) * sc),
                      bf=(f32(self.final_proj.bias) * sc).contiguous(), stages=[])
            c = cfin
            for k in range(1, 5):
                st = []
                for conv, bn in self._pairs(getattr(self, f"conv_layers_{k}")):
                    d, c = layer(conv, bn, c)
                    st.append(d)
                pk["stages"].append(st)
            return pk
        return self._pk.get(self, build)

    @staticmethod
    def _conv_bn_relu(x, d):
        if d["kind"] == "direct":
            y = ops.conv2d_direct(x, d["w"], d["stride"], 1, bias=d["b"])
        else:
            y = ops.conv3x3(x, d["w"], d["cout"], bias=d["b"], stride=d["stride"])
        return ops.batch_norm_train(y, d["g"], d["be"], d["eps"], relu=True)

    @torch.no_grad()
    def forward_nhwc(self, x: torch.Tensor):
        """x: [frames, 3, H, W] -> 5 channels-last maps [frames, h, w, C] fp16 (shared by both CFG branches: duplicating
        the batch leaves BatchNorm's batch statistics unchanged)."""
        if not x.is_cuda:
            raise RuntimeError("aniportrait_b200.PoseGuider runs on CUDA only: no CPU fallback")
        if self.dtype != torch.float16:
            raise RuntimeError("aniportrait_b200.PoseGuider runs in fp16 (`.to(device, torch.float16)`): no fp32 / library path")
        pk = self.packed()
        frames, cin, H, W = x.shape
        ctx = RunCtx(1, frames, None, None, emit_stats=False)     # no GroupNorm consumes the transformer outputs here
        x = ops.ncfhw_to_nhwc(x.to(torch.float16).contiguous().view(frames, cin, 1, H, W), 8)      # [frames, H, W, 8]
        for d in pk["stem"]:
            x = self._conv_bn_relu(x, d)
        nf, h, w, c = x.shape
        x = ops.gemm(x.view(-1, c), pk["wf"], bias=pk["bf"]).view(nf, h, w, -1)
        fea = [x]
        for k in range(1, 5):
            for d in pk["stages"][k - 1]:
                x = self._conv_bn_relu(x, d)
            if self.use_ca:
                x = getattr(self, f"cross_attn{k}").run(x, ctx)
            fea.append(x)
        return fea

    def forward(self, x, ref_x=None):
        """Reference signature: x [B, 3, F, H, W] -> list of 5 maps [B, C, F, h, w] (pose_guider.py:124-162)."""
        b, _, f, _, _ = x.shape
        xs = x.permute(0, 2, 1, 3, 4).reshape(b * f, 3, x.shape[3], x.shape[4]).to(self.dtype)
        fea = self.forward_nhwc(xs)
        return [t.view(b, f, t.shape[1], t.shape[2], t.shape[3]).permute(0, 4, 1, 2, 3) for t in fea]
