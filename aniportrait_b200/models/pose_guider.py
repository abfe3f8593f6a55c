"""PoseGuider (mirror of the reference's src/models/pose_guider.py; same parameter names: conv_layers.N, final_proj,
conv_layers_{1..4}, cross_attn{1..4}.*, scale).

Semantics preserved on purpose:
  * the BatchNorm2d layers run with BATCH statistics (the reference never calls .eval(); scripts/pose2vid.py:102-110),
    so a frame's features depend on the window it is evaluated with;
  * the `ref_x` branch is dead code in the reference (its Transformer2DModel blocks are built with
    cross_attention_dim=None, pose_guider.py:86-89, so `ref_x` is never read): it is accepted and ignored.
The four self-attention Transformer2DModel blocks (16 heads x 88) run on the sm_100a kernels; the small-channel conv /
BatchNorm / ReLU stem (3..128 channels, 4x4 stride-2 convs, ~2.5 GFLOP/frame) currently uses fp16 torch library ops.
It does not depend on the timestep, so the pipeline evaluates it once per window instead of once per step.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .blocks import RunCtx, Transformer2DModel as _KernelTransformer2D
from .modeling import ModelBase


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

    @staticmethod
    def _bn_stage(seq, x):
        for m in seq:
            if isinstance(m, nn.BatchNorm2d):
                # batch statistics, biased variance; running stats are not part of the forward value
                x = F.batch_norm(x, None, None, m.weight, m.bias, training=True, momentum=0.0, eps=m.eps)
            else:
                x = m(x)
        return x

    @torch.no_grad()
    def forward_nhwc(self, x: torch.Tensor):
        """x: [frames, 3, H, W] -> 5 channels-last maps [frames, h, w, C] fp16 (shared by both CFG branches: duplicating
        the batch leaves BatchNorm's batch statistics unchanged)."""
        if not x.is_cuda:
            raise RuntimeError("aniportrait_b200.PoseGuider runs on CUDA only: no CPU fallback")
        ctx = RunCtx(1, x.shape[0], None, None)
        fea = []
        x = self._bn_stage(self.conv_layers, x)
        x = self.final_proj(x) * self.scale
        x = x.permute(0, 2, 3, 1).contiguous()
        fea.append(x)
        for k in range(1, 5):
            xc = self._bn_stage(getattr(self, f"conv_layers_{k}"), x.permute(0, 3, 1, 2))
            x = xc.permute(0, 2, 3, 1).contiguous()
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
