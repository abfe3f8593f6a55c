"""Seeded synthetic weights / inputs (there are no checkpoints or datasets offline). SURVEY.md §8d recipe: normal(0, 0.02)
weights, zero biases, unit norm scales, PoseGuider scale 2, sinusoidal PE buffers as computed."""
from __future__ import annotations

import math
import zlib

import torch


def randomize_state_dict(sd: dict, seed: int = 0, std: float = 0.02) -> dict:
    """Deterministically re-initialise a state dict. Every tensor draws from its own CPU generator seeded by
    (seed, crc32(name)), so the values do not depend on the key ORDER (the reference's modules register their children
    in a different order than ours, same keys)."""
    out = {}
    for name, t in sd.items():
        g = torch.Generator(device="cpu").manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
        if name.endswith(".pe"):  # positional-encoding buffers keep their analytic values
            out[name] = _sinusoid_pe(t.shape[1], t.shape[2]).to(t.dtype) if t.is_meta else t
            continue
        if not t.is_floating_point():
            out[name] = torch.zeros(t.shape, dtype=t.dtype) if t.is_meta else t
            continue
        leaf = name.rsplit(".", 1)[-1]
        is_norm = any(k in name for k in (".norm", "norm1", "norm2", "norm3", "ff_norm", "norms.", "group_norm",
                                          "conv_norm_out"))
        is_bn = t.dim() == 1 and ("conv_layers" in name)
        if name == "scale":
            v = torch.full(t.shape, 2.0)
        elif "running_mean" in name:
            v = torch.zeros(t.shape)
        elif "running_var" in name:
            v = torch.ones(t.shape)
        elif (is_norm or (is_bn and leaf == "weight" and t.dim() == 1 and _is_bn_name(name, sd))) and leaf == "weight":
            v = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
        elif leaf == "bias":
            v = 0.02 * torch.randn(t.shape, generator=g)
        else:
            fan_in = t[0].numel() if t.dim() > 1 else t.numel()
            v = torch.randn(t.shape, generator=g) * min(std * 2.5, 1.0 / math.sqrt(max(fan_in, 1)))
        out[name] = v.to(t.dtype)
    return out


def _sinusoid_pe(max_len, d_model):
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


def meta_state_dict(factory):
    """Shapes/dtypes of a model's state dict without allocating or initialising it."""
    with torch.device("meta"):
        return factory().state_dict()


def _is_bn_name(name, sd):
    base = name.rsplit(".", 1)[0]
    return (base + ".running_mean") in sd
