"""Minimal model plumbing shared by the host-side mirrors of the reference's modules: config registration (the subset of
diffusers' ConfigMixin/ModelMixin behaviour the reference's scripts rely on: `.config`, attribute fall-through such as
`unet.in_channels`, `.dtype`, `.device`, `from_config`) and packed-weight caching."""
from __future__ import annotations

import inspect
import itertools
import json

import torch
from torch import nn


class FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ModelBase(nn.Module):
    """nn.Module + `.config` / `.dtype` / `.device` (mirrors diffusers ModelMixin+ConfigMixin as used by
    scripts/pose2vid.py:59-110 and pipeline_pose2vid_long.py:408,444)."""

    config_name = "config.json"
    # substrings of state-dict keys that from_pretrained() tolerates as absent from / unknown to the checkpoint
    _allow_missing_keys = ()
    _allow_unexpected_keys = ()

    def register_to_config(self, **kwargs):
        self.__dict__["_internal_dict"] = FrozenDict(kwargs)

    @property
    def config(self):
        return self.__dict__.get("_internal_dict", FrozenDict())

    def __getattr__(self, name):
        d = self.__dict__.get("_internal_dict")
        if d is not None and name in d:
            return d[name]
        return super().__getattr__(name)

    @property
    def device(self):
        for t in itertools.chain(self.parameters(), self.buffers()):
            return t.device
        return torch.device("cpu")

    @property
    def dtype(self):
        for t in itertools.chain(self.parameters(), self.buffers()):
            if t.is_floating_point():
                return t.dtype
        return torch.float32

    @classmethod
    def load_config(cls, path, **kwargs):
        with open(path) as f:
            return json.load(f)

    @classmethod
    def from_pretrained(cls, pretrained_model_path, subfolder=None, **kwargs):
        """diffusers-style loader: <path>/<subfolder>/config.json + diffusion_pytorch_model.{safetensors,bin}
        (scripts/pose2vid.py:59-64)."""
        import os
        path = os.path.join(str(pretrained_model_path), subfolder) if subfolder else str(pretrained_model_path)
        model = cls.from_config(cls.load_config(os.path.join(path, cls.config_name)), **kwargs)
        st = os.path.join(path, "diffusion_pytorch_model.safetensors")
        pt = os.path.join(path, "diffusion_pytorch_model.bin")
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st, device="cpu")
        elif os.path.exists(pt):
            sd = torch.load(pt, map_location="cpu", weights_only=True)
        else:
            raise FileNotFoundError(f"no weights file found in {path}")
        load_checked(model, sd, cls._allow_missing_keys, cls._allow_unexpected_keys)
        return model

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__).parameters
        init = {k: v for k, v in dict(config).items() if k in sig and not k.startswith("_")}
        init.update({k: v for k, v in kwargs.items() if k in sig})
        return cls(**init)


# diffusers < 0.19 checkpoints (sd-vae-ft-mse as published) name the VAE mid-block attention query/key/value/proj_attn;
# diffusers 0.24 remaps them in ModelMixin._convert_deprecated_attention_blocks [dep]. Same remap here.
_LEGACY_ATTENTION_KEYS = (("query", "to_q"), ("key", "to_k"), ("value", "to_v"), ("proj_attn", "to_out.0"))


def remap_legacy_attention_keys(state_dict: dict, target_keys) -> dict:
    out = {}
    target_keys = set(target_keys)
    for k, v in state_dict.items():
        nk = k
        if k not in target_keys:
            for old, new in _LEGACY_ATTENTION_KEYS:
                for leaf in ("weight", "bias"):
                    if k.endswith(f".{old}.{leaf}"):
                        cand = k[: -len(f"{old}.{leaf}")] + f"{new}.{leaf}"
                        if cand in target_keys:
                            nk = cand
        out[nk] = v
    return out


def load_checked(model: nn.Module, state_dict: dict, allow_missing=(), allow_unexpected=()):
    """load_state_dict that does not fail silently: legacy attention names are remapped, and any key that is missing from the
    checkpoint or unknown to the model raises unless it matches one of the allowed substrings (e.g. "motion_modules." when
    the 2-D SD1.5 weights are loaded into the 3-D UNet before the motion-module file is merged)."""
    own = model.state_dict().keys()
    sd = remap_legacy_attention_keys(state_dict, own)
    res = model.load_state_dict(sd, strict=False)
    missing = [k for k in res.missing_keys if not any(a in k for a in allow_missing)]
    unexpected = [k for k in res.unexpected_keys if not any(a in k for a in allow_unexpected)]
    if missing or unexpected:
        def head(keys):
            return ", ".join(keys[:6]) + (f", ... (+{len(keys) - 6})" if len(keys) > 6 else "")
        raise RuntimeError(f"{type(model).__name__}: checkpoint does not match the model: "
                           f"{len(missing)} missing [{head(missing)}]; {len(unexpected)} unexpected [{head(unexpected)}]")
    return res


class PackedCache:
    """Kernel-layout copies of a module's parameters, rebuilt whenever a parameter object is replaced / modified
    (load_state_dict, .to(), .half())."""

    def __init__(self):
        self._key = None
        self._val = None

    def get(self, module: nn.Module, build):
        key = tuple((p.data_ptr(), p._version, p.dtype, p.device) for p in module.parameters(recurse=True))
        if key != self._key:
            self._val = build()
            self._key = key
        return self._val


def f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def f16(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float16).contiguous()
