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
        model.load_state_dict(sd, strict=False)
        return model

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__).parameters
        init = {k: v for k, v in dict(config).items() if k in sig and not k.startswith("_")}
        init.update({k: v for k, v in kwargs.items() if k in sig})
        return cls(**init)


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
