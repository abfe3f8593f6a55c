import logging as _pylogging
from collections import OrderedDict
from dataclasses import fields, is_dataclass

import torch
from packaging import version

SAFETENSORS_WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"
WEIGHTS_NAME = "diffusion_pytorch_model.bin"
USE_PEFT_BACKEND = True  # -> plain nn.Linear / nn.Conv2d (identical math to LoRACompatible* with no LoRA attached)


class BaseOutput(OrderedDict):
    """Dataclass-style output container (diffusers.utils.BaseOutput)."""

    def __post_init__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                self[f.name] = v

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


class _Logging:
    @staticmethod
    def get_logger(name=None):
        return _pylogging.getLogger(name)


logging = _Logging()


def deprecate(*args, **kwargs):
    return None


def is_accelerate_available():
    return False


def is_torch_version(op, ver):
    import operator
    ops = {">": operator.gt, ">=": operator.ge, "<": operator.lt, "<=": operator.le, "==": operator.eq}
    return ops[op](version.parse(torch.__version__.split("+")[0]), version.parse(ver))


def scale_lora_layers(model, weight):
    return None


def unscale_lora_layers(model, weight=None):
    return None
