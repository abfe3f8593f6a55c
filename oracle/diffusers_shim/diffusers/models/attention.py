"""diffusers 0.24.0 models/attention.py restated: FeedForward (GEGLU), GELU, AdaLayerNorm stubs."""
from typing import Any, Dict, Optional  # noqa: F401

import torch
import torch.nn.functional as F
from torch import nn

from .attention_processor import Attention  # noqa: F401
from .embeddings import SinusoidalPositionalEmbedding  # noqa: F401


class GELU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int, approximate: str = "none"):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)
        self.approximate = approximate

    def forward(self, hidden_states):
        return F.gelu(self.proj(hidden_states), approximate=self.approximate)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, hidden_states, scale: float = 1.0):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int, dim_out: Optional[int] = None, mult: int = 4, dropout: float = 0.0,
                 activation_fn: str = "geglu", final_dropout: bool = False):
        super().__init__()
        inner_dim = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        if activation_fn == "gelu":
            act_fn = GELU(dim, inner_dim)
        elif activation_fn == "gelu-approximate":
            act_fn = GELU(dim, inner_dim, approximate="tanh")
        elif activation_fn == "geglu":
            act_fn = GEGLU(dim, inner_dim)
        else:
            raise ValueError(activation_fn)
        self.net = nn.ModuleList([])
        self.net.append(act_fn)
        self.net.append(nn.Dropout(dropout))
        self.net.append(nn.Linear(inner_dim, dim_out))
        if final_dropout:
            self.net.append(nn.Dropout(dropout))

    def forward(self, hidden_states, scale: float = 1.0):
        for module in self.net:
            hidden_states = module(hidden_states)
        return hidden_states


def _stub(name):
    class _S(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            raise NotImplementedError(f"{name} is not on the AniPortrait hot path")
    _S.__name__ = name
    return _S


AdaLayerNorm = _stub("AdaLayerNorm")
AdaLayerNormZero = _stub("AdaLayerNormZero")
GatedSelfAttentionDense = _stub("GatedSelfAttentionDense")
