"""diffusers 0.24.0 models/attention_processor.py restated: Attention, AttnProcessor, AttnProcessor2_0.

Semantics kept (the ones the AniPortrait hot path exercises, reference src/models/attention.py:323-347,
src/models/motion_module.py:280-388, src/models/mutual_self_attention.py:158-205):
  to_q: Linear(query_dim, inner, bias=bias); to_k/to_v: Linear(cross_attention_dim or query_dim, inner, bias=bias);
  to_out = [Linear(inner, query_dim, bias=out_bias), Dropout]; scale = dim_head**-0.5; residual_connection False;
  rescale_output_factor 1; default processor = AttnProcessor2_0 when F.scaled_dot_product_attention exists.
"""
from typing import Callable, Optional, Union  # noqa: F401  (re-exported through the reference's star import)

import torch
import torch.nn.functional as F  # noqa: F401
from torch import nn


class Attention(nn.Module):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 dropout: float = 0.0, bias=False, upcast_attention: bool = False, upcast_softmax: bool = False,
                 cross_attention_norm: Optional[str] = None, cross_attention_norm_num_groups: int = 32,
                 added_kv_proj_dim: Optional[int] = None, norm_num_groups: Optional[int] = None,
                 spatial_norm_dim: Optional[int] = None, out_bias: bool = True, scale_qk: bool = True,
                 only_cross_attention: bool = False, eps: float = 1e-5, rescale_output_factor: float = 1.0,
                 residual_connection: bool = False, _from_deprecated_attn_block=False, processor=None):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.upcast_attention = upcast_attention
        self.upcast_softmax = upcast_softmax
        self.rescale_output_factor = rescale_output_factor
        self.residual_connection = residual_connection
        self.dropout = dropout
        self._from_deprecated_attn_block = _from_deprecated_attn_block
        self.scale_qk = scale_qk
        self.scale = dim_head ** -0.5 if self.scale_qk else 1.0
        self.heads = heads
        self.sliceable_head_dim = heads
        self.added_kv_proj_dim = added_kv_proj_dim
        self.only_cross_attention = only_cross_attention
        self.group_norm = (nn.GroupNorm(num_channels=query_dim, num_groups=norm_num_groups, eps=eps, affine=True)
                           if norm_num_groups is not None else None)
        self.spatial_norm = None
        self.norm_cross = None
        assert cross_attention_norm is None and added_kv_proj_dim is None and spatial_norm_dim is None
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])
        if processor is None:
            processor = AttnProcessor2_0() if hasattr(F, "scaled_dot_product_attention") and self.scale_qk \
                else AttnProcessor()
        self.set_processor(processor)

    def set_use_memory_efficient_attention_xformers(self, use, attention_op=None):
        raise NotImplementedError("xformers is not part of the oracle")

    def set_processor(self, processor, _remove_lora=False):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)

    def batch_to_head_dim(self, tensor):
        head_size = self.heads
        batch_size, seq_len, dim = tensor.shape
        tensor = tensor.reshape(batch_size // head_size, head_size, seq_len, dim)
        return tensor.permute(0, 2, 1, 3).reshape(batch_size // head_size, seq_len, dim * head_size)

    def head_to_batch_dim(self, tensor, out_dim=3):
        head_size = self.heads
        batch_size, seq_len, dim = tensor.shape
        tensor = tensor.reshape(batch_size, seq_len, head_size, dim // head_size).permute(0, 2, 1, 3)
        if out_dim == 3:
            tensor = tensor.reshape(batch_size * head_size, seq_len, dim // head_size)
        return tensor

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        if self.upcast_attention:
            query = query.float()
            key = key.float()
        if attention_mask is None:
            baddbmm_input = torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype,
                                        device=query.device)
            beta = 0
        else:
            baddbmm_input = attention_mask
            beta = 1
        attention_scores = torch.baddbmm(baddbmm_input, query, key.transpose(-1, -2), beta=beta, alpha=self.scale)
        if self.upcast_softmax:
            attention_scores = attention_scores.float()
        attention_probs = attention_scores.softmax(dim=-1)
        return attention_probs.to(dtype)

    def prepare_attention_mask(self, attention_mask, target_length, batch_size, out_dim=3):
        if attention_mask is None:
            return attention_mask
        raise NotImplementedError("attention masks are not used on the AniPortrait hot path")

    def norm_encoder_hidden_states(self, encoder_hidden_states):
        return encoder_hidden_states


class AttnProcessor:
    """Classic bmm-softmax-bmm processor (used by the motion module when set_use_memory_efficient... is called)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0):
        residual = hidden_states
        input_ndim = hidden_states.ndim
        assert input_ndim == 3
        batch_size, sequence_length, _ = (hidden_states.shape if encoder_hidden_states is None
                                          else encoder_hidden_states.shape)
        attention_mask = attn.prepare_attention_mask(attention_mask, sequence_length, batch_size)
        query = attn.to_q(hidden_states)
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        key = attn.to_k(encoder_hidden_states)
        value = attn.to_v(encoder_hidden_states)
        query = attn.head_to_batch_dim(query)
        key = attn.head_to_batch_dim(key)
        value = attn.head_to_batch_dim(value)
        attention_probs = attn.get_attention_scores(query, key, attention_mask)
        hidden_states = torch.bmm(attention_probs, value)
        hidden_states = attn.batch_to_head_dim(hidden_states)
        hidden_states = attn.to_out[0](hidden_states)
        hidden_states = attn.to_out[1](hidden_states)
        if attn.residual_connection:
            hidden_states = hidden_states + residual
        return hidden_states / attn.rescale_output_factor


class AttnProcessor2_0:
    """scaled_dot_product_attention processor (the default on torch >= 2.0)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0):
        residual = hidden_states
        input_ndim = hidden_states.ndim
        assert input_ndim == 3
        batch_size, sequence_length, _ = (hidden_states.shape if encoder_hidden_states is None
                                          else encoder_hidden_states.shape)
        assert attention_mask is None
        query = attn.to_q(hidden_states)
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        key = attn.to_k(encoder_hidden_states)
        value = attn.to_v(encoder_hidden_states)
        inner_dim = key.shape[-1]
        head_dim = inner_dim // attn.heads
        query = query.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        key = key.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        value = value.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        hidden_states = F.scaled_dot_product_attention(query, key, value, attn_mask=None, dropout_p=0.0,
                                                       is_causal=False)
        hidden_states = hidden_states.transpose(1, 2).reshape(batch_size, -1, attn.heads * head_dim)
        hidden_states = hidden_states.to(query.dtype)
        hidden_states = attn.to_out[0](hidden_states)
        hidden_states = attn.to_out[1](hidden_states)
        if attn.residual_connection:
            hidden_states = hidden_states + residual
        return hidden_states / attn.rescale_output_factor


class XFormersAttnProcessor:
    def __init__(self, attention_op=None):
        raise NotImplementedError("xformers is not part of the oracle")


class AttnAddedKVProcessor:
    pass


class AttnAddedKVProcessor2_0:
    pass


class LoRAAttnProcessor:
    pass


class LoRAAttnProcessor2_0:
    pass


class LoRAXFormersAttnProcessor:
    pass


class LoRAAttnAddedKVProcessor:
    pass


class CustomDiffusionAttnProcessor:
    pass


class CustomDiffusionAttnProcessor2_0:
    pass


class CustomDiffusionXFormersAttnProcessor:
    pass


class SlicedAttnProcessor:
    pass


class SlicedAttnAddedKVProcessor:
    pass


class XFormersAttnAddedKVProcessor:
    pass


ADDED_KV_ATTENTION_PROCESSORS = (AttnAddedKVProcessor, SlicedAttnAddedKVProcessor, AttnAddedKVProcessor2_0,
                                 XFormersAttnAddedKVProcessor, LoRAAttnAddedKVProcessor)
CROSS_ATTENTION_PROCESSORS = (AttnProcessor, AttnProcessor2_0, XFormersAttnProcessor, SlicedAttnProcessor,
                              LoRAAttnProcessor, LoRAAttnProcessor2_0, LoRAXFormersAttnProcessor)
LORA_ATTENTION_PROCESSORS = (LoRAAttnProcessor, LoRAAttnProcessor2_0, LoRAXFormersAttnProcessor,
                             LoRAAttnAddedKVProcessor)
AttentionProcessor = Union[AttnProcessor, AttnProcessor2_0]
