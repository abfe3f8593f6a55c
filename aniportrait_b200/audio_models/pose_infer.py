"""Audio2PoseModel.infer without the O(T^2) re-decoding (reference src/audio_models/pose_model.py:97-124).

The reference predicts the head pose of frame i by running its 8-layer nn.TransformerDecoder over ALL i+1 pose tokens
produced so far, for i = 0 .. T-1 (T = 150..299 per 5 s audio chunk, scripts/audio2vid.py:176-195): T(T+1)/2 token passes,
of which only the last row of every pass is new. Three facts of that loop make it incremental:

* the self-attention mask (causal + ALiBi, `biased_mask`) lets token i see tokens <= i only, so the activations of earlier
  tokens never change once computed: their per-layer keys / values can be cached;
* the cross-attention mask (`enc_dec_mask`) leaves token i exactly ONE audio frame, frame i: the softmax is over a single
  key, its weight is 1, and the cross-attention output is out_proj(v_proj(audio_i)) — independent of the pose tokens.
  It is evaluated for all T frames and all layers with two GEMMs per layer before the loop;
* the positional encoding and the identity embedding are per-position additions.

What remains per frame is one token through 8 layers: q/k/v of the new token, attention of one query over the cache, two
small MLPs. Every step has the same tensor shapes (the cache is allocated for T tokens, entries beyond the current position are
masked by the -inf the reference's own mask already holds there; position-dependent reads / writes go through an index
TENSOR), so on a GPU one step is captured into a CUDA graph once and replayed T times — the step is launch-bound (~100 small
kernels), not FLOP-bound. The arithmetic is torch library code on the module's own parameters and mask buffer (nothing of the
reference model is restated or copied: `enable_kv_cache(model)` rebinds `infer` on an instance of the reference class).

Parity: tests/test_host_cpu.py::test_kv_cached_pose_infer_matches_reference_infer (unmodified reference module, seeded
weights, CPU fp32: <= 1e-4 relative). Not measured on the GPU (no budget left when it was written); graph capture falls back to
eager stepping with a warning if capture fails.
"""
from __future__ import annotations

import math
import types
import warnings

import torch
import torch.nn.functional as F


def _layer_step(layer, x, pos, cache_k, cache_v, mask_row, cross_row, heads):
    """One post-norm nn.TransformerDecoderLayer (eval) for ONE new token x [1, E] at position `pos` ([1] int64 tensor)."""
    E = x.shape[-1]
    d = E // heads
    sa = layer.self_attn
    q, k, v = F.linear(x, sa.in_proj_weight, sa.in_proj_bias).view(3, heads, d)
    cache_k.index_copy_(1, pos, k.unsqueeze(1))                   # [heads, T, d]
    cache_v.index_copy_(1, pos, v.unsqueeze(1))
    scores = torch.bmm(cache_k, q.unsqueeze(-1)).squeeze(-1) * (1.0 / math.sqrt(d)) + mask_row      # [heads, T]
    p = torch.softmax(scores, dim=-1)
    a = torch.bmm(p.unsqueeze(1), cache_v).reshape(1, E)
    x = layer.norm1(x + sa.out_proj(a))
    x = layer.norm2(x + cross_row)
    x = layer.norm3(x + layer.linear2(layer.activation(layer.linear1(x))))
    return x


@torch.no_grad()
def kv_cached_infer(model, input_value, seq_len, id_seed=None, use_cuda_graph=None):
    """Drop-in for `Audio2PoseModel.infer(input_value, seq_len, id_seed)` -> [1, seq_len, out_dim] (reference :97-124)."""
    emb = model.audio_encoder(input_value, seq_len=seq_len, output_hidden_states=True)
    if model._only_last_features:
        hidden = emb.last_hidden_state
    else:
        hidden = sum(emb.hidden_states) / len(emb.hidden_states)
    hidden = model.in_fn(hidden)                                       # [1, S, E] audio memory
    if hidden.shape[0] != 1:
        raise ValueError("Audio2PoseModel.infer works on one clip at a time (its attention mask is built for batch 1)")
    T = int(seq_len)
    layers = list(model.transformer_decoder.layers)
    if model.transformer_decoder.norm is not None or any(getattr(l, "norm_first", False) for l in layers):
        raise NotImplementedError("kv_cached_infer mirrors the post-norm decoder the reference builds")
    dev, dt = hidden.device, hidden.dtype
    heads = layers[0].self_attn.num_heads
    E = hidden.shape[-1]
    if hidden.shape[1] < T:
        raise ValueError(f"audio memory has {hidden.shape[1]} frames, fewer than seq_len={T}")
    # cross-attention with a one-key softmax: out_proj(v_proj(audio_i)), all frames at once, per layer  -> [layers, T, E]
    mem = hidden[0, :T]
    cross = []
    for l in layers:
        ca = l.multihead_attn
        cross.append(ca.out_proj(F.linear(mem, ca.in_proj_weight[2 * E:], ca.in_proj_bias[2 * E:])))
    cross = torch.stack(cross)
    # per-position additions: sinusoidal table + identity embedding; the reference's own causal + ALiBi mask rows
    add = model.PPE.pe[0, :T].to(dev, dt) + model.id_embed(id_seed).to(dt)                # [T, E]
    mask = model.biased_mask[:, :T, :T].to(dev, dt)                                       # [heads, T(query), T(key)]
    cache_k = [torch.zeros(heads, T, E // heads, device=dev, dtype=dt) for _ in layers]
    cache_v = [torch.zeros(heads, T, E // heads, device=dev, dtype=dt) for _ in layers]
    out = torch.zeros(T, model.out_dim, device=dev, dtype=dt)
    pos = torch.zeros(1, dtype=torch.long, device=dev)
    token = model.pose_map(torch.zeros(1, model.out_dim, device=dev, dtype=dt))           # embedding of the zero pose

    def step():
        x = token + add.index_select(0, pos)
        row = mask.index_select(1, pos).squeeze(1)
        for li, l in enumerate(layers):
            x = _layer_step(l, x, pos, cache_k[li], cache_v[li], row, cross[li].index_select(0, pos), heads)
        pose = model.pose_map_r(x)
        out.index_copy_(0, pos, pose)
        token.copy_(model.pose_map(pose))
        pos.add_(1)

    graph = None
    if use_cuda_graph is None:
        use_cuda_graph = dev.type == "cuda"
    if use_cuda_graph and dev.type == "cuda" and T > 2:
        step()                                                   # position 0 eagerly (warms up every kernel)
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):                        # capturing does not execute: position 1 is the first replay
                step()
        except Exception as exc:                                  # noqa: BLE001
            warnings.warn(f"kv_cached_infer: CUDA-graph capture failed ({exc!r}); stepping eagerly")
            graph = None
        for _ in range(1, T):
            graph.replay() if graph is not None else step()
    else:
        for _ in range(T):
            step()
    return out.unsqueeze(0)


def enable_kv_cache(model, use_cuda_graph=None):
    """Rebind `infer` on an instance of the reference's Audio2PoseModel (scripts/audio2vid.py:70-72 builds it)."""
    def infer(self, input_value, seq_len, id_seed=None):
        return kv_cached_infer(self, input_value, seq_len, id_seed, use_cuda_graph=use_cuda_graph)
    model.infer = types.MethodType(infer, model)
    return model
