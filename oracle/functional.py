"""TEST INFRASTRUCTURE — CPU oracle of AniPortrait's denoising hot path (plain torch, fp32 by default).

A flat functional restatement, keyed by the reference's state-dict names, of exactly what the reference executes on
this path. Each function cites the reference file:line it follows (paths relative to Zejun-Yang/AniPortrait). The
leaf semantics of diffusers==0.24.0 (Attention/SDPA, FeedForward-GEGLU, Timesteps, ResnetBlock2D, DDIM, AutoencoderKL)
are restated from that pinned version, which is NOT vendored in the reference; the oracle is pinned against the
reference's own wiring imported unmodified through oracle/diffusers_shim (tests/test_oracle_vs_reference.py, golden
vectors in tests/golden/) — "parity unpinned" with respect to real diffusers/torch-2.0.1 library numerics.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module. The
product (aniportrait_b200/) never does.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

SD15 = dict(block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, heads=8, norm_groups=32, norm_eps=1e-5,
            cross_attention_dim=768, in_channels=4, out_channels=4, motion_max_len=32)


# The checker evaluates attention as explicit softmax(q k^T d^-1/2) v. The reference itself reaches
# F.scaled_dot_product_attention through diffusers' AttnProcessor2_0 [dep]; bench.py's CPU-baseline leg sets USE_SDPA so that
# the TIMED restatement runs the same library kernel the reference would (6x faster on CPU at 4096 x 8192 tokens; values agree
# to 1e-7, tests/test_oracle_vs_reference.py runs both).
USE_SDPA = False


# ------------------------------------------------------------------------------------------------ leaves
def linear(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def conv2d(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def group_norm(sd, p, x, groups, eps):
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def layer_norm(sd, p, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def attention(sd, p, x, kv, heads):
    """diffusers Attention + AttnProcessor2_0 [dep 0.24.0]: no qkv bias, out bias, scale d^-1/2, softmax over keys."""
    q = F.linear(x, sd[p + ".to_q.weight"])
    k = F.linear(kv, sd[p + ".to_k.weight"])
    v = F.linear(kv, sd[p + ".to_v.weight"])
    b, n, c = q.shape
    d = c // heads
    q = q.view(b, n, heads, d).transpose(1, 2)
    k = k.view(b, -1, heads, d).transpose(1, 2)
    v = v.view(b, -1, heads, d).transpose(1, 2)
    if USE_SDPA:
        o = F.scaled_dot_product_attention(q, k, v)
    else:
        s = (q @ k.transpose(-1, -2)) * (d ** -0.5)
        o = s.softmax(dim=-1) @ v
    o = o.transpose(1, 2).reshape(b, n, c)
    return F.linear(o, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def feed_forward(sd, p, x):
    """diffusers FeedForward(activation_fn='geglu') [dep]: proj -> (h, gate) -> h * gelu_erf(gate) -> Linear."""
    h, g = linear(sd, p + ".net.0.proj", x).chunk(2, dim=-1)
    return linear(sd, p + ".net.2", h * F.gelu(g))


def timestep_embedding(t, dim=320):
    """diffusers Timesteps(320, flip_sin_to_cos=True, freq_shift=0) [dep]; reference src/models/unet_3d.py:95,463."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


# ------------------------------------------------------------------------------------------------ blocks
def resnet_block(sd, p, x, temb, groups=32, eps=1e-5):
    """ResnetBlock3D.forward (src/models/resnet.py:218-248) on [(b f), C, H, W]; temb already per frame [(b f), 1280].
    Identical math to diffusers ResnetBlock2D used by the ReferenceNet."""
    h = F.silu(group_norm(sd, p + ".norm1", x, groups, eps))
    h = conv2d(sd, p + ".conv1", h)
    if temb is not None:
        h = h + linear(sd, p + ".time_emb_proj", F.silu(temb))[:, :, None, None]
    h = F.silu(group_norm(sd, p + ".norm2", h, groups, eps))
    h = conv2d(sd, p + ".conv2", h)
    if (p + ".conv_shortcut.weight") in sd:
        x = conv2d(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def transformer_block(sd, p, x, ehs, heads, bank=None, uncond_frames=0, write_bank=None):
    """(Temporal)BasicTransformerBlock as patched by ReferenceAttentionControl
    (src/models/mutual_self_attention.py:93-265). x: [frames, N, C]; ehs: [frames, S, 768].
    write mode (write_bank is a list): append norm1(x) (:137-138), plain self-attention.
    read mode (bank [frames, N, C] or None): keys/values = cat([norm1(x), bank]) (:147-165); the first `uncond_frames`
    frames are recomputed with plain self-attention (:166-186)."""
    n1 = layer_norm(sd, p + ".norm1", x)
    if write_bank is not None:
        write_bank.append(n1.clone())
    if bank is not None:
        out = attention(sd, p + ".attn1", n1, torch.cat([n1, bank], dim=1), heads) + x
        if uncond_frames > 0:
            u = slice(0, uncond_frames)
            out = out.clone()
            out[u] = attention(sd, p + ".attn1", n1[u], n1[u], heads) + x[u]
        x = out
    else:
        x = attention(sd, p + ".attn1", n1, n1, heads) + x
    if (p + ".attn2.to_q.weight") in sd:
        x = attention(sd, p + ".attn2", layer_norm(sd, p + ".norm2", x), ehs, heads) + x
    return feed_forward(sd, p + ".ff", layer_norm(sd, p + ".norm3", x)) + x


def spatial_transformer(sd, p, x, ehs, heads, groups=32, **kw):
    """Transformer3DModel.forward (src/models/transformer_3d.py:103-169) / Transformer2DModel
    (src/models/transformer_2d.py:213-396) on [(b f), C, H, W]: GN(eps 1e-6) -> 1x1 -> block -> 1x1 -> + residual."""
    b, c, h, w = x.shape
    t = group_norm(sd, p + ".norm", x, groups, 1e-6)
    t = conv2d(sd, p + ".proj_in", t, padding=0)
    t = t.permute(0, 2, 3, 1).reshape(b, h * w, -1)
    t = transformer_block(sd, p + ".transformer_blocks.0", t, ehs, heads, **kw)
    t = t.reshape(b, h, w, -1).permute(0, 3, 1, 2)
    return conv2d(sd, p + ".proj_out", t, padding=0) + x


def positional_encoding(d_model, max_len=32):
    """PositionalEncoding buffer (src/models/motion_module.py:262-277)."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def motion_module(sd, p, x, video_length, heads=8, groups=32):
    """VanillaTemporalModule -> TemporalTransformer3DModel.forward (src/models/motion_module.py:146-182),
    TemporalTransformerBlock.forward (:236-259), VersatileAttention.forward (:351-388). x: [(b f), C, H, W]."""
    p = p + ".temporal_transformer"
    bf, c, h, w = x.shape
    b = bf // video_length
    t = group_norm(sd, p + ".norm", x, groups, 1e-6)
    t = t.permute(0, 2, 3, 1).reshape(bf, h * w, c)
    t = linear(sd, p + ".proj_in", t)
    blk = p + ".transformer_blocks.0"
    pe_key = blk + ".attention_blocks.0.pos_encoder.pe"
    pe = sd[pe_key][0] if pe_key in sd else positional_encoding(c)
    for i in range(2):
        n = layer_norm(sd, f"{blk}.norms.{i}", t)
        # "(b f) d c -> (b d) f c"
        n = n.view(b, video_length, h * w, c).permute(0, 2, 1, 3).reshape(b * h * w, video_length, c)
        n = n + pe[None, :video_length].to(n.dtype)
        a = attention(sd, f"{blk}.attention_blocks.{i}", n, n, heads)
        a = a.view(b, h * w, video_length, c).permute(0, 2, 1, 3).reshape(bf, h * w, c)
        t = a + t
    t = feed_forward(sd, blk + ".ff", layer_norm(sd, blk + ".ff_norm", t)) + t
    t = linear(sd, p + ".proj_out", t)
    return t.reshape(bf, h, w, c).permute(0, 3, 1, 2) + x


# ------------------------------------------------------------------------------------------------ UNets
def _time_embed(sd, t, batch, dtype):
    temb = timestep_embedding(t.expand(batch), sd["time_embedding.linear_1.weight"].shape[1]).to(dtype)
    temb = linear(sd, "time_embedding.linear_1", temb)
    return linear(sd, "time_embedding.linear_2", F.silu(temb))


def unet3d_forward(sd, sample, timestep, ehs, pose_fea=None, banks=None, cfg=True, use_motion=True, c=SD15):
    """UNet3DConditionModel.forward (src/models/unet_3d.py:399-580) with ReferenceAttentionControl in read mode.
    sample [B, 4, F, h, w]; ehs [B, 1, 768]; pose_fea: 5 tensors [B, C, F, h', w'] or None;
    banks: list of 16 tensors [B, N, C] in reader DFS order (unsorted; pairing is positional per width, see
    reference_banks()) or None (plain self-attention)."""
    B, _, Fr, H, W = sample.shape
    heads, groups, eps = c["heads"], c["norm_groups"], c["norm_eps"]
    t = timestep if torch.is_tensor(timestep) else torch.tensor([timestep])
    t = t.reshape(-1)
    temb = _time_embed(sd, t, B, sample.dtype).repeat_interleave(Fr, dim=0)           # per frame
    ehs_f = ehs.repeat_interleave(Fr, dim=0)                                          # "(b f) n c"
    x = sample.permute(0, 2, 1, 3, 4).reshape(B * Fr, -1, H, W)                       # "(b f) c h w"
    uncond_frames = (B // 2) * Fr if cfg else 0
    bank_iter = iter(banks) if banks is not None else None

    def to_bf(pf):
        return pf.permute(0, 2, 1, 3, 4).reshape(B * Fr, pf.shape[1], pf.shape[3], pf.shape[4])

    def tr(p, x):
        kw = {}
        if bank_iter is not None:
            bk = next(bank_iter)                                                     # [B, N, C]
            kw = dict(bank=bk.repeat_interleave(Fr, dim=0), uncond_frames=uncond_frames)
        return spatial_transformer(sd, p, x, ehs_f, heads, groups, **kw)

    def mm(p, x):
        return motion_module(sd, p, x, Fr, heads, groups) if use_motion and (p + ".temporal_transformer.norm.weight") in sd else x

    x = conv2d(sd, "conv_in", x)
    if pose_fea is not None:
        x = x + to_bf(pose_fea[0])
    skips = [x]
    for i in range(4):
        for j in range(c["layers_per_block"]):
            x = resnet_block(sd, f"down_blocks.{i}.resnets.{j}", x, temb, groups, eps)
            if i < 3:
                x = tr(f"down_blocks.{i}.attentions.{j}", x)
            x = mm(f"down_blocks.{i}.motion_modules.{j}", x)
            skips.append(x)
        if i < 3:
            x = conv2d(sd, f"down_blocks.{i}.downsamplers.0.conv", x, stride=2)
            skips.append(x)
        if pose_fea is not None:
            x = x + to_bf(pose_fea[i + 1])
    x = resnet_block(sd, "mid_block.resnets.0", x, temb, groups, eps)
    x = tr("mid_block.attentions.0", x)
    x = mm("mid_block.motion_modules.0", x)
    x = resnet_block(sd, "mid_block.resnets.1", x, temb, groups, eps)
    for i in range(4):
        for j in range(c["layers_per_block"] + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block(sd, f"up_blocks.{i}.resnets.{j}", x, temb, groups, eps)
            if i > 0:
                x = tr(f"up_blocks.{i}.attentions.{j}", x)
            x = mm(f"up_blocks.{i}.motion_modules.{j}", x)
        if i < 3:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = conv2d(sd, f"up_blocks.{i}.upsamplers.0.conv", x)
    x = F.silu(group_norm(sd, "conv_norm_out", x, groups, eps))
    x = conv2d(sd, "conv_out", x)
    return x.reshape(B, Fr, -1, H, W).permute(0, 2, 1, 3, 4)


def reference_unet_banks(sd, latent, ehs, c=SD15):
    """ReferenceNet write pass: UNet2DConditionModel.forward (src/models/unet_2d_condition.py:872-1308, conv_out
    removed :645-653) at t=0 with every BasicTransformerBlock in write mode. latent [B,4,h,w]; ehs [B,1,768].
    Returns the 16 bank tensors [B, N, C] in DFS (execution) order."""
    B = latent.shape[0]
    heads, groups, eps = c["heads"], c["norm_groups"], c["norm_eps"]
    temb = _time_embed(sd, torch.zeros(1, dtype=torch.long), B, latent.dtype)
    banks = []

    def tr(p, x):
        return spatial_transformer(sd, p, x, ehs, heads, groups, write_bank=banks)

    x = conv2d(sd, "conv_in", latent)
    skips = [x]
    for i in range(4):
        for j in range(c["layers_per_block"]):
            x = resnet_block(sd, f"down_blocks.{i}.resnets.{j}", x, temb, groups, eps)
            if i < 3:
                x = tr(f"down_blocks.{i}.attentions.{j}", x)
            skips.append(x)
        if i < 3:
            x = conv2d(sd, f"down_blocks.{i}.downsamplers.0.conv", x, stride=2)
            skips.append(x)
    x = resnet_block(sd, "mid_block.resnets.0", x, temb, groups, eps)
    x = tr("mid_block.attentions.0", x)
    x = resnet_block(sd, "mid_block.resnets.1", x, temb, groups, eps)
    for i in range(4):
        for j in range(c["layers_per_block"] + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block(sd, f"up_blocks.{i}.resnets.{j}", x, temb, groups, eps)
            if i > 0:
                x = tr(f"up_blocks.{i}.attentions.{j}", x)
        if i < 3:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = conv2d(sd, f"up_blocks.{i}.upsamplers.0.conv", x)
    return banks


def pair_banks(banks):
    """ReferenceAttentionControl.update (src/models/mutual_self_attention.py:302-339): reader and writer block lists are
    both stably sorted by descending width and zipped. Both UNets share one DFS topology, so the pairing is positional:
    the k-th block of the reader receives the k-th bank of the writer. (Identity, kept to document the argument.)"""
    return list(banks)


# ------------------------------------------------------------------------------------------------ PoseGuider
def _bn_train(sd, p, x, eps=1e-5):
    """nn.BatchNorm2d in TRAIN mode (the scripts never call .eval(): scripts/pose2vid.py:102-110): batch statistics,
    biased variance (src/models/pose_guider.py:19-89)."""
    # y = (x - mean_batch) / sqrt(var_batch_biased + eps) * gamma + beta
    return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], training=True, momentum=0.0, eps=eps)


def pose_guider_forward(sd, x, c0=320):
    """PoseGuider.forward (src/models/pose_guider.py:124-162). x: [B, 3, F, H, W]. The ref_x branch is dead code
    (cross_attention_dim=None => BasicTransformerBlock never reads encoder_hidden_states), so it is not evaluated.
    Returns 5 feature maps [B, C, F, h, w]."""
    B, _, Fr, H, W = x.shape
    x = x.permute(0, 2, 1, 3, 4).reshape(B * Fr, 3, H, W)
    spec = [(0, 3, 1, 1), (3, 4, 2, 1), (6, 3, 1, 1), (9, 4, 2, 1), (12, 3, 1, 1), (15, 4, 2, 1), (18, 3, 1, 1),
            (21, 3, 1, 1)]
    for idx, k, s, pad in spec:
        x = conv2d(sd, f"conv_layers.{idx}", x, stride=s, padding=pad)
        x = F.relu(_bn_train(sd, f"conv_layers.{idx + 1}", x))
    x = conv2d(sd, "final_proj", x, padding=0) * sd["scale"]
    fea = [x]

    def stage(name, x, strides):
        for n, s in enumerate(strides):
            x = conv2d(sd, f"{name}.{3 * n}", x, stride=s)
            x = F.relu(_bn_train(sd, f"{name}.{3 * n + 1}", x))
        return x

    for k, strides in enumerate([(1, 2), (1, 2), (1, 2), (1,)], start=1):
        x = stage(f"conv_layers_{k}", x, strides)
        if f"cross_attn{k}.norm.weight" in sd:
            x = spatial_transformer(sd, f"cross_attn{k}", x, None, heads=16)
        fea.append(x)
    return [f.reshape(B, Fr, f.shape[1], f.shape[2], f.shape[3]).permute(0, 2, 1, 3, 4) for f in fea]


# ------------------------------------------------------------------------------------------------ scheduler
class DDIM:
    """diffusers DDIMScheduler [dep 0.24.0] with configs/inference/inference_v2.yaml:24-33."""

    def __init__(self, num_train=1000, beta_start=0.00085, beta_end=0.012):
        betas = torch.linspace(beta_start, beta_end, num_train, dtype=torch.float32)
        abar_sqrt = torch.cumprod(1.0 - betas, 0).sqrt()
        a0, aT = abar_sqrt[0].clone(), abar_sqrt[-1].clone()
        abar_sqrt = (abar_sqrt - aT) * (a0 / (a0 - aT))
        abar = abar_sqrt ** 2
        alphas = torch.cat([abar[0:1], abar[1:] / abar[:-1]])
        betas = 1 - alphas                                   # diffusers stores betas, then alphas = 1 - betas
        self.alphas_cumprod = torch.cumprod(1.0 - betas, 0)
        self.num_train = num_train

    def timesteps(self, n):
        import numpy as np
        return (np.round(np.arange(self.num_train, 0, -self.num_train / n)).astype("int64") - 1).tolist()

    def step(self, v, t, x, n):
        prev = t - self.num_train // n
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else torch.tensor(1.0)
        x0 = a_t.sqrt() * x - (1 - a_t).sqrt() * v
        eps = a_t.sqrt() * v + (1 - a_t).sqrt() * x
        return a_p.sqrt() * x0 + (1 - a_p).sqrt() * eps


def context_windows(num_frames, size=16, overlap=4):
    """`uniform` window scheduler (src/pipelines/context.py:15-42) as the pipeline calls it: step=0, stride=1."""
    if num_frames <= size:
        return [list(range(num_frames))]
    out = []
    for j in range(0, num_frames, size - overlap):
        out.append([e % num_frames for e in range(j, j + size)])
    return out


def denoise_loop(sd_unet, sd_ref, sd_pose, latents, ref_latents, clip_embed, pose_cond, steps, guidance=3.5,
                 context_frames=16, context_overlap=4, c=SD15):
    """Pose2VideoPipeline.__call__ denoising loop (src/pipelines/pipeline_pose2vid_long.py:459-567), CFG on.
    latents [1,4,L,h,w]; ref_latents [1,4,h,w]; clip_embed [1,768]; pose_cond [1,3,L,H,W] (already preprocessed).
    With context_frames >= L this is also the single-window loop of src/pipelines/pipeline_pose2vid.py:396-437 (one
    PoseGuider pass, one UNet call per step over all frames). `c`: UNet config (block widths) for reduced-size tests."""
    ehs = torch.cat([torch.zeros_like(clip_embed), clip_embed], 0).unsqueeze(1)
    banks = reference_unet_banks(sd_ref, ref_latents.repeat(2, 1, 1, 1), ehs, c=c)
    sched = DDIM()
    L = latents.shape[2]
    windows = context_windows(L, context_frames, context_overlap)
    for t in sched.timesteps(steps):
        noise = torch.zeros(2, *latents.shape[1:], dtype=latents.dtype)
        counter = torch.zeros(1, 1, L, 1, 1, dtype=latents.dtype)
        for wdw in windows:
            lat = latents[:, :, wdw].repeat(2, 1, 1, 1, 1)
            pose_fea = pose_guider_forward(sd_pose, pose_cond[:, :, wdw].repeat(2, 1, 1, 1, 1),
                                           c0=c["block_out_channels"][0])
            pred = unet3d_forward(sd_unet, lat, t, ehs, pose_fea, banks, cfg=True, c=c)
            noise[:, :, wdw] += pred
            counter[:, :, wdw] += 1
        u, cnd = (noise / counter).chunk(2)
        latents = sched.step(u + guidance * (cnd - u), t, latents, steps)
    return latents


# ------------------------------------------------------------------------------------------------ VAE decode
def vae_decode(sd, z, groups=32):
    """AutoencoderKL.decode [dep 0.24.0] for sd-vae-ft-mse as called by decode_latents
    (src/pipelines/pipeline_pose2vid_long.py:113-126; the caller divides by 0.18215 first). z: [n,4,h,w] -> [n,3,8h,8w]."""
    def res(p, x):
        return resnet_block(sd, p, x, None, groups, 1e-6)

    x = conv2d(sd, "post_quant_conv", z, padding=0)
    x = conv2d(sd, "decoder.conv_in", x)
    x = res("decoder.mid_block.resnets.0", x)
    p = "decoder.mid_block.attentions.0"
    b, c, h, w = x.shape
    t = group_norm(sd, p + ".group_norm", x, groups, 1e-6).view(b, c, h * w).transpose(1, 2)
    q, k, v = linear(sd, p + ".to_q", t), linear(sd, p + ".to_k", t), linear(sd, p + ".to_v", t)
    a = ((q @ k.transpose(1, 2)) * c ** -0.5).softmax(-1) @ v
    a = linear(sd, p + ".to_out.0", a).transpose(1, 2).reshape(b, c, h, w)
    x = x + a
    x = res("decoder.mid_block.resnets.1", x)
    for i in range(4):
        for j in range(3):
            x = res(f"decoder.up_blocks.{i}.resnets.{j}", x)
        if i < 3:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = conv2d(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", x)
    x = F.silu(group_norm(sd, "decoder.conv_norm_out", x, groups, 1e-6))
    return conv2d(sd, "decoder.conv_out", x)


def vae_encode(sd, x, groups=32):
    """AutoencoderKL.encode [dep 0.24.0] -> moments (mean | logvar) as the pipeline uses it for the reference image
    (src/pipelines/pipeline_pose2vid_long.py:430-431 takes latent_dist.mean * 0.18215). x: [n,3,H,W] -> [n,8,H/8,W/8].
    Downsample2D(padding=0): pad right/bottom by one, 3x3 conv stride 2."""
    def res(p, x):
        return resnet_block(sd, p, x, None, groups, 1e-6)

    x = conv2d(sd, "encoder.conv_in", x)
    n_blocks = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.down_blocks."))
    for i in range(n_blocks):
        j = 0
        while f"encoder.down_blocks.{i}.resnets.{j}.norm1.weight" in sd:
            x = res(f"encoder.down_blocks.{i}.resnets.{j}", x)
            j += 1
        if f"encoder.down_blocks.{i}.downsamplers.0.conv.weight" in sd:
            x = conv2d(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", F.pad(x, (0, 1, 0, 1)), stride=2, padding=0)
    x = res("encoder.mid_block.resnets.0", x)
    p = "encoder.mid_block.attentions.0"
    b, c, h, w = x.shape
    t = group_norm(sd, p + ".group_norm", x, groups, 1e-6).view(b, c, h * w).transpose(1, 2)
    q, k, v = linear(sd, p + ".to_q", t), linear(sd, p + ".to_k", t), linear(sd, p + ".to_v", t)
    a = ((q @ k.transpose(1, 2)) * c ** -0.5).softmax(-1) @ v
    x = x + linear(sd, p + ".to_out.0", a).transpose(1, 2).reshape(b, c, h, w)
    x = res("encoder.mid_block.resnets.1", x)
    x = F.silu(group_norm(sd, "encoder.conv_norm_out", x, groups, 1e-6))
    return conv2d(sd, "quant_conv", conv2d(sd, "encoder.conv_out", x), padding=0)

