"""Building blocks shared by the denoising UNet3D and the ReferenceNet UNet2D mirrors.

Every class keeps the reference's attribute tree so that `state_dict()` keys (and the published checkpoints:
denoising_unet.pth, motion_module.pth, reference_unet.pth) are interchangeable with the reference's
src/models/{resnet,transformer_3d,attention,motion_module,unet_3d_blocks,unet_2d_blocks}.py. Their forward passes,
however, are sequences of the sm_100a kernels in aniportrait_b200/csrc operating on channels-last fp16 activations
`[frames, H, W, C]` (== token matrices `[frames*H*W, C]`, so conv <-> attention hand-offs are zero-copy). There is no
torch-op fallback: running a block on a non-CUDA tensor raises.
"""
from __future__ import annotations

import math

import torch
from torch import nn

import os

from .. import ops
from .modeling import PackedCache, f16, f32

# Normalisation fusion: LayerNorms are folded into the consuming GEMM (row statistics from the producer's
# epilogue), GroupNorm statistics come from the producing conv / GEMM epilogue. AP_FUSE_NORMS=0 restores the standalone
# LayerNorm / GroupNorm-statistics kernels (A/B timing, and the path the ReferenceNet write pass still needs for norm1).
FUSE_NORMS = os.environ.get("AP_FUSE_NORMS", "1") != "0"


def _cs(t):
    """ColStats attached to an activation by the op that produced it (None: the GroupNorm runs its own statistics pass)."""
    return getattr(t, "_ap_cs", None)


def _tag(t, cs):
    if cs is not None:
        t._ap_cs = cs
    return t


def fold_layer_norm(w: torch.Tensor, b, gamma: torch.Tensor, beta: torch.Tensor):
    """LN(x) W^T + b = rstd (x (W diag(gamma))^T - mean colsum) + (W beta + b). Returns (W'' fp16 [N, K + 8], bias' fp32):
    W'' = [W diag(gamma) | cs_hi, cs_hi, cs_lo, 0 x 5] where cs = colsum of the ROUNDED W diag(gamma) split into two fp16
    halves; multiplied by the activation operand [x | -mean_hi, -mean_lo, -mean_hi, 0 x 5] (ops.LNFold) the accumulator is
    x W'^T - mean cs up to ~2^-22 relative (the lo x lo cross term is dropped). w: [N, K] any dtype; b: [N] or None."""
    w32 = w.detach().to(torch.float32)
    wg = (w32 * gamma.detach().to(torch.float32)[None, :]).to(torch.float16)
    cs = wg.to(torch.float32).sum(dim=1)
    cs_hi = cs.to(torch.float16)
    cs_lo = (cs - cs_hi.to(torch.float32)).to(torch.float16)
    ext = torch.zeros(w.shape[0], ops.LN_EXTRA_K, dtype=torch.float16, device=w.device)
    ext[:, 0], ext[:, 1], ext[:, 2] = cs_hi, cs_hi, cs_lo
    bias = w32 @ beta.detach().to(torch.float32)
    if b is not None:
        bias = bias + b.detach().to(torch.float32)
    return torch.cat([wg, ext], dim=1).contiguous(), bias.contiguous()


class RunCtx:
    """Per-forward state threaded through the blocks."""

    def __init__(self, batch: int, frames: int, temb_act: torch.Tensor | None, ehs: torch.Tensor | None,
                 ehs_key=None, ref_branch=None, emit_stats: bool = True, temb_bias=None, group=None):
        # ref_branch: None = the reference's layout (both CFG branches in the batch when the reader was built with
        # do_classifier_free_guidance); "uncond" / "cond" = this call carries ONE branch of a CFG reader (multi-GPU
        # (window, branch) work units): uncond never reads the bank, cond reads the conditional bank for every frame
        self.ref_branch = ref_branch
        # group = (n_uncond, n_cond): the batch elements are WINDOWS OF ONE VIDEO, the first n_uncond unconditional, the rest
        # conditional (sharded mode: all units of a rank in one call -> larger GEMMs, fewer wave-quantisation losses than one
        # batch-1 call per unit). Unconditional elements never read the bank, every conditional element reads the video's
        # conditional bank.
        self.group = group
        self.B = batch              # CFG branches x videos
        self.F = frames             # frames per batch element (1 for the 2-D ReferenceNet)
        self.temb_act = temb_act    # SiLU(time embedding) [B, 1280] fp16
        self.ehs = ehs              # encoder hidden states [B, S, 768] fp16
        # identity of `ehs` for the per-block attn2-constant cache: an explicit token handed out per video by the caller
        # (pipeline session). None = no caching (the constant is recomputed on every call): tensor addresses / versions
        # are NOT an identity — a fresh clone of another video's embedding routinely lands on the same address.
        self.ehs_key = ehs_key
        # emit_stats: outputs of transformer / motion blocks feed a GroupNorm (UNets: yes; PoseGuider: no)
        self.emit_stats = emit_stats and FUSE_NORMS
        # temb_bias: {id(resnet): fp32 [B, cout] view}: every resnet's time-embedding bias from ONE GEMM per call
        self.temb_bias = temb_bias


# ------------------------------------------------------------------------------------------------------------
# parameter holders that mirror diffusers' module tree
# ------------------------------------------------------------------------------------------------------------
class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    """diffusers FeedForward(activation_fn='geglu'): net.0.proj, net.2 (reference src/models/attention.py:361)."""

    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])
        self._pk = PackedCache()

    def packed(self):
        def build():
            w1, b1 = ops.interleave_geglu(f16(self.net[0].proj.weight), f32(self.net[0].proj.bias))
            return dict(w1=w1, b1=b1, w2=f16(self.net[2].weight), b2=f32(self.net[2].bias))
        return self._pk.get(self, build)

    def run(self, x_norm: torch.Tensor, residual: torch.Tensor, **kw) -> torch.Tensor:
        pk = self.packed()
        h = ops.gemm(x_norm, pk["w1"], bias=pk["b1"], geglu=True)
        return ops.gemm(h, pk["w2"], bias=pk["b2"], residual=residual, **kw)

    def folded(self, gamma: torch.Tensor, beta: torch.Tensor):
        """GEGLU projection with the preceding LayerNorm (gamma, beta) folded in: (w1'' [8C, C + 8], b1'), value / gate rows
        interleaved like packed()['w1']."""
        wg, bias = fold_layer_norm(self.net[0].proj.weight, self.net[0].proj.bias, gamma, beta)
        return ops.interleave_geglu(wg, bias)

    def run_folded(self, x: torch.Tensor, stats, fold, residual: torch.Tensor, eps: float = 1e-5, **kw) -> torch.Tensor:
        """x is the UN-normalised input; `stats` its RowStats from the producer's epilogue; fold = self.folded(...)."""
        pk = self.packed()
        w1, b1 = fold
        h = ops.gemm(x, w1, bias=b1, geglu=True, ln=ops.LNFold(stats, eps))
        return ops.gemm(h, pk["w2"], bias=pk["b2"], residual=residual, **kw)


class Attention(nn.Module):
    """diffusers Attention parameters: to_q, to_k, to_v (no bias), to_out.0 (bias)."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64):
        super().__init__()
        inner = heads * dim_head
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.dim_head = dim_head
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv_dim, inner, bias=False)
        self.to_v = nn.Linear(kv_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])


# ------------------------------------------------------------------------------------------------------------
# ResnetBlock (3-D inflated == per-frame 2-D)
# ------------------------------------------------------------------------------------------------------------
class ResnetBlock(nn.Module):
    """ResnetBlock3D / diffusers ResnetBlock2D (reference src/models/resnet.py:124-248):
    GN+SiLU -> conv3x3 (+ time-embedding bias) -> GN+SiLU -> conv3x3 (+ residual | 1x1 shortcut)."""

    def __init__(self, in_channels, out_channels, temb_channels=1280, groups=32, eps=1e-5):
        super().__init__()
        self.in_channels, self.out_channels, self.groups, self.eps = in_channels, out_channels, groups, eps
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None
        self._pk = PackedCache()

    def packed(self):
        def build():
            d = dict(g1=f32(self.norm1.weight), b1=f32(self.norm1.bias), g2=f32(self.norm2.weight),
                     b2=f32(self.norm2.bias), w1=ops.pack_conv3x3_weight(self.conv1.weight.detach()),
                     w2=ops.pack_conv3x3_weight(self.conv2.weight.detach()), cb2=f32(self.conv2.bias))
            if self.time_emb_proj is not None:
                d["wt"] = f16(self.time_emb_proj.weight)
                d["bt"] = f32(self.time_emb_proj.bias) + f32(self.conv1.bias)   # conv1 bias folded into the temb bias
            else:
                d["cb1"] = f32(self.conv1.bias)
            if self.conv_shortcut is not None:
                d["ws"] = f16(self.conv_shortcut.weight.reshape(self.out_channels, self.in_channels))
                d["bs"] = f32(self.conv_shortcut.bias)
            return d
        return self._pk.get(self, build)

    def run(self, x: torch.Tensor, ctx: RunCtx, skip: torch.Tensor | None = None) -> torch.Tensor:
        """x: [Nf,H,W,C1]; skip: optional [Nf,H,W,C2] (the reference's torch.cat([x, skip], dim=1))."""
        pk = self.packed()
        nf, h, w, _ = x.shape
        cout = self.out_channels
        # GroupNorm statistics come from the epilogue of whatever produced x / skip / hc (ColStats tags); a source without a
        # tag (conv_in + pose, x + pose feature) makes that GroupNorm run its own statistics pass
        emit = FUSE_NORMS and ops.conv_col_stats_ok(nf, h, w) and cout % 32 == 0
        hn = ops.group_norm(x, pk["g1"], pk["b1"], self.groups, self.eps, True, x2=skip, stats=_cs(x),
                            stats2=_cs(skip) if skip is not None else None)
        if self.time_emb_proj is not None:
            tb = ctx.temb_bias.get(id(self)) if ctx.temb_bias is not None else None
            bias1 = tb if tb is not None else ops.gemm(ctx.temb_act, pk["wt"], bias=pk["bt"], out_f32=True)   # [B, cout]
            hc = ops.conv3x3(hn, pk["w1"], cout, bias=bias1, bias_group_rows=ctx.F * h * w, col_stats=emit)
        else:
            hc = ops.conv3x3(hn, pk["w1"], cout, bias=pk["cb1"], col_stats=emit)
        hc, cs_hc = hc if emit else (hc, None)
        hn2 = ops.group_norm(hc, pk["g2"], pk["b2"], self.groups, self.eps, True, stats=cs_hc)
        if self.conv_shortcut is not None:
            c1 = x.shape[-1]
            a = x.view(-1, c1)
            a2 = skip.view(-1, skip.shape[-1]) if skip is not None else None
            res = ops.gemm(a, pk["ws"], bias=pk["bs"], a2=a2).view(nf, h, w, cout)
        else:
            assert skip is None
            res = x
        out = ops.conv3x3(hn2, pk["w2"], cout, bias=pk["cb2"], residual=res, col_stats=emit)
        return _tag(*out) if emit else out


class Downsample(nn.Module):
    """Downsample3D / Downsample2D(name='op'): conv3x3 stride 2 padding 1 (reference src/models/resnet.py:94-121)."""

    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)
        self._pk = PackedCache()

    def run(self, x):
        pk = self._pk.get(self, lambda: dict(w=ops.pack_conv3x3_weight(self.conv.weight.detach()), b=f32(self.conv.bias)))
        nf, h, w, _ = x.shape
        emit = FUSE_NORMS and ops.conv_col_stats_ok(nf, h // 2, w // 2) and self.conv.out_channels % 32 == 0
        out = ops.conv3x3(x, pk["w"], self.conv.out_channels, bias=pk["b"], stride=2, col_stats=emit)
        return _tag(*out) if emit else out


class Upsample(nn.Module):
    """Upsample3D / Upsample2D: nearest 2x + conv3x3 (reference src/models/resnet.py:32-91)."""

    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)
        self._pk = PackedCache()

    def run(self, x):
        pk = self._pk.get(self, lambda: dict(w=ops.pack_conv3x3_weight(self.conv.weight.detach()), b=f32(self.conv.bias)))
        nf, h, w, _ = x.shape
        emit = FUSE_NORMS and ops.conv_col_stats_ok(nf, 2 * h, 2 * w) and self.conv.out_channels % 32 == 0
        out = ops.conv3x3(ops.upsample2x(x), pk["w"], self.conv.out_channels, bias=pk["b"], col_stats=emit)
        return _tag(*out) if emit else out


# ------------------------------------------------------------------------------------------------------------
# spatial transformer (self / reference attention + constant cross-attention + GEGLU FF)
# ------------------------------------------------------------------------------------------------------------
class BasicTransformerBlock(nn.Module):
    """Parameters of (Temporal)BasicTransformerBlock (reference src/models/attention.py:14-445) and the forward the
    reference installs on it through ReferenceAttentionControl (src/models/mutual_self_attention.py:93-265).

    `bank` / `_ref_mode` / `_ref_cfg` are managed by aniportrait_b200.models.mutual_self_attention."""

    def __init__(self, dim, heads, dim_head, cross_attention_dim=768):
        super().__init__()
        self.dim, self.heads, self.dim_head = dim, heads, dim_head
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        if cross_attention_dim is not None:
            self.norm2 = nn.LayerNorm(dim)
            self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        else:
            self.norm2 = None
            self.attn2 = None
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)
        self.bank = []
        self._ref_mode = None       # None | "write" | "read"
        self._ref_cfg = False
        self._pk = PackedCache()
        self._bank_kv = None        # (key, [n_banks*N, 2*heads*dpad] fp16)
        self._attn2_const = None    # (key, fp32 [B, dim])

    def packed(self):
        def build():
            dpad = ops.head_pad(self.dim_head)
            wq = ops.pad_head_rows(f16(self.attn1.to_q.weight), self.heads, dpad)
            wk = ops.pad_head_rows(f16(self.attn1.to_k.weight), self.heads, dpad)
            wv = ops.pad_head_rows(f16(self.attn1.to_v.weight), self.heads, dpad)
            d = dict(dpad=dpad, wqkv=torch.cat([wq, wk, wv], 0).contiguous(), wkv=torch.cat([wk, wv], 0).contiguous(),
                     wo=f16(self.attn1.to_out[0].weight), bo=f32(self.attn1.to_out[0].bias),
                     g1=f32(self.norm1.weight), b1=f32(self.norm1.bias), g3=f32(self.norm3.weight),
                     b3=f32(self.norm3.bias))
            if self.attn2 is not None:
                d.update(wv2=f16(self.attn2.to_v.weight), wo2=f16(self.attn2.to_out[0].weight),
                         bo2o=(f32(self.attn2.to_out[0].bias) + f32(self.attn1.to_out[0].bias)).contiguous())
            # norm1 folded into the q/k/v projection, norm3 into the GEGLU projection (fold_layer_norm); the padded head
            # rows stay zero (zero weight rows -> zero colsum and bias)
            qkv_w = torch.cat([self.attn1.to_q.weight, self.attn1.to_k.weight, self.attn1.to_v.weight], 0)
            wg, bias = fold_layer_norm(qkv_w, None, self.norm1.weight, self.norm1.bias)
            inner = self.heads * self.dim_head
            d["wqkv_g"] = torch.cat([ops.pad_head_rows(wg[i * inner:(i + 1) * inner], self.heads, dpad) for i in range(3)],
                                    0).contiguous()
            pad1 = lambda v: ops.pad_head_rows(v[:, None], self.heads, dpad).reshape(-1)    # noqa: E731
            d["b_qkv"] = torch.cat([pad1(bias[i * inner:(i + 1) * inner]) for i in range(3)]).contiguous()
            d["ff_fold"] = self.ff.folded(self.norm3.weight, self.norm3.bias)
            d["eps1"], d["eps3"] = self.norm1.eps, self.norm3.eps
            return d
        return self._pk.get(self, build)

    def _attn_out_bias(self, pk, ctx: RunCtx):
        """attn1 out bias (+ the attn2 term). With a single encoder token softmax == 1, so attn2's output is the
        per-batch constant to_out(to_v(e)) + b, independent of the query (reference attention.py:339-347 runs the full
        q / out GEMMs for it every step). Folded into the to_out epilogue bias of attn1."""
        if self.attn2 is None:
            return pk["bo"], 0
        ehs = ctx.ehs
        if ehs is None or ehs.shape[1] != 1:
            raise NotImplementedError("attn2 with more than one encoder token is not part of the AniPortrait hot path")
        def compute():
            e = ehs.reshape(ehs.shape[0], -1).contiguous()
            v = ops.gemm(e, pk["wv2"])
            return ops.gemm(v, pk["wo2"], bias=pk["bo2o"], out_f32=True)
        if ctx.ehs_key is None:
            return compute(), 1
        # (video, CFG branch layout, batch, packed weights): a video needs at most 3 entries (both branches together,
        # uncond alone, cond alone)
        key = (ctx.ehs_key, ctx.ref_branch, ctx.group, ehs.shape[0], id(pk))
        cache = self._attn2_const if isinstance(self._attn2_const, dict) else {}
        if key not in cache:
            if len(cache) >= 4:
                cache.clear()
            cache[key] = compute()
            self._attn2_const = cache
        return cache[key], 1

    def _bank_projection(self, pk, n_tokens):
        """K/V of the ReferenceNet bank under THIS block's to_k/to_v. The bank is step- and frame-invariant, so it is
        projected once per video (the reference re-projects it for 16 frames x 25 steps)."""
        bank = self.bank[0]
        key = (bank.data_ptr(), bank._version, id(pk))
        if self._bank_kv is None or self._bank_kv[0] != key:
            nb = bank.shape[0]
            kv = ops.gemm(bank.reshape(nb * bank.shape[1], bank.shape[2]).contiguous(), pk["wkv"])
            self._bank_kv = (key, kv, nb)
        return self._bank_kv[1], self._bank_kv[2]

    def prepare(self, ctx: RunCtx):
        """Per-video precompute (outside any CUDA-graph capture): attn2 constant and the bank's K/V projection."""
        pk = self.packed()
        self._attn_out_bias(pk, ctx)
        if self._ref_mode == "read" and len(self.bank) > 0:
            self._bank_projection(pk, self.bank[0].shape[1])

    def fused(self) -> bool:
        """LayerNorm folding is used unless the block has to materialise norm1(x) for the bank (ReferenceNet write pass)."""
        return FUSE_NORMS and self._ref_mode != "write"

    def run(self, t0: torch.Tensor, n_frames: int, tokens: int, ctx: RunCtx, t0_stats=None, out_kw=None) -> torch.Tensor:
        """t0: [n_frames*tokens, dim] fp16. t0_stats: RowStats of t0 from its producer's epilogue (required when fused()).
        out_kw: extra ops.gemm arguments for the block's LAST GEMM (the caller may ask for statistics of the output)."""
        pk = self.packed()
        dpad, heads, d = pk["dpad"], self.heads, self.dim_head
        hp = heads * dpad
        fused = self.fused() and t0_stats is not None
        if fused:
            qkv = ops.gemm(t0, pk["wqkv_g"], bias=pk["b_qkv"], ln=ops.LNFold(t0_stats, pk["eps1"]))
        else:
            n1 = ops.layer_norm(t0, pk["g1"], pk["b1"], pk["eps1"])
            if self._ref_mode == "write":
                self.bank.append(n1.view(n_frames, tokens, self.dim).clone())
            qkv = ops.gemm(n1, pk["wqkv"])
        q, k, v = qkv[:, :hp], qkv[:, hp:2 * hp], qkv[:, 2 * hp:]
        kw = {}
        if self._ref_mode == "read" and len(self.bank) > 0 and ctx.ref_branch != "uncond" and \
                not (ctx.group is not None and ctx.group[1] == 0):
            kv, nb = self._bank_projection(pk, tokens)
            bank_tokens = self.bank[0].shape[1]
            if self._ref_cfg and ctx.group is not None:
                # windows of one video: elements [0, n_uncond) skip the bank, all the others read the conditional bank
                n_uncond, n_cond = ctx.group
                nbh = nb // 2
                kw = dict(bank_k=kv[nbh * bank_tokens:, :hp], bank_v=kv[nbh * bank_tokens:, hp:],
                          bank_tokens=bank_tokens, n_banks=nb - nbh, first_bank_frame=n_uncond * ctx.F,
                          frames_per_bank=n_cond * ctx.F)
            elif self._ref_cfg and ctx.ref_branch == "cond":
                # single-branch call of a CFG reader: every frame reads the bank of the conditional ReferenceNet pass
                nbh = nb // 2
                kw = dict(bank_k=kv[nbh * bank_tokens:, :hp], bank_v=kv[nbh * bank_tokens:, hp:],
                          bank_tokens=bank_tokens, n_banks=nb - nbh, first_bank_frame=0, frames_per_bank=ctx.F)
            elif self._ref_cfg:
                # frames of the first half of the batch are the unconditional branch: no reference keys
                # (mutual_self_attention.py:166-186); frame f of the second half reads bank[f // F] (:148-157)
                first = (ctx.B // 2) * ctx.F
                kw = dict(bank_k=kv[:, :hp], bank_v=kv[:, hp:], bank_tokens=bank_tokens, n_banks=nb,
                          first_bank_frame=first, frames_per_bank=ctx.F)
                if nb > 1:   # bank rows of the conditional half only
                    nbh = nb // 2
                    kw["bank_k"] = kv[nbh * bank_tokens:, :hp]
                    kw["bank_v"] = kv[nbh * bank_tokens:, hp:]
                    kw["n_banks"] = nb - nbh
            else:
                kw = dict(bank_k=kv[:, :hp], bank_v=kv[:, hp:], bank_tokens=bank_tokens, n_banks=nb,
                          first_bank_frame=0, frames_per_bank=ctx.F)
        a = ops.attention(q, k, v, n_frames, tokens, heads, d, dpad, **kw)
        bias, grouped = self._attn_out_bias(pk, ctx)
        gr = ctx.F * tokens if grouped else 0
        out_kw = out_kw or {}
        if fused:
            t1, st1 = ops.gemm(a, pk["wo"], bias=bias, residual=t0, bias_group_rows=gr, row_stats=True)
            return self.ff.run_folded(t1, st1, pk["ff_fold"], t1, pk["eps3"], **out_kw)
        t1 = ops.gemm(a, pk["wo"], bias=bias, residual=t0, bias_group_rows=gr)
        n3 = ops.layer_norm(t1, pk["g3"], pk["b3"], pk["eps3"])
        return self.ff.run(n3, t1, **out_kw)


class TemporalBasicTransformerBlock(BasicTransformerBlock):
    """Same parameters/forward as BasicTransformerBlock; separate class name as in the reference
    (src/models/attention.py:300) so reader/writer selection by type keeps working."""


class SpatialTransformer(nn.Module):
    """Transformer3DModel / Transformer2DModel with use_linear_projection=False (reference
    src/models/transformer_3d.py:27-169): GroupNorm(eps 1e-6) -> 1x1 conv -> block -> 1x1 conv -> + residual."""

    block_cls = BasicTransformerBlock

    def __init__(self, heads, dim_head, in_channels, cross_attention_dim=768, groups=32):
        super().__init__()
        inner = heads * dim_head
        self.in_channels, self.inner, self.groups = in_channels, inner, groups
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([self.block_cls(inner, heads, dim_head, cross_attention_dim)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)
        self._pk = PackedCache()

    def packed(self):
        def build():
            return dict(g=f32(self.norm.weight), b=f32(self.norm.bias),
                        wi=f16(self.proj_in.weight.reshape(self.inner, self.in_channels)), bi=f32(self.proj_in.bias),
                        wo=f16(self.proj_out.weight.reshape(self.in_channels, self.inner)), bo=f32(self.proj_out.bias))
        return self._pk.get(self, build)

    def run(self, x: torch.Tensor, ctx: RunCtx) -> torch.Tensor:
        pk = self.packed()
        nf, h, w, c = x.shape
        blk = self.transformer_blocks[0]
        hn = ops.group_norm(x, pk["g"], pk["b"], self.groups, 1e-6, False, stats=_cs(x))
        if blk.fused():
            t0, st0 = ops.gemm(hn.view(-1, c), pk["wi"], bias=pk["bi"], row_stats=True)
        else:
            t0, st0 = ops.gemm(hn.view(-1, c), pk["wi"], bias=pk["bi"]), None
        t = blk.run(t0, nf, h * w, ctx, t0_stats=st0)
        emit = ctx.emit_stats and (h * w) % 32 == 0 and c % 32 == 0
        out = ops.gemm(t, pk["wo"], bias=pk["bo"], residual=x.view(-1, c), col_stats=emit)
        if emit:
            return _tag(out[0].view(nf, h, w, c), out[1])
        return out.view(nf, h, w, c)


class Transformer3DModel(SpatialTransformer):
    block_cls = TemporalBasicTransformerBlock


class Transformer2DModel(SpatialTransformer):
    block_cls = BasicTransformerBlock


# ------------------------------------------------------------------------------------------------------------
# motion module
# ------------------------------------------------------------------------------------------------------------
class PositionalEncoding(nn.Module):
    """Sinusoidal buffer `pe` [1, max_len, d] (reference src/models/motion_module.py:262-277)."""

    def __init__(self, d_model, max_len=32):
        super().__init__()
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(1, max_len, d_model)
        pe[0, :, 0::2] = torch.sin(position * div_term)
        pe[0, :, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)


class VersatileAttention(Attention):
    def __init__(self, query_dim, heads, dim_head, max_len):
        super().__init__(query_dim, None, heads, dim_head)
        self.pos_encoder = PositionalEncoding(query_dim, max_len)


class TemporalTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, max_len):
        super().__init__()
        self.attention_blocks = nn.ModuleList([VersatileAttention(dim, heads, dim_head, max_len) for _ in range(2)])
        self.norms = nn.ModuleList([nn.LayerNorm(dim) for _ in range(2)])
        self.ff = FeedForward(dim)
        self.ff_norm = nn.LayerNorm(dim)


class TemporalTransformer3DModel(nn.Module):
    """TemporalTransformer3DModel (reference src/models/motion_module.py:94-182) with one TemporalTransformerBlock
    (two Temporal_Self attentions + GEGLU FF)."""

    def __init__(self, in_channels, heads=8, max_len=32, groups=32):
        super().__init__()
        self.channels, self.heads, self.groups, self.max_len = in_channels, heads, groups, max_len
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, in_channels)
        self.transformer_blocks = nn.ModuleList([TemporalTransformerBlock(in_channels, heads, in_channels // heads,
                                                                          max_len)])
        self.proj_out = nn.Linear(in_channels, in_channels)
        self._pk = PackedCache()

    def packed(self):
        def build():
            blk = self.transformer_blocks[0]
            d = dict(g=f32(self.norm.weight), b=f32(self.norm.bias), wi=f16(self.proj_in.weight),
                     bi=f32(self.proj_in.bias), wo=f16(self.proj_out.weight), bo=f32(self.proj_out.bias),
                     gf=f32(blk.ff_norm.weight), bf=f32(blk.ff_norm.bias), attn=[])
            for i in range(2):
                at = blk.attention_blocks[i]
                d["attn"].append(dict(
                    wqkv=torch.cat([f16(at.to_q.weight), f16(at.to_k.weight), f16(at.to_v.weight)], 0).contiguous(),
                    wo=f16(at.to_out[0].weight), bo=f32(at.to_out[0].bias), g=f32(blk.norms[i].weight),
                    b=f32(blk.norms[i].bias),
                    # the reference adds the (fp16) buffer to the LayerNorm output: keep its rounding
                    pe=at.pos_encoder.pe[0].detach().to(torch.float16).to(torch.float32).contiguous()))
                # norms[i] folded into the q/k/v projection; beta and the positional encoding become a per-frame bias table
                qkv_w = torch.cat([at.to_q.weight, at.to_k.weight, at.to_v.weight], 0)
                wg, _ = fold_layer_norm(qkv_w, None, blk.norms[i].weight, blk.norms[i].bias)
                d["attn"][i].update(wqkv_g=wg, eps=blk.norms[i].eps)
            d["ff_fold"] = blk.ff.folded(blk.ff_norm.weight, blk.ff_norm.bias)
            d["eps_f"] = blk.ff_norm.eps
            d["pe_bias"] = {}
            return d
        return self._pk.get(self, build)

    @staticmethod
    def _pe_bias(pk, i, B, F):
        """Bias table of the folded q/k/v projection: row b * F + f = (beta + pe[f]) . Wqkv^T  (LN(m) + pe[f] is what the
        reference projects, motion_module.py:365-366). Built once per (batch, window length)."""
        key = (i, B, F)
        tab = pk["pe_bias"].get(key)
        if tab is None:
            a = pk["attn"][i]
            t = (a["b"][None, :] + a["pe"][:F]) @ a["wqkv"].to(torch.float32).t()         # [F, 3C]
            tab = t.repeat(B, 1).contiguous()
            pk["pe_bias"][key] = tab
        return tab

    def prepare(self, batch: int, frames: int):
        """Per-geometry precompute outside any CUDA-graph capture (torch math: a 32 x C by C x 3C matmul per attention)."""
        if FUSE_NORMS:
            pk = self.packed()
            for i in range(2):
                self._pe_bias(pk, i, batch, frames)

    def run(self, x: torch.Tensor, ctx: RunCtx) -> torch.Tensor:
        pk = self.packed()
        nf, h, w, c = x.shape
        n_tok = h * w
        if ctx.F > self.max_len:
            raise ValueError(f"window of {ctx.F} frames exceeds temporal_position_encoding_max_len={self.max_len}")
        hn = ops.group_norm(x, pk["g"], pk["b"], self.groups, 1e-6, False, stats=_cs(x))
        emit = ctx.emit_stats and n_tok % 32 == 0 and c % 32 == 0
        if FUSE_NORMS:
            m, st = ops.gemm(hn.view(-1, c), pk["wi"], bias=pk["bi"], row_stats=True)
            for i in range(2):
                a = pk["attn"][i]
                qkv = ops.gemm(m, a["wqkv_g"], bias=self._pe_bias(pk, i, ctx.B, ctx.F), bias_group_rows=n_tok,
                               ln=ops.LNFold(st, a["eps"]))
                o = ops.temporal_attention(qkv, ctx.B, ctx.F, n_tok, c, self.heads)
                m, st = ops.gemm(o, a["wo"], bias=a["bo"], residual=m, row_stats=True)
            m = self.transformer_blocks[0].ff.run_folded(m, st, pk["ff_fold"], m, pk["eps_f"])
        else:
            m = ops.gemm(hn.view(-1, c), pk["wi"], bias=pk["bi"])
            for i in range(2):
                a = pk["attn"][i]
                n = ops.layer_norm(m, a["g"], a["b"], a["eps"], pe=a["pe"], rows_per_pe=n_tok, pe_period=ctx.F)
                qkv = ops.gemm(n, a["wqkv"])
                o = ops.temporal_attention(qkv, ctx.B, ctx.F, n_tok, c, self.heads)
                m = ops.gemm(o, a["wo"], bias=a["bo"], residual=m)
            n = ops.layer_norm(m, pk["gf"], pk["bf"], pk["eps_f"])
            m = self.transformer_blocks[0].ff.run(n, m)
        out = ops.gemm(m, pk["wo"], bias=pk["bo"], residual=x.view(-1, c), col_stats=emit)
        if emit:
            return _tag(out[0].view(nf, h, w, c), out[1])
        return out.view(nf, h, w, c)


class VanillaTemporalModule(nn.Module):
    """motion_modules.N = VanillaTemporalModule(temporal_transformer=...) (reference motion_module.py:44-91)."""

    def __init__(self, in_channels, num_attention_heads=8, temporal_position_encoding_max_len=32, **_):
        super().__init__()
        self.temporal_transformer = TemporalTransformer3DModel(in_channels, num_attention_heads,
                                                               temporal_position_encoding_max_len)

    def run(self, x, ctx):
        return self.temporal_transformer.run(x, ctx)
