"""GPU parity of the norm / attention / temporal / elementwise kernels against plain fp32 torch evaluations of the
same ops on the same fp16 inputs. Tolerances (rel-L2): 2e-3 for single-rounding ops, 3e-3 for attention (P is rounded
to fp16 before the P.V tensor-core product, exp2 uses ex2.approx)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _mk(shape, dev, scale=1.0, seed=0, shift=0.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale + shift).to(torch.float16).to(dev)


@pytest.mark.parametrize("nf,hw,c,silu", [(4, 4096, 320, True), (3, 1024, 640, False), (5, 64, 1280, True),
                                          (2, 256, 128, True), (2, 4096, 64, False)])
def test_group_norm(cuda_dev, nf, hw, c, silu):
    from aniportrait_b200 import ops
    x = _mk((nf, hw, c), cuda_dev, 2.0, 1, shift=0.7)
    gamma = torch.randn(c, device=cuda_dev)
    beta = torch.randn(c, device=cuda_dev)
    out = ops.group_norm(x, gamma, beta, 32, 1e-5, silu)
    ref = F.group_norm(x.float().permute(0, 2, 1), 32, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    assert rel_l2(out, ref.permute(0, 2, 1)) < 2e-3


def test_group_norm_two_source(cuda_dev):
    from aniportrait_b200 import ops
    nf, hw, c1, c2 = 3, 1024, 640, 320
    x1 = _mk((nf, hw, c1), cuda_dev, 1.5, 2)
    x2 = _mk((nf, hw, c2), cuda_dev, 0.5, 3, shift=-1.0)
    gamma = torch.randn(c1 + c2, device=cuda_dev)
    beta = torch.randn(c1 + c2, device=cuda_dev)
    out = ops.group_norm(x1, gamma, beta, 32, 1e-5, True, x2=x2)
    ref = F.silu(F.group_norm(torch.cat([x1, x2], -1).float().permute(0, 2, 1), 32, gamma, beta, 1e-5))
    assert out.shape == (nf, hw, c1 + c2)
    assert rel_l2(out, ref.permute(0, 2, 1)) < 2e-3


@pytest.mark.parametrize("rows,c", [(4096, 320), (1000, 640), (77, 1280), (256, 1408), (128, 64)])
def test_layer_norm(cuda_dev, rows, c):
    from aniportrait_b200 import ops
    x = _mk((rows, c), cuda_dev, 3.0, 4, shift=1.0)
    gamma = torch.randn(c, device=cuda_dev)
    beta = torch.randn(c, device=cuda_dev)
    out = ops.layer_norm(x, gamma, beta)
    ref = F.layer_norm(x.float(), (c,), gamma, beta, 1e-5)
    assert rel_l2(out, ref) < 2e-3


def test_layer_norm_with_pe(cuda_dev):
    from aniportrait_b200 import ops
    B, Fr, N, c = 2, 16, 64, 320
    x = _mk((B * Fr * N, c), cuda_dev, 1.0, 5)
    gamma = torch.randn(c, device=cuda_dev)
    beta = torch.randn(c, device=cuda_dev)
    pe = torch.randn(32, c, device=cuda_dev)
    out = ops.layer_norm(x, gamma, beta, pe=pe, rows_per_pe=N, pe_period=Fr)
    ref = F.layer_norm(x.float(), (c,), gamma, beta, 1e-5).view(B, Fr, N, c) + pe[None, :Fr, None]
    assert rel_l2(out, ref.reshape(-1, c)) < 2e-3


def _attn_ref(q, k, v, scale):
    s = (q.float() @ k.float().transpose(-1, -2)) * scale
    return s.softmax(-1) @ v.float()


@pytest.mark.parametrize("tokens,heads,d,frames,bank_from", [
    (1024, 8, 40, 4, 2),     # CFG layout: frames 0-1 uncond (own keys), 2-3 cond (own + bank)
    (256, 8, 80, 4, 2),
    (256, 8, 160, 4, 2),
    (64, 8, 160, 6, 3),      # 8x8 level: fewer tokens than a tile
    (4096, 8, 40, 2, 1),     # full 64x64 level
    (1024, 8, 40, 3, None),  # plain self-attention (ReferenceNet / PoseGuider style)
    (1024, 16, 88, 2, None), # PoseGuider head layout
    (100, 2, 40, 3, 1),      # ragged token count
])
def test_attention(cuda_dev, tokens, heads, d, frames, bank_from):
    from aniportrait_b200 import ops
    dpad = ops.head_pad(d)
    rows = frames * tokens
    g = torch.Generator(device="cpu").manual_seed(7)
    qkv_true = (torch.randn(rows, 3, heads, d, generator=g) * 1.5).to(torch.float16)
    qkv = torch.zeros(rows, 3, heads, dpad, dtype=torch.float16)
    qkv[..., :d] = qkv_true
    qkv = qkv.reshape(rows, 3 * heads * dpad).to(cuda_dev)
    hp = heads * dpad
    q, k, v = qkv[:, :hp], qkv[:, hp:2 * hp], qkv[:, 2 * hp:]
    kwargs = {}
    bank_true = None
    if bank_from is not None:
        bank_true = (torch.randn(tokens, 2, heads, d, generator=g) * 1.5).to(torch.float16)
        bank = torch.zeros(tokens, 2, heads, dpad, dtype=torch.float16)
        bank[..., :d] = bank_true
        bank = bank.reshape(tokens, 2 * hp).to(cuda_dev)
        kwargs = dict(bank_k=bank[:, :hp], bank_v=bank[:, hp:], bank_tokens=tokens, n_banks=1,
                      first_bank_frame=bank_from, frames_per_bank=frames)
    out = ops.attention(q, k, v, frames, tokens, heads, d, dpad, **kwargs)
    torch.cuda.synchronize()
    qt = qkv_true.to(cuda_dev).view(frames, tokens, 3, heads, d).permute(2, 0, 3, 1, 4)  # [3, f, h, n, d]
    refs = []
    for f in range(frames):
        kk, vv = qt[1, f], qt[2, f]
        if bank_from is not None and f >= bank_from:
            bt = bank_true.to(cuda_dev).permute(1, 2, 0, 3)  # [2, h, n, d]
            kk = torch.cat([kk, bt[0]], dim=1)
            vv = torch.cat([vv, bt[1]], dim=1)
        refs.append(_attn_ref(qt[0, f], kk, vv, d ** -0.5))
    ref = torch.stack(refs).permute(0, 2, 1, 3).reshape(rows, heads * d)
    assert out.shape == ref.shape
    err = rel_l2(out, ref)
    assert err < 3e-3, err


def test_attention_large_logits(cuda_dev):
    """Rows whose running max jumps by far more than the lazy-rescale threshold between key tiles."""
    from aniportrait_b200 import ops
    tokens, heads, d, frames = 512, 8, 40, 2
    dpad = ops.head_pad(d)
    g = torch.Generator(device="cpu").manual_seed(11)
    qkv_true = torch.randn(frames * tokens, 3, heads, d, generator=g)
    qkv_true[:, 1] *= torch.linspace(0.2, 6.0, frames * tokens)[:, None, None]  # key norms grow along the sequence
    qkv_true = qkv_true.to(torch.float16)
    qkv = torch.zeros(frames * tokens, 3, heads, dpad, dtype=torch.float16)
    qkv[..., :d] = qkv_true
    hp = heads * dpad
    qkv = qkv.reshape(-1, 3 * hp).to(cuda_dev)
    out = ops.attention(qkv[:, :hp], qkv[:, hp:2 * hp], qkv[:, 2 * hp:], frames, tokens, heads, d, dpad)
    qt = qkv_true.to(cuda_dev).view(frames, tokens, 3, heads, d).permute(2, 0, 3, 1, 4)
    ref = _attn_ref(qt[0], qt[1], qt[2], d ** -0.5).permute(0, 2, 1, 3).reshape(frames * tokens, heads * d)
    assert rel_l2(out, ref) < 3e-3


@pytest.mark.parametrize("B,Fr,N,C", [(2, 16, 256, 320), (2, 16, 64, 1280), (1, 4, 1024, 640), (2, 5, 33, 320),
                                      (1, 24, 16, 64)])
def test_temporal_attention(cuda_dev, B, Fr, N, C):
    from aniportrait_b200 import ops
    heads = 8
    d = C // heads
    qkv = _mk((B * Fr * N, 3 * C), cuda_dev, 1.0, 12)
    out = ops.temporal_attention(qkv, B, Fr, N, C, heads)
    t = qkv.float().view(B, Fr, N, 3, heads, d).permute(3, 0, 2, 4, 1, 5)  # [3, B, N, h, F, d]
    ref = _attn_ref(t[0], t[1], t[2], d ** -0.5)                           # [B, N, h, F, d]
    ref = ref.permute(0, 3, 1, 2, 4).reshape(B * Fr * N, C)
    assert rel_l2(out, ref) < 2e-3


def test_elementwise_and_layout(cuda_dev):
    from aniportrait_b200 import ops
    a = _mk((6, 16, 16, 320), cuda_dev, 1.0, 13)
    b = _mk((6, 16, 16, 320), cuda_dev, 1.0, 14)
    assert rel_l2(ops.add(a, b), a.float() + b.float()) < 1e-3
    assert rel_l2(ops.silu(a), F.silu(a.float())) < 1e-3
    up = ops.upsample2x(a)
    ref = F.interpolate(a.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), ref)
    x = _mk((2, 4, 3, 16, 16), cuda_dev, 1.0, 15)
    nhwc = ops.ncfhw_to_nhwc(x, 64)
    assert nhwc.shape == (6, 16, 16, 64)
    assert torch.equal(nhwc[..., :4], x.permute(0, 2, 3, 4, 1).reshape(6, 16, 16, 4))
    assert nhwc[..., 4:].abs().max().item() == 0
    back = ops.nhwc_to_ncfhw(nhwc, 2, 4, 3)
    assert torch.equal(back, x)


def test_denoise_elementwise(cuda_dev):
    """gather -> (identity 'UNet') -> scatter-accumulate over overlapping windows -> CFG + DDIM step, vs torch."""
    from aniportrait_b200 import ops
    L, h, w = 6, 8, 8
    lat = _mk((L, h, w, 4), cuda_dev, 1.0, 16)
    lat0 = lat.clone()
    windows = [[0, 1, 2, 3], [2, 3, 4, 5], [4, 5, 0, 1]]
    acc = torch.zeros(2, L, h, w, 4, dtype=torch.float32, device=cuda_dev)
    cnt = torch.zeros(L)
    preds = []
    for k, wd in enumerate(windows):
        idx = torch.tensor(wd, dtype=torch.int32, device=cuda_dev)
        x = ops.gather_window(lat, idx, 2, 64)
        assert torch.equal(x[:4, ..., :4], lat[wd]) and torch.equal(x[4:, ..., :4], lat[wd])
        pred = _mk((8, h, w, 32), cuda_dev, 1.0, 20 + k)
        preds.append(pred)
        ops.scatter_accumulate(pred, idx, acc)
        for f in wd:
            cnt[f] += 1
    inv = (1.0 / cnt).to(cuda_dev)
    a_t, a_p, gs = 0.37, 0.52, 3.5
    ops.cfg_ddim_step(acc, inv, gs, a_t, a_p, lat)
    ref_acc = torch.zeros(2, L, h, w, 4, device=cuda_dev)
    for wd, pred in zip(windows, preds):
        p4 = pred[..., :4].float().view(2, 4, h, w, 4)
        for j, f in enumerate(wd):
            ref_acc[:, f] += p4[:, j]
    ref_acc = ref_acc * inv[None, :, None, None, None]
    v = ref_acc[0] + gs * (ref_acc[1] - ref_acc[0])
    x = lat0.float()
    x0 = math.sqrt(a_t) * x - math.sqrt(1 - a_t) * v
    eps = math.sqrt(a_t) * v + math.sqrt(1 - a_t) * x
    ref = math.sqrt(a_p) * x0 + math.sqrt(1 - a_p) * eps
    assert rel_l2(lat, ref) < 1e-3
    assert acc.abs().max().item() == 0


@pytest.mark.parametrize("pred_type,clip", [("v_prediction", 0.0), ("epsilon", 0.0), ("epsilon", 1.0), ("sample", 1.0)])
def test_ddim_step_prediction_types_match_scheduler(cuda_dev, pred_type, clip):
    """The fused CFG + DDIM kernel against the host DDIMScheduler.step (diffusers semantics [dep]) for every prediction type
    the reference's configs use (inference_v2.yaml: v_prediction; inference_v1.yaml: epsilon, clip_sample default True)."""
    from aniportrait_b200 import ops
    from aniportrait_b200.pipelines.scheduler import DDIMScheduler
    sch = DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="linear", clip_sample=clip > 0,
                        clip_sample_range=clip if clip > 0 else 1.0, prediction_type=pred_type,
                        timestep_spacing="leading", steps_offset=1)
    sch.set_timesteps(10)
    t = int(sch.timesteps[3])
    L, h, w = 3, 8, 8
    lat = _mk((L, h, w, 4), cuda_dev, 1.5, 31)
    lat0 = lat.clone()
    acc = _mk((2, L, h, w, 4), cuda_dev, 1.0, 32).float().contiguous()
    acc0 = acc.clone()
    inv = torch.full((L,), 0.5, device=cuda_dev)
    a_t, a_p = sch.alpha_pair(t)
    ops.cfg_ddim_step(acc, inv, 2.0, a_t, a_p, lat, pred_type, clip)
    avg = acc0 * 0.5
    v = avg[0] + 2.0 * (avg[1] - avg[0])
    ref = sch.step(v, t, lat0.float()).prev_sample
    assert rel_l2(lat, ref) < 1e-3
    with pytest.raises(ValueError):
        ops.cfg_ddim_step(acc, inv, 2.0, a_t, a_p, lat, "flow", 0.0)


def test_scatter_accumulate_repeated_frame_counts_once(cuda_dev):
    """A window holding the same frame twice (dilated windows wrapping around a short clip, context_stride > 1): the frame
    receives ONE contribution (its last occurrence), like the reference's index assignment; plan_windows counts it once."""
    from aniportrait_b200 import ops
    from aniportrait_b200.pipelines.sharding import accumulate, plan_windows
    windows, inv = plan_windows(24, 25, "uniform", 16, 2, 4)
    dup = [wd for wd in windows if len(set(wd)) < len(wd)]
    assert dup, "expected a window with a repeated frame at L=24, size=16, stride=2"
    wd = dup[0]
    h = w = 4
    acc = torch.zeros(2, 24, h, w, 4, dtype=torch.float32, device=cuda_dev)
    pred = _mk((2 * len(wd), h, w, 8), cuda_dev, 1.0, 41)
    idx = torch.tensor(wd, dtype=torch.int32, device=cuda_dev)
    for _ in range(3):   # deterministic (no race between the two writers of the repeated frame)
        acc.zero_()
        ops.scatter_accumulate(pred, idx, acc)
        ref = accumulate(torch.zeros(2, 24, h, w, 4), pred[..., :4].float().cpu().view(2, len(wd), h, w, 4), wd)
        assert torch.equal(acc.cpu(), ref)


@pytest.mark.parametrize("rescale", [False, True])
def test_pack_frames_u8_matches_host_bytes(cuda_dev, rescale):
    """On-device 8-bit frame packing == the bytes the reference's save_videos_grid makes on the host from the fp32 tensor
    (src/utils/util.py:94-98: `(x + 1) / 2` if rescale, `(x * 255).numpy().astype(np.uint8)`), bit for bit: every fp16 value
    of the range (incl. the k/255 neighbourhoods where truncation decides the byte), both the four-pixel and the
    any-stride kernel, the decoder's [F, 3, H, W] memory order and a plain contiguous video."""
    import numpy as np
    from aniportrait_b200 import ops

    def host_bytes(video):          # what the scripts do with pipe(...).videos
        x = video.float().cpu()
        if rescale:
            x = (x + 1.0) / 2.0
        return (x * 255).numpy().astype(np.uint8).transpose(0, 2, 3, 4, 1)    # b c f h w -> b f h w c

    lo, hi = (-1.0, 1.0) if rescale else (0.0, 1.0)
    # every finite fp16 bit pattern inside [lo, hi]
    allh = torch.arange(0, 1 << 16, dtype=torch.int32).to(torch.int16).view(torch.float16)
    allh = allh[torch.isfinite(allh) & (allh >= lo) & (allh <= hi)]
    g = torch.Generator().manual_seed(5)
    for (B, F_, H, W), decoder_order in [((1, 4, 64, 64), True), ((2, 3, 16, 32), False), ((1, 2, 10, 21), True),
                                         ((1, 3, 8, 20), False)]:
        n = B * 3 * F_ * H * W
        vals = allh[torch.randint(0, allh.numel(), (n,), generator=g)]
        vals[:allh.numel()] = allh[:n] if allh.numel() > n else allh
        if decoder_order:           # memory [B, F, 3, H, W], handed over as the [B, 3, F, H, W] view the pipeline makes
            video = vals.view(B, F_, 3, H, W).to(cuda_dev).permute(0, 2, 1, 3, 4)
        else:
            video = vals.view(B, 3, F_, H, W).to(cuda_dev)
        out = ops.pack_frames_u8(video, rescale=rescale)
        assert out.shape == (B, F_, H, W, 3) and out.dtype == torch.uint8 and out.is_contiguous()
        assert np.array_equal(out.cpu().numpy(), host_bytes(video)), (B, F_, H, W, decoder_order)
    # a view with a pixel stride of 2 (takes the any-stride kernel even though W % 4 == 0)
    wide = allh[torch.randint(0, allh.numel(), (1 * 3 * 2 * 8 * 32,), generator=g)].view(1, 3, 2, 8, 32).to(cuda_dev)
    sub = wide[..., ::2]
    assert np.array_equal(ops.pack_frames_u8(sub, rescale=rescale).cpu().numpy(), host_bytes(sub))
    if not rescale:                 # out of range saturates, NaN -> 0 (the host cast is undefined there)
        odd = torch.tensor([-0.5, 1.5, float("nan"), 65504.0] * 6, dtype=torch.float16, device=cuda_dev).view(1, 3, 1, 2, 4)
        got = ops.pack_frames_u8(odd).cpu().view(-1)
        exp = torch.tensor([0, 255, 0, 255] * 6, dtype=torch.uint8).view(1, 3, 1, 2, 4).permute(0, 2, 3, 4, 1).reshape(-1)
        assert torch.equal(got, exp)
