"""Python-side operator wrappers over the C ABI (torch tensors in, torch tensors out).

torch is used for device memory and streams only; every arithmetic op below is a hand-written sm_100a kernel in
aniportrait_b200/csrc reached through libaniportrait_b200.so. Activations are fp16 channels-last token matrices.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import I, LL, check, fptr, lib, ptr, stream_ptr

KERNEL_LAUNCHES = 0  # incremented by every wrapper; bench.py reports it as gpu_launches


def _count(n=1):
    global KERNEL_LAUNCHES
    KERNEL_LAUNCHES += n


def _ensure(t: torch.Tensor):
    if not t.is_cuda:
        raise _lib.ApError("aniportrait_b200 ops need CUDA tensors (no CPU fallback)")
    _lib.init(t.device.index if t.device.index is not None else torch.cuda.current_device())


# --------------------------------------------------------------------------------------------------------------
# weight repacking (done once at load time)
# --------------------------------------------------------------------------------------------------------------
def pack_conv3x3_weight(w: torch.Tensor, cin_pad_to: int = 64, cout_pad_to: int = 32) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [Cout_p, 9*Cin_p] fp16, tap-major / channel-minor, zero padded."""
    cout, cin = w.shape[0], w.shape[1]
    cin_p = (cin + cin_pad_to - 1) // cin_pad_to * cin_pad_to
    cout_p = (cout + cout_pad_to - 1) // cout_pad_to * cout_pad_to
    wp = torch.zeros(cout_p, 3, 3, cin_p, dtype=torch.float16, device=w.device)
    wp[:cout, :, :, :cin] = w.permute(0, 2, 3, 1).to(torch.float16)
    return wp.reshape(cout_p, 9 * cin_p).contiguous()


def pack_conv3x3_weight_two_source(w: torch.Tensor, c1: int) -> torch.Tensor:
    """Conv over cat([x1 (c1 ch), x2]) -> same layout; channel order inside a tap is already [x1, x2]."""
    return pack_conv3x3_weight(w)


def interleave_geglu(w: torch.Tensor, b: torch.Tensor | None):
    """FeedForward.net.0.proj weight [8C, C] (value half then gate half) -> rows interleaved in blocks of 16:
    [v0..15, g0..15, v16..31, g16..31, ...] so that a 32-column accumulator chunk holds matching value/gate pairs."""
    n2 = w.shape[0]
    half = n2 // 2
    assert half % 16 == 0
    idx = torch.arange(half, device=w.device).reshape(-1, 16)
    order = torch.cat([idx, idx + half], dim=1).reshape(-1)
    wi = w[order].contiguous()
    bi = b[order].contiguous() if b is not None else None
    return wi, bi


# --------------------------------------------------------------------------------------------------------------
# GEMM / conv
# --------------------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, residual: torch.Tensor | None = None,
         a2: torch.Tensor | None = None, geglu: bool = False, out: torch.Tensor | None = None,
         bias_group_rows: int = 0, n_valid: int = 0, block_n: int = 0) -> torch.Tensor:
    """out = [a | a2] @ w.T (+bias) (+residual); a:[M,K1] fp16 (row stride may exceed K1), w:[N,K1+K2] fp16."""
    _ensure(a)
    assert a.dtype == torch.float16 and w.dtype == torch.float16 and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.is_contiguous()
    M, K1 = a.shape
    N = w.shape[0]
    K2 = 0
    if a2 is not None:
        assert a2.dtype == torch.float16 and a2.shape[0] == M and a2.stride(1) == 1
        K2 = a2.shape[1]
    assert w.shape[1] == K1 + K2, (w.shape, K1, K2)
    nout = N // 2 if geglu else N
    if n_valid:
        nout = n_valid
    if out is None:
        out = torch.empty(M, nout, dtype=torch.float16, device=a.device)
    assert out.stride(1) == 1
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.shape[-1] == N
    if residual is not None:
        assert residual.dtype == torch.float16 and residual.stride(1) == 1 and residual.shape[0] == M
    rc = lib().ap_gemm_f16(ptr(a), LL(a.stride(0)), I(K1), ptr(a2), LL(a2.stride(0) if a2 is not None else 0), I(K2),
                           ptr(w), LL(M), I(N), fptr(bias), LL(bias_group_rows), ptr(residual),
                           LL(residual.stride(0) if residual is not None else 0), ptr(out), LL(out.stride(0)),
                           I(nout), I(1 if geglu else 0), I(block_n), stream_ptr())
    check(rc, "ap_gemm_f16")
    _count()
    return out


def conv3x3(x: torch.Tensor, w_packed: torch.Tensor, cout: int, bias: torch.Tensor | None = None,
            residual: torch.Tensor | None = None, x2: torch.Tensor | None = None, stride: int = 1,
            out: torch.Tensor | None = None, bias_group_rows: int = 0, block_n: int = 0) -> torch.Tensor:
    """x: [Nf, H, W, C1] fp16 channels-last (C1 % 64 == 0); w_packed: pack_conv3x3_weight(...); returns
    [Nf, H/stride, W/stride, cout]."""
    _ensure(x)
    assert x.dtype == torch.float16 and x.is_contiguous() and x.dim() == 4
    nf, h, wd, c1 = x.shape
    c2 = 0
    if x2 is not None:
        assert x2.is_contiguous() and x2.shape[:3] == x.shape[:3]
        c2 = x2.shape[3]
    cout_p = w_packed.shape[0]
    assert w_packed.shape[1] == 9 * (c1 + c2), (w_packed.shape, c1, c2)
    ho, wo = h // stride, wd // stride
    if out is None:
        out = torch.empty(nf, ho, wo, cout, dtype=torch.float16, device=x.device)
    assert out.is_contiguous()
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.shape[-1] == cout_p
    if residual is not None:
        assert residual.is_contiguous() and residual.shape == out.shape
    rc = lib().ap_conv3x3_nhwc_f16(ptr(x), I(c1), ptr(x2), I(c2), I(nf), I(h), I(wd), I(stride), ptr(w_packed),
                                   I(cout_p), fptr(bias), LL(bias_group_rows), ptr(residual), ptr(out), LL(cout),
                                   I(cout), I(block_n), stream_ptr())
    check(rc, "ap_conv3x3_nhwc_f16")
    _count()
    return out
