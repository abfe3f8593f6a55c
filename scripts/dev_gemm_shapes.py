"""GPU-box helper: time the distinct GEMM / conv shapes of one UNet3D call (512x512, F=16, CFG) and print TFLOP/s."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aniportrait_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
_lib.init(0)
BN = int(os.environ.get("BN", 0))


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gemm_case(name, M, N, K, count, geglu=False, residual=False, bias=True):
    a = torch.randn(M, K, device=dev, dtype=torch.float16)
    w = torch.randn(N, K, device=dev, dtype=torch.float16) * 0.02
    b = torch.zeros(N, device=dev, dtype=torch.float32) if bias else None
    nout = N // 2 if geglu else N
    r = torch.randn(M, nout, device=dev, dtype=torch.float16) if residual else None
    out = torch.empty(M, nout, device=dev, dtype=torch.float16)
    ms = timeit(lambda: ops.gemm(a, w, bias=b, residual=r, geglu=geglu, out=out, block_n=BN))
    fl = 2.0 * M * N * K
    print(f"{name:34s} M={M:6d} N={N:5d} K={K:5d} x{count:3d}  {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TF/s  "
          f"total {ms * count:6.2f} ms")
    return ms * count


def conv_case(name, nf, h, cin, cout, count, stride=1, cin2=0):
    x = torch.randn(nf, h, h, cin, device=dev, dtype=torch.float16)
    x2 = torch.randn(nf, h, h, cin2, device=dev, dtype=torch.float16) if cin2 else None
    w = ops.pack_conv3x3_weight(torch.randn(cout, cin + cin2, 3, 3, device=dev, dtype=torch.float16) * 0.02)
    b = torch.zeros(w.shape[0], device=dev, dtype=torch.float32)
    ho = h // stride
    out = torch.empty(nf, ho, ho, cout, device=dev, dtype=torch.float16)
    ms = timeit(lambda: ops.conv3x3(x, w, cout, bias=b, x2=x2, stride=stride, out=out, block_n=BN))
    fl = 2.0 * nf * ho * ho * cout * 9 * (cin + cin2)
    print(f"{name:34s} {nf}x{h}x{h} {cin + cin2:4d}->{cout:4d} s{stride} x{count:3d}  {ms * 1e3:8.1f} us  "
          f"{fl / ms / 1e9:7.1f} TF/s  total {ms * count:6.2f} ms")
    return ms * count


tot = 0.0
for lvl, (hw, C, nres, ntr, nmm) in enumerate([(64, 320, 7, 5, 5), (32, 640, 7, 5, 5), (16, 1280, 7, 5, 5),
                                               (8, 1280, 9, 1, 6)]):
    M = 32 * hw * hw
    print(f"--- level {lvl}: {hw}x{hw}, C={C}, M={M}")
    tot += conv_case("resnet conv CxC", 32, hw, C, C, 2 * nres)
    tot += gemm_case("linear CxC (+res)", M, C, C, 4 * ntr // 1 + 5 * nmm, residual=True)
    if ntr:
        dp = ops.head_pad(C // 8)
        tot += gemm_case("spatial qkv (head-padded)", M, 3 * 8 * dp, C, ntr, bias=False)
    tot += gemm_case("temporal qkv", M, 3 * C, C, 2 * nmm, bias=False)
    tot += gemm_case("FF1 GEGLU", M, 8 * C, C, ntr + nmm, geglu=True)
    tot += gemm_case("FF2 (+res)", M, C, 4 * C, ntr + nmm, residual=True)
tot += conv_case("up conv 960->320 two-source", 32, 64, 640, 320, 1, cin2=320)
tot += conv_case("up conv 1920->640 two-source", 32, 32, 1280, 640, 1, cin2=640)
tot += conv_case("up conv 2560->1280 two-source", 32, 16, 1280, 1280, 2, cin2=1280)
tot += conv_case("down conv s2 320", 32, 64, 320, 320, 1, stride=2)
tot += conv_case("down conv s2 640", 32, 32, 640, 640, 1, stride=2)
print(f"sum of listed: {tot:.2f} ms")
