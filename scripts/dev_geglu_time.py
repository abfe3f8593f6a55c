"""Dev helper: time the FeedForward GEGLU projections of the three UNet levels (hot loop, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aniportrait_b200 import _lib, ops
dev = torch.device("cuda:0"); _lib.init(0)
for M, C in [(131072, 320), (32768, 640), (8192, 1280)]:
    a = torch.randn(M, C, device=dev, dtype=torch.float16)
    w = torch.randn(8 * C, C, device=dev, dtype=torch.float16) * 0.05
    b = torch.randn(8 * C, device=dev) * 0.1
    wi, bi = ops.interleave_geglu(w, b)
    out = torch.empty(M, 4 * C, device=dev, dtype=torch.float16)
    for _ in range(3): ops.gemm(a, wi, bias=bi, geglu=True, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.gemm(a, wi, bias=bi, geglu=True, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"GEGLU {M}x{8*C}x{C}: {ms*1e3:.1f} us  {2.0*M*8*C*C/ms/1e9:.0f} TF/s", flush=True)
