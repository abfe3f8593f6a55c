"""Dev helper: time the HBM-bound kernels at their level-0 shapes (L2 flushed between launches) against the HBM roofline."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aniportrait_b200 import _lib, ops
dev = torch.device("cuda:0"); _lib.init(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def t(name, fn, nbytes, iters=10):
    for _ in range(2): fn()
    ms = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    ms /= iters
    print(f"{name:44s} {ms*1e3:8.1f} us  {nbytes/ms/1e6:7.0f} GB/s", flush=True)

for (B, F, N, C) in [(2, 16, 4096, 320), (2, 16, 1024, 640), (2, 16, 256, 1280), (2, 16, 64, 1280)]:
    qkv = torch.randn(B * F * N, 3 * C, device=dev, dtype=torch.float16)
    o = torch.empty(B * F * N, C, device=dev, dtype=torch.float16)
    t(f"temporal_attention N={N} C={C}", lambda: ops.temporal_attention(qkv, B, F, N, C, 8, out=o), qkv.numel() * 2 + o.numel() * 2)
x = torch.randn(32, 4096, 320, device=dev, dtype=torch.float16)
g = torch.ones(320, device=dev); b = torch.zeros(320, device=dev)
y = torch.empty_like(x)
t("group_norm+silu 32x4096x320 (stats+apply)", lambda: ops.group_norm(x, g, b, 32, 1e-5, silu=True, out=y), x.numel() * 2 * 3)
x2 = x.view(-1, 320)
y2 = torch.empty_like(x2)
t("layer_norm 131072x320", lambda: ops.layer_norm(x2, g, b, out=y2), x.numel() * 4)
for rows, c in [(32768, 640), (8192, 1280), (2048, 1280)]:
    xx = torch.randn(rows, c, device=dev, dtype=torch.float16)
    gg = torch.ones(c, device=dev); bb = torch.zeros(c, device=dev)
    yy = torch.empty_like(xx)
    t(f"layer_norm {rows}x{c}", lambda: ops.layer_norm(xx, gg, bb, out=yy), xx.numel() * 4)
