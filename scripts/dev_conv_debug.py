"""GPU-box experiment: pace of the conv kernel with TMA loads or MMAs disabled (AP_GEMM_DEBUG=1|2), per BLOCK_N / CG."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aniportrait_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
_lib.init(0)
x = torch.randn(32, 64, 64, 320, device=dev, dtype=torch.float16)
wt = ops.pack_conv3x3_weight(torch.randn(320, 320, 3, 3, device=dev, dtype=torch.float16) * 0.02)
b = torch.zeros(320, device=dev, dtype=torch.float32)
out = torch.empty(32, 64, 64, 320, device=dev, dtype=torch.float16)
x2 = torch.randn(32, 16, 16, 1280, device=dev, dtype=torch.float16)
wt2 = ops.pack_conv3x3_weight(torch.randn(1280, 1280, 3, 3, device=dev, dtype=torch.float16) * 0.02)
b2 = torch.zeros(1280, device=dev, dtype=torch.float32)
out2 = torch.empty(32, 16, 16, 1280, device=dev, dtype=torch.float16)


def t(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20


ms = t(lambda: ops.conv3x3(x, wt, 320, bias=b, out=out))
ms2 = t(lambda: ops.conv3x3(x2, wt2, 1280, bias=b2, out=out2))
print(f"DEBUG={os.environ.get('AP_GEMM_DEBUG', '0')} CG={os.environ.get('AP_GEMM_CG', 'auto')}: conv320@64 {ms * 1e3:.1f} us "
      f"({2.0 * 131072 * 320 * 2880 / ms / 1e9:.0f} TF/s), conv1280@16 {ms2 * 1e3:.1f} us "
      f"({2.0 * 8192 * 1280 * 11520 / ms2 / 1e9:.0f} TF/s)")
