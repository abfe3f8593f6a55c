"""GPU-box helper for `ncu --set full --profile-from-start off`: launches the two dominant kernels once each
(implicit-GEMM conv 320->320 @64x64x32 frames; reference attention at the 64x64 level) inside a profiler range."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aniportrait_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
_lib.init(0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
x = torch.randn(32, 64, 64, 320, device=dev, dtype=torch.float16)
wt = ops.pack_conv3x3_weight(torch.randn(320, 320, 3, 3, device=dev, dtype=torch.float16) * 0.02)
b = torch.zeros(320, device=dev, dtype=torch.float32)
out = torch.empty(32, 64, 64, 320, device=dev, dtype=torch.float16)
n, heads, d, dpad, fr = 4096, 8, 40, 64, 32
hp = heads * dpad
qkv = torch.randn(fr * n, 3 * hp, device=dev, dtype=torch.float16)
bank = torch.randn(n, 2 * hp, device=dev, dtype=torch.float16)
o = torch.empty(fr * n, heads * d, device=dev, dtype=torch.float16)
a = torch.randn(131072, 320, device=dev, dtype=torch.float16)
w2 = torch.randn(320, 320, device=dev, dtype=torch.float16) * 0.05
res = torch.randn(131072, 320, device=dev, dtype=torch.float16)
o2 = torch.empty(131072, 320, device=dev, dtype=torch.float16)


def conv():
    ops.conv3x3(x, wt, 320, bias=b, out=out)


def attn():
    ops.attention(qkv[:, :hp], qkv[:, hp:2 * hp], qkv[:, 2 * hp:], fr, n, heads, d, dpad, bank_k=bank[:, :hp],
                  bank_v=bank[:, hp:], bank_tokens=n, n_banks=1, first_bank_frame=16, frames_per_bank=16, out=o)


def lin():
    ops.gemm(a, w2, bias=b, residual=res, out=o2)


# round 2: the epilogue-bound K = 320 consumers of a folded LayerNorm (FF1 GEGLU 2560 x 320, spatial q/k/v 1536 x 320), fed by
# a producer that emits the row statistics
from aniportrait_b200.models.blocks import fold_layer_norm  # noqa: E402

gam, bet = torch.ones(320, device=dev), torch.zeros(320, device=dev)
wg1, bg1 = fold_layer_norm(torch.randn(2560, 320, device=dev) * 0.05, torch.zeros(2560, device=dev), gam, bet)
wg1, bg1 = ops.interleave_geglu(wg1, bg1)
wq, bq = fold_layer_norm(torch.randn(1536, 320, device=dev) * 0.05, None, gam, bet)
_, rs = ops.gemm(a, w2, bias=b, residual=res, out=o2, row_stats=True)
og = torch.empty(131072, 1280, device=dev, dtype=torch.float16)
oq = torch.empty(131072, 1536, device=dev, dtype=torch.float16)


def geglu():
    ops.gemm(o2, wg1, bias=bg1, geglu=True, out=og, ln=ops.LNFold(rs, 1e-5))


def qkv_fold():
    ops.gemm(o2, wq, bias=bq, out=oq, ln=ops.LNFold(rs, 1e-5))


for f in (conv, attn, lin, geglu, qkv_fold):
    f()
torch.cuda.synchronize()
torch.cuda.profiler.start()
if which in ("all", "conv"):
    conv()
if which in ("all", "attn"):
    attn()
if which in ("all", "lin"):
    lin()
if which in ("all", "geglu"):
    geglu()
if which in ("all", "qkv"):
    qkv_fold()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
