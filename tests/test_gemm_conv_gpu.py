"""GPU parity: tcgen05 GEMM / implicit-GEMM conv vs a plain fp32 torch evaluation of the same op on the same fp16
inputs. Tolerance: the kernel accumulates in fp32 and rounds once to fp16 -> rel-L2 <= 2e-3 (fp16 eps ~ 1e-3)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 2e-3


def rel_l2(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _mk(shape, dev, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.float16).to(dev)


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 256, 64, 0), (256, 256, 320, 0), (1000, 320, 320, 0), (4096, 640, 1280, 0), (333, 1280, 2560, 0),
    (2, 1280, 320, 0), (8192, 960, 320, 0), (512, 128, 128, 0), (512, 64, 192, 0), (300, 32, 320, 0),
    (16384, 2560, 320, 256), (4096, 320, 2880, 160),
])
def test_gemm_plain(cuda_dev, M, N, K, bn):
    from aniportrait_b200 import ops
    a = _mk((M, K), cuda_dev, 1.0, 1)
    w = _mk((N, K), cuda_dev, K ** -0.5, 2)
    out = ops.gemm(a, w, block_n=bn)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    assert rel_l2(out, ref) < TOL


def test_gemm_bias_residual_groups(cuda_dev):
    from aniportrait_b200 import ops
    M, N, K = 2048, 640, 640
    a = _mk((M, K), cuda_dev, 1.0, 3)
    w = _mk((N, K), cuda_dev, K ** -0.5, 4)
    bias = torch.randn(2, N, device=cuda_dev, dtype=torch.float32)
    res = _mk((M, N), cuda_dev, 1.0, 5)
    out = ops.gemm(a, w, bias=bias, residual=res, bias_group_rows=M // 2)
    ref = a.float() @ w.float().t() + bias.repeat_interleave(M // 2, dim=0) + res.float()
    assert rel_l2(out, ref) < TOL


def test_gemm_two_source(cuda_dev):
    from aniportrait_b200 import ops
    M, N, K1, K2 = 1024, 640, 1280, 640
    a1 = _mk((M, K1), cuda_dev, 1.0, 6)
    a2 = _mk((M, K2), cuda_dev, 1.0, 7)
    w = _mk((N, K1 + K2), cuda_dev, (K1 + K2) ** -0.5, 8)
    out = ops.gemm(a1, w, a2=a2)
    ref = torch.cat([a1, a2], 1).float() @ w.float().t()
    assert rel_l2(out, ref) < TOL


def test_gemm_strided_a_and_out(cuda_dev):
    """A is a column slice of a wider matrix (e.g. one of q/k/v inside a fused qkv buffer)."""
    from aniportrait_b200 import ops
    M, N, K = 512, 320, 320
    big = _mk((M, 3 * K), cuda_dev, 1.0, 9)
    a = big[:, K:2 * K]
    w = _mk((N, K), cuda_dev, K ** -0.5, 10)
    outbig = torch.zeros(M, 2 * N, dtype=torch.float16, device=cuda_dev)
    ops.gemm(a, w, out=outbig[:, N:])
    ref = a.float() @ w.float().t()
    assert rel_l2(outbig[:, N:], ref) < TOL
    assert outbig[:, :N].abs().max().item() == 0


def test_gemm_geglu(cuda_dev):
    from aniportrait_b200 import ops
    M, C = 1024, 320
    a = _mk((M, C), cuda_dev, 1.0, 11)
    w = _mk((8 * C, C), cuda_dev, C ** -0.5, 12)
    b = torch.randn(8 * C, device=cuda_dev, dtype=torch.float32) * 0.1
    wi, bi = ops.interleave_geglu(w, b)
    out = ops.gemm(a, wi, bias=bi, geglu=True)
    h = a.float() @ w.float().t() + b
    v, g = h.chunk(2, dim=-1)
    ref = v * F.gelu(g)
    assert out.shape == (M, 4 * C)
    assert rel_l2(out, ref) < TOL


def _conv_ref(x_nhwc, w, bias, stride):
    x = x_nhwc.float().permute(0, 3, 1, 2)
    y = F.conv2d(x, w.float(), bias, stride=stride, padding=1)
    return y.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("nf,h,w,cin,cout,stride", [
    (2, 64, 64, 320, 320, 1), (3, 32, 32, 640, 640, 1), (4, 16, 16, 1280, 1280, 1), (5, 8, 8, 1280, 1280, 1),
    (2, 64, 64, 320, 320, 2), (2, 32, 32, 640, 640, 2), (3, 16, 16, 1280, 1280, 2),
    (1, 64, 64, 64, 320, 1), (2, 64, 64, 320, 4, 1), (1, 128, 128, 128, 128, 1), (1, 48, 96, 64, 64, 1),
])
def test_conv3x3(cuda_dev, nf, h, w, cin, cout, stride):
    from aniportrait_b200 import ops
    x = _mk((nf, h, w, cin), cuda_dev, 1.0, 20)
    wt = _mk((cout, cin, 3, 3), cuda_dev, (9 * cin) ** -0.5, 21)
    bias = torch.randn(cout, device=cuda_dev, dtype=torch.float32)
    wp = ops.pack_conv3x3_weight(wt)
    bp = torch.zeros(wp.shape[0], device=cuda_dev, dtype=torch.float32)
    bp[:cout] = bias
    out = ops.conv3x3(x, wp, cout, bias=bp, stride=stride)
    ref = _conv_ref(x, wt, bias, stride)
    assert out.shape == ref.shape
    assert rel_l2(out, ref) < TOL


def test_conv3x3_two_source_residual(cuda_dev):
    from aniportrait_b200 import ops
    nf, h, w, c1, c2, cout = 2, 32, 32, 640, 320, 640
    x1 = _mk((nf, h, w, c1), cuda_dev, 1.0, 30)
    x2 = _mk((nf, h, w, c2), cuda_dev, 1.0, 31)
    wt = _mk((cout, c1 + c2, 3, 3), cuda_dev, (9 * (c1 + c2)) ** -0.5, 32)
    res = _mk((nf, h, w, cout), cuda_dev, 1.0, 33)
    wp = ops.pack_conv3x3_weight(wt)
    out = ops.conv3x3(x1, wp, cout, x2=x2, residual=res)
    ref = _conv_ref(torch.cat([x1, x2], -1), wt, None, 1) + res.float()
    assert rel_l2(out, ref) < TOL
