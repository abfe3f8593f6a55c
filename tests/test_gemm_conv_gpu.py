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


# ------------------------------------------------------------------------------------------------------------------
# statistics fused into the epilogue + LayerNorm folding (reference: the standalone nn.LayerNorm / nn.GroupNorm passes of
# src/models/attention.py:331-362, motion_module.py:228-241, resnet.py:221-238)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(1000, 320, 320), (4096, 640, 640), (131072 // 8, 320, 320), (700, 1280, 1280),
                                   (2048, 128, 64)])
def test_gemm_row_stats(cuda_dev, M, N, K):
    """{sum, sumsq} of every output row, as partials written by the epilogue warps, equal the sums of the fp16 output."""
    from aniportrait_b200 import ops
    a = _mk((M, K), cuda_dev, 1.0, 31)
    w = _mk((N, K), cuda_dev, K ** -0.5, 32)
    res = _mk((M, N), cuda_dev, 1.0, 33) + 0.5
    bias = torch.randn(N, device=cuda_dev, dtype=torch.float32)
    out, rs = ops.gemm(a, w, bias=bias, residual=res, row_stats=True)
    assert rel_l2(out, a.float() @ w.float().t() + bias + res.float()) < TOL
    tot = rs.buf[:, :M].sum(0)                                    # [M, 2]
    o = out.float()
    assert rel_l2(tot[:, 0], o.sum(1)) < 1e-4
    assert rel_l2(tot[:, 1], (o * o).sum(1)) < 1e-4
    out2, rs2 = ops.gemm(a, w, bias=bias, residual=res, row_stats=True)
    assert torch.equal(rs.buf[:, :M], rs2.buf[:, :M]), "partials must be bit-reproducible"


@pytest.mark.parametrize("M,C,N,geglu", [(1000, 320, 960, False), (4096, 640, 1920, False), (640, 1280, 3840, False),
                                        (1000, 320, 2560, True), (2048, 640, 5120, True), (16384, 320, 1536, False)])
def test_gemm_layernorm_folding(cuda_dev, M, C, N, geglu):
    """x = producer GEMM output (row statistics from its epilogue); consumer = LN(x) W^T + b (optionally GEGLU) with the
    LayerNorm folded: weights [W diag(gamma) | colsum], activation [x | -mean] (one extra k-block on the tensor core),
    epilogue rstd * acc + (W beta + b). vs fp32 torch."""
    from aniportrait_b200 import ops
    from aniportrait_b200.models.blocks import fold_layer_norm
    a = _mk((M, C), cuda_dev, 1.0, 41)
    w0 = _mk((C, C), cuda_dev, C ** -0.5, 42)
    res = _mk((M, C), cuda_dev, 1.0, 43) + 0.3          # non-zero row means
    x, rs = ops.gemm(a, w0, residual=res, row_stats=True)
    g = torch.Generator().manual_seed(44)
    gamma = (1.0 + 0.2 * torch.randn(C, generator=g)).to(cuda_dev)
    beta = (0.1 * torch.randn(C, generator=g)).to(cuda_dev)
    w = (torch.randn(N, C, generator=g) * C ** -0.5).to(cuda_dev)
    b = (0.1 * torch.randn(N, generator=g)).to(cuda_dev)
    n = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    ref = n @ w.half().float().t() + b
    wg, bias = fold_layer_norm(w, b, gamma, beta)          # [N, C + 8]: W diag(gamma) | colsum (hi, hi, lo) | 0
    assert wg.shape == (N, C + ops.LN_EXTRA_K)
    if geglu:
        wi, bi = ops.interleave_geglu(wg, bias)
        out = ops.gemm(x, wi, bias=bi, geglu=True, ln=ops.LNFold(rs, 1e-5))
        half = N // 2
        ref = ref[:, :half] * F.gelu(ref[:, half:])
    else:
        out = ops.gemm(x, wg, bias=bias, ln=ops.LNFold(rs, 1e-5))
    err = rel_l2(out, ref)
    print(f"LN folding M={M} C={C} N={N} geglu={geglu}: rel-L2 = {err:.3e}")
    assert err < 3e-3


def test_gemm_bias_table_slice(cuda_dev):
    """bias given as a column slice of a wider fp32 table (row stride > N): the batched time-embedding projection."""
    from aniportrait_b200 import ops
    M, N, K = 1024, 320, 320
    a = _mk((M, K), cuda_dev, 1.0, 51)
    w = _mk((N, K), cuda_dev, K ** -0.5, 52)
    table = torch.randn(2, 3 * N, device=cuda_dev, dtype=torch.float32)
    out = ops.gemm(a, w, bias=table[:, N:2 * N], bias_group_rows=M // 2)
    ref = a.float() @ w.float().t() + table[:, N:2 * N].repeat_interleave(M // 2, dim=0)
    assert rel_l2(out, ref) < TOL
    x = _mk((4, 16, 16, 64), cuda_dev, 1.0, 53)
    wc = _mk((320, 64, 3, 3), cuda_dev, (9 * 64) ** -0.5, 54)
    outc = ops.conv3x3(x, ops.pack_conv3x3_weight(wc), 320, bias=table[:, :N], bias_group_rows=2 * 256)
    refc = F.conv2d(x.permute(0, 3, 1, 2).float(), wc.float(), padding=1).permute(0, 2, 3, 1)
    refc = refc + table[:, :N].repeat_interleave(2, dim=0)[:, None, None, :]
    assert rel_l2(outc, refc) < TOL


@pytest.mark.parametrize("nf,h,w,cin,cout,stride", [(4, 32, 32, 320, 320, 1), (3, 16, 16, 640, 640, 1), (6, 8, 8, 1280, 1280, 1),
                                                   (2, 64, 64, 320, 320, 2), (5, 16, 24, 128, 320, 1), (8, 4, 8, 256, 256, 1)])
def test_conv_col_stats_feed_group_norm(cuda_dev, nf, h, w, cin, cout, stride):
    """GroupNorm whose statistics come from the producing conv's epilogue == GroupNorm with its own statistics pass == torch."""
    from aniportrait_b200 import ops
    x = _mk((nf, h, w, cin), cuda_dev, 1.0, 61)
    wt = _mk((cout, cin, 3, 3), cuda_dev, (9 * cin) ** -0.5, 62)
    b = torch.randn(cout, device=cuda_dev, dtype=torch.float32)
    ho, wo = h // stride, w // stride
    assert ops.conv_col_stats_ok(nf, ho, wo)
    y, cs = ops.conv3x3(x, ops.pack_conv3x3_weight(wt), cout, bias=b, stride=stride, col_stats=True)
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), wt.float(), b, stride=stride, padding=1).permute(0, 2, 3, 1)
    assert rel_l2(y, ref) < TOL
    # partials: entry e = rows [32e, 32e+32) of the output, frame-major
    yf = y.float().view(nf, ho * wo // 32, 32, cout)
    tot = cs.buf[:nf * ho * wo // 32].view(nf, ho * wo // 32, cout, 2).sum(1)          # per frame, per channel
    assert rel_l2(tot[..., 0], yf.sum((1, 2))) < 1e-4
    assert rel_l2(tot[..., 1], (yf * yf).sum((1, 2))) < 1e-4
    gamma = torch.randn(cout, device=cuda_dev)
    beta = torch.randn(cout, device=cuda_dev)
    fused = ops.group_norm(y, gamma, beta, 32, 1e-5, True, stats=cs)
    plain = ops.group_norm(y, gamma, beta, 32, 1e-5, True)
    tref = F.silu(F.group_norm(y.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)).permute(0, 2, 3, 1)
    assert rel_l2(fused, tref) < TOL and rel_l2(plain, tref) < TOL
    assert rel_l2(fused, plain) < 1e-3


def test_gemm_col_stats_two_source_group_norm(cuda_dev):
    """The up-block case: GroupNorm over the channel concat of two tensors, each with ColStats from its own producer (a GEMM
    with residual and a conv); groups straddle the two sources (C = 640 + 320 = 960, 30 channels per group)."""
    from aniportrait_b200 import ops
    nf, h, w = 4, 16, 16
    a = _mk((nf * h * w, 640), cuda_dev, 1.0, 71)
    w1 = _mk((640, 640), cuda_dev, 640 ** -0.5, 72)
    r1 = _mk((nf * h * w, 640), cuda_dev, 1.0, 73)
    x1, cs1 = ops.gemm(a, w1, residual=r1, col_stats=True)
    xin = _mk((nf, h, w, 320), cuda_dev, 1.0, 74)
    wc = _mk((320, 320, 3, 3), cuda_dev, (9 * 320) ** -0.5, 75)
    x2, cs2 = ops.conv3x3(xin, ops.pack_conv3x3_weight(wc), 320, col_stats=True)
    gamma = torch.randn(960, device=cuda_dev)
    beta = torch.randn(960, device=cuda_dev)
    x1v = x1.view(nf, h, w, 640)
    fused = ops.group_norm(x1v, gamma, beta, 32, 1e-5, True, x2=x2, stats=cs1, stats2=cs2)
    cat = torch.cat([x1v, x2], -1).float().permute(0, 3, 1, 2)
    tref = F.silu(F.group_norm(cat, 32, gamma, beta, 1e-5)).permute(0, 2, 3, 1)
    assert rel_l2(fused, tref) < TOL
    # a source without statistics -> the standalone pass for both (same values)
    plain = ops.group_norm(x1v, gamma, beta, 32, 1e-5, True, x2=x2, stats=cs1, stats2=None)
    assert rel_l2(plain, tref) < TOL
