"""tcgen05 GEMM numerics vs an fp32 PyTorch reference (all three training layouts, grouped, fp8)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b, layout):
    a32, b32 = a.float(), b.float()
    if layout == "TN":
        return a32 @ b32.transpose(-1, -2)
    if layout == "NN":
        return a32 @ b32
    if layout == "NT":
        return a32.transpose(-1, -2) @ b32
    raise ValueError(layout)


def _mk(shape, dtype=torch.bfloat16):
    return (torch.randn(shape, device="cuda", dtype=torch.float32) * 0.5).to(dtype)


def _operands(layout, M, N, K, G=None):
    pre = () if G is None else (G,)
    if layout == "TN":
        return _mk(pre + (M, K)), _mk(pre + (N, K))
    if layout == "NN":
        return _mk(pre + (M, K)), _mk(pre + (K, N))
    return _mk(pre + (K, M)), _mk(pre + (K, N))


@pytest.mark.parametrize("layout", ["TN", "NN", "NT"])
@pytest.mark.parametrize("shape", [(256, 256, 128), (512, 768, 320), (384, 160, 512), (136, 72, 200), (130, 70, 100), (4096, 1024, 1024)])
@pytest.mark.parametrize("block_n", [0, 64, 128, 256])
def test_gemm_layouts(layout, shape, block_n):
    from simumax_b200.ops.gemm import gemm
    torch.manual_seed(0)
    M, N, K = shape
    a, b = _operands(layout, M, N, K)
    out = gemm(a, b, layout, block_n=block_n)
    ref = _ref(a, b, layout)
    torch.testing.assert_close(out.float(), ref, atol=2e-2 * (K ** 0.5) / 8, rtol=2e-2)


@pytest.mark.parametrize("layout", ["TN", "NT"])
def test_gemm_fp32_accumulate(layout):
    from simumax_b200.ops.gemm import gemm
    torch.manual_seed(1)
    M, N, K = 512, 384, 1024
    a, b = _operands(layout, M, N, K)
    acc = torch.randn(M, N, device="cuda", dtype=torch.float32)
    ref = acc + _ref(a, b, layout)
    out = gemm(a, b, layout, out=acc, accumulate=True)
    assert out.data_ptr() == acc.data_ptr()
    torch.testing.assert_close(out, ref, atol=1e-2, rtol=1e-3)
    o2 = gemm(a, b, layout, out_dtype=torch.float32)
    torch.testing.assert_close(o2, _ref(a, b, layout), atol=1e-2, rtol=1e-3)


@pytest.mark.parametrize("layout", ["TN", "NN", "NT"])
def test_grouped_gemm(layout):
    from simumax_b200.ops.gemm import gemm
    torch.manual_seed(2)
    G, M, N, K = 5, 320, 384, 256
    a, b = _operands(layout, M, N, K, G)
    out = gemm(a, b, layout)
    ref = _ref(a, b, layout)
    torch.testing.assert_close(out.float(), ref, atol=5e-2, rtol=2e-2)


def test_gemm_strided_view():
    from simumax_b200.ops.gemm import gemm
    torch.manual_seed(3)
    big = _mk((512, 1024))
    a = big[:, 256:768]           # row stride 1024, inner contiguous
    w = _mk((384, 512))
    out = gemm(a, w, "TN")
    torch.testing.assert_close(out.float(), a.float() @ w.float().t(), atol=5e-2, rtol=2e-2)


def test_gemm_fp8():
    from simumax_b200.ops.gemm import gemm
    torch.manual_seed(4)
    M, N, K = 512, 512, 1024
    a = _mk((M, K)).to(torch.float8_e4m3fn)
    b = _mk((N, K)).to(torch.float8_e4m3fn)
    out = gemm(a, b, "TN", alpha=0.5)
    ref = 0.5 * (a.float() @ b.float().t())
    torch.testing.assert_close(out.float(), ref, atol=1e-1, rtol=2e-2)


@pytest.mark.parametrize("layout", ["TN", "NN", "NT", "TT"])
@pytest.mark.parametrize("shape", [(256, 256, 64), (512, 768, 320), (1000, 520, 200), (2048, 4096, 512), (4096, 4096 + 256, 192)])
@pytest.mark.parametrize("mode", ["bf16", "f32_accumulate"])
def test_gemm_cta_pair_kernel(layout, shape, mode):
    """cta_group::2 kernel (256 x 256 pair tiles, tail round sliced) against fp32 torch, every layout."""
    from simumax_b200.ops.gemm import gemm
    torch.manual_seed(11)
    M, N, K = shape
    a = torch.randn((K, M) if layout in ("NT", "TT") else (M, K), device="cuda").bfloat16()
    b = torch.randn((K, N) if layout in ("NN", "NT") else (N, K), device="cuda").bfloat16()
    af = a.float().t() if layout in ("NT", "TT") else a.float()
    bf = b.float() if layout in ("NN", "NT") else b.float().t()
    ref = af @ bf
    if mode == "bf16":
        out = gemm(a, b, layout, two_cta=True)
        torch.testing.assert_close(out.float(), ref, atol=2e-2 * ref.abs().max().item(), rtol=2e-2)
        one = gemm(a, b, layout, two_cta=False)
        torch.testing.assert_close(out.float(), one.float(), atol=1e-2 * ref.abs().max().item(), rtol=1e-2)
    else:
        c = torch.full((M, N), 0.5, device="cuda", dtype=torch.float32)
        gemm(a, b, layout, out=c, accumulate=True, two_cta=True)
        torch.testing.assert_close(c, ref + 0.5, atol=1e-3 * ref.abs().max().item(), rtol=1e-4)


@pytest.mark.parametrize("shape", [(128, 256, 128), (200, 300, 384), (1024, 768, 1024), (4096, 512, 2048)])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_gemm_mxfp8_block_scaled(shape, out_dtype):
    """tcgen05 kind::mxf8f6f4.block_scale: result must equal the matmul of the dequantised operands; every 32-block gets
    a different power-of-two magnitude so a misplaced scale factor cannot go unnoticed."""
    from simumax_b200.ops.gemm import gemm_mxfp8, mxfp8_dequantize, mxfp8_quantize
    torch.manual_seed(13)
    M, N, K = shape
    x = torch.randn(M, K, device="cuda") * torch.exp2(torch.randint(-6, 6, (M, K // 32), device="cuda").float()).repeat_interleave(32, 1)
    w = torch.randn(N, K, device="cuda") * torch.exp2(torch.randint(-6, 6, (N, K // 32), device="cuda").float()).repeat_interleave(32, 1)
    qa, sa = mxfp8_quantize(x)
    qb, sb = mxfp8_quantize(w)
    ref = mxfp8_dequantize(qa, sa) @ mxfp8_dequantize(qb, sb).t()
    out = gemm_mxfp8(qa, sa, qb, sb, out_dtype=out_dtype)
    tol = 2e-5 if out_dtype == torch.float32 else 1e-2
    assert ((out.float() - ref).abs().max() / ref.abs().max()).item() < tol
    # and the quantisation itself is faithful to the unquantised product to fp8 precision
    assert ((ref - x @ w.t()).norm() / (x @ w.t()).norm()).item() < 6e-2
