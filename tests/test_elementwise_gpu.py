"""Memory-bound sm_100a kernels vs plain PyTorch fp32 references."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _r(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).bfloat16()


@pytest.mark.parametrize("H", [1024, 4096, 5120, 8192, 16384])
def test_rmsnorm_fwd_bwd(H):
    from simumax_b200.ops import elementwise as E
    torch.manual_seed(0)
    x, g, dy = _r(37, H), (1 + 0.1 * torch.randn(H, device="cuda")).bfloat16(), _r(37, H)
    y, rstd = E.rmsnorm_fwd(x, g, 1e-5)
    xf = x.float().requires_grad_(True)
    gf = g.float().requires_grad_(True)
    ref = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * gf
    torch.testing.assert_close(y.float(), ref, atol=3e-2, rtol=2e-2)
    ref.backward(dy.float())
    dgamma = torch.zeros(H, device="cuda")
    dx = E.rmsnorm_bwd(x, g, dy, rstd, dgamma)
    torch.testing.assert_close(dx.float(), xf.grad, atol=3e-2, rtol=2e-2)
    torch.testing.assert_close(dgamma, gf.grad, atol=5e-2, rtol=2e-2)
    base = _r(37, H)
    dx2 = E.rmsnorm_bwd(x, g, dy, rstd, torch.zeros(H, device="cuda"), dx=base.clone(), accumulate_dx=True)
    torch.testing.assert_close(dx2.float(), base.float() + xf.grad, atol=6e-2, rtol=2e-2)


def test_swiglu():
    from simumax_b200.ops import elementwise as E
    torch.manual_seed(1)
    x, dy = _r(64, 2 * 1536), _r(64, 1536)
    y = E.swiglu_fwd(x)
    xf = x.float().requires_grad_(True)
    a, b = xf.chunk(2, -1)
    ref = torch.nn.functional.silu(a) * b
    torch.testing.assert_close(y.float(), ref, atol=2e-2, rtol=2e-2)
    ref.backward(dy.float())
    dx = E.swiglu_bwd(x, dy)
    torch.testing.assert_close(dx.float(), xf.grad, atol=3e-2, rtol=2e-2)


def test_rope_roundtrip_and_reference():
    from simumax_b200.ops import elementwise as E
    torch.manual_seed(2)
    B, S, nq, nk, d = 2, 128, 8, 2, 128
    qkv = _r(B * S, (nq + 2 * nk) * d)
    orig = qkv.clone()
    cos, sin = E.rope_tables(S, d, 10000.0)
    E.rope_inplace(qkv, cos, sin, S, nq + nk, d)
    x = orig.float().view(B, S, nq + 2 * nk, d)[:, :, : nq + nk]
    x1, x2 = x[..., : d // 2], x[..., d // 2:]
    c, s = cos[None, :, None, :], sin[None, :, None, :]
    ref = torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], -1)
    got = qkv.float().view(B, S, nq + 2 * nk, d)
    torch.testing.assert_close(got[:, :, : nq + nk], ref, atol=3e-2, rtol=2e-2)
    torch.testing.assert_close(got[:, :, nq + nk:], orig.float().view(B, S, nq + 2 * nk, d)[:, :, nq + nk:])
    E.rope_inplace(qkv, cos, sin, S, nq + nk, d, inverse=True)
    torch.testing.assert_close(qkv.float(), orig.float(), atol=5e-2, rtol=2e-2)


@pytest.mark.parametrize("V", [16128, 128256, 50264])
def test_cross_entropy(V):
    from simumax_b200.ops import elementwise as E
    torch.manual_seed(3)
    rows = 64
    logits = _r(rows, V, scale=2.0)
    tgt = torch.randint(0, V, (rows,), device="cuda")
    lf = logits.float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lf, tgt, reduction="none")
    (ref.sum() * 0.5).backward()
    work = logits.clone()
    loss = E.cross_entropy_(work, tgt, grad_scale=0.5)
    torch.testing.assert_close(loss, ref.detach(), atol=2e-2, rtol=1e-2)
    torch.testing.assert_close(work.float(), lf.grad, atol=2e-3, rtol=5e-2)


def test_adamw_and_sqnorm():
    from simumax_b200.ops import elementwise as E
    torch.manual_seed(4)
    n = 1 << 20
    w = torch.randn(n, device="cuda"); g = torch.randn(n, device="cuda") * 0.1
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    p = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    sq = E.sqnorm(g)
    torch.testing.assert_close(sq[0], g.pow(2).sum(), rtol=1e-4, atol=1e-2)
    wr = w.clone().requires_grad_(True)
    opt = torch.optim.AdamW([wr], lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    wr.grad = g.clone()
    opt.step()
    E.adamw_(w, m, v, g, p, 1e-2, 0.9, 0.95, 1e-8, 0.1, 1, sqnorm_t=None, zero_grad=True)
    torch.testing.assert_close(w, wr.detach(), atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(p.float(), wr.detach(), atol=2e-2, rtol=1e-2)
    assert float(g.abs().max()) == 0.0


def test_permute_unpermute():
    from simumax_b200.ops import elementwise as E
    torch.manual_seed(5)
    T, H, topk, Ex = 256, 1024, 2, 8
    x = _r(T, H)
    expert = torch.stack([torch.randperm(Ex, device="cuda")[:topk] for _ in range(T)])           # [T, topk]
    flat_expert = expert.flatten()
    order = torch.argsort(flat_expert, stable=True)
    row_map = (order // topk).int()
    permuted = E.permute_rows(x, row_map)
    torch.testing.assert_close(permuted, x[row_map.long()])
    inv = torch.empty_like(order)
    inv[order] = torch.arange(order.numel(), device="cuda")
    probs = torch.rand(T, topk, device="cuda")
    out = E.unpermute_combine(permuted, inv.int().view(T, topk), probs, T, topk)
    ref = (x.float()[:, None, :] * probs[:, :, None]).sum(1)
    torch.testing.assert_close(out.float(), ref, atol=3e-2, rtol=2e-2)


def test_add_copy():
    from simumax_b200.ops import elementwise as E
    a, b = _r(1000, 512), _r(1000, 512)
    torch.testing.assert_close(E.add(a, b).float(), a.float() + b.float(), atol=2e-2, rtol=2e-2)
    dst = torch.empty_like(a)
    E.copy_(dst, a)
    assert torch.equal(dst, a)


@pytest.mark.parametrize("world,H", [(2, 1024), (4, 4096), (8, 8192)])
def test_rs_consume_sum_residual_rmsnorm(world, H):
    """Owner-side kernel of the fused GEMM -> reduce-scatter: slot sum (+ residual) (+ RMSNorm), no counter wait."""
    import ctypes
    from simumax_b200.ops import _lib
    torch.manual_seed(3)
    rows = 96
    slots, resid = _r(world, rows, H), _r(rows, H)
    gamma = (1 + 0.1 * torch.randn(H, device="cuda")).bfloat16()
    out, normed = torch.empty_like(resid), torch.empty_like(resid)
    rstd = torch.empty(rows, device="cuda")
    f = _lib.lib().smx_rs_consume
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_int] + [ctypes.c_void_p] * 5 + [
        ctypes.c_float, ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p]
    st = torch.cuda.current_stream().cuda_stream
    assert f(slots.data_ptr(), world, rows, H, resid.data_ptr(), out.data_ptr(), gamma.data_ptr(), normed.data_ptr(),
             rstd.data_ptr(), 1e-5, None, 0, st) == 0
    ref = resid.float() + slots.float().sum(0)
    torch.testing.assert_close(out.float(), ref, atol=6e-2, rtol=2e-2)
    x2 = out.float()                       # the norm is taken of the bf16-rounded residual stream, as stored
    r_ref = torch.rsqrt(x2.pow(2).mean(-1) + 1e-5)
    torch.testing.assert_close(rstd, r_ref, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(normed.float(), x2 * r_ref[:, None] * gamma.float(), atol=3e-2, rtol=2e-2)
    plain = torch.empty_like(resid)
    assert f(slots.data_ptr(), world, rows, H, None, plain.data_ptr(), None, None, None, 1e-5, None, 0, st) == 0
    torch.testing.assert_close(plain.float(), slots.float().sum(0), atol=6e-2, rtol=2e-2)


@pytest.mark.parametrize("world", [2, 4])
def test_vocab_parallel_cross_entropy(world):
    """Two-phase vocab-parallel CE (local stats -> gathered -> loss + in-place grad) vs fp32 CE on the full vocabulary."""
    from simumax_b200.ops import elementwise as E
    torch.manual_seed(4)
    rows, V = 64, 2048 * world
    VW = V // world
    logits = _r(rows, V, scale=2.0)
    tgt = torch.randint(0, V, (rows,), device="cuda")
    lf = logits.float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lf, tgt, reduction="none")
    ref.sum().backward()
    shards = [logits[:, r * VW:(r + 1) * VW].contiguous() for r in range(world)]
    stats = torch.empty(world, rows, 4, device="cuda")
    for r in range(world):
        E.ce_vocab_parallel_stats(shards[r], tgt, r * VW, stats[r])
    for r in range(world):
        loss = E.ce_vocab_parallel_grad_(shards[r], tgt, r * VW, stats, world, grad_scale=1.0)
        torch.testing.assert_close(loss, ref.detach(), atol=2e-2, rtol=1e-2)
        torch.testing.assert_close(shards[r].float(), lf.grad[:, r * VW:(r + 1) * VW], atol=2e-2, rtol=2e-2)


def test_embedding_gather_and_scatter_add():
    from simumax_b200.ops import elementwise as E
    torch.manual_seed(0)
    V, H, T, v0 = 512, 256, 1000, 128
    table = torch.randn(V, H, device="cuda").bfloat16()
    tok = torch.randint(0, V + 2 * v0, (T,), device="cuda")          # some ids fall outside the shard [v0, v0 + V)
    out = torch.empty(T, H, device="cuda", dtype=torch.bfloat16)
    E.embedding_fwd(tok, table, out, v0=v0)
    local = tok - v0
    mine = (local >= 0) & (local < V)
    ref = table[local.clamp(0, V - 1)] * mine.unsqueeze(1)
    assert torch.equal(out, ref)
    dy = torch.randn(T, H, device="cuda").bfloat16()
    grad = torch.zeros(V, H, device="cuda", dtype=torch.float32)
    E.embedding_bwd(tok, dy, grad, v0=v0)
    ref_g = torch.zeros_like(grad).index_add_(0, local.clamp(0, V - 1), dy.float() * mine.unsqueeze(1))
    assert torch.allclose(grad, ref_g, atol=1e-4, rtol=1e-5)


def test_fused_add_rmsnorm():
    from simumax_b200.ops import elementwise as E
    torch.manual_seed(1)
    for H in (1024, 4096, 8192):
        T = 300
        x, d = torch.randn(T, H, device="cuda").bfloat16(), torch.randn(T, H, device="cuda").bfloat16()
        gamma = (1 + 0.1 * torch.randn(H, device="cuda")).bfloat16()
        s, n = torch.empty_like(x), torch.empty_like(x)
        rstd = torch.empty(T, device="cuda", dtype=torch.float32)
        E.add_rmsnorm_fwd(x, d, gamma, 1e-5, out_sum=s, out_normed=n, rstd=rstd)
        ref_s = (x.float() + d.float()).bfloat16()
        assert torch.equal(s, ref_s)
        r = torch.rsqrt(ref_s.float().pow(2).mean(-1) + 1e-5)
        assert torch.allclose(rstd, r, rtol=1e-4)
        ref_n = ref_s.float() * r[:, None] * gamma.float()
        assert torch.allclose(n.float(), ref_n, atol=2e-2, rtol=1e-2)
