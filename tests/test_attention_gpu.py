"""tcgen05 flash-attention forward vs an fp32 PyTorch reference."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, k, v, causal, scale):
    B, Sq, H, D = q.shape
    Hk = k.shape[2]
    qf = q.float().permute(0, 2, 1, 3)
    kf = k.float().permute(0, 2, 1, 3).repeat_interleave(H // Hk, dim=1)
    vf = v.float().permute(0, 2, 1, 3).repeat_interleave(H // Hk, dim=1)
    s = qf @ kf.transpose(-1, -2) * scale
    if causal:
        Skv = k.shape[1]
        mask = torch.ones(Sq, Skv, device=q.device, dtype=torch.bool).tril(Skv - Sq)
        s = s.masked_fill(~mask, float("-inf"))
    lse = torch.logsumexp(s, -1)
    o = torch.softmax(s, -1) @ vf
    return o.permute(0, 2, 1, 3), lse


@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("cfg", [(1, 256, 4, 4, 128, 128), (2, 512, 8, 2, 128, 128), (1, 384, 4, 1, 64, 64),
                                 (1, 1024, 4, 4, 192, 128), (1, 200, 2, 2, 128, 128), (1, 640, 6, 2, 128, 128), (1, 1152, 8, 4, 128, 128),
                                 (2, 2048, 4, 1, 128, 128)])
def test_fmha_fwd(cfg, causal):
    from simumax_b200.ops.attention import fmha_fwd
    torch.manual_seed(0)
    B, S, H, Hk, D, Dv = cfg
    q = torch.randn(B, S, H, D, device="cuda").bfloat16()
    k = torch.randn(B, S, Hk, D, device="cuda").bfloat16()
    v = torch.randn(B, S, Hk, Dv, device="cuda").bfloat16()
    scale = 1.0 / math.sqrt(D)
    o, lse = fmha_fwd(q, k, v, causal=causal, scale=scale)
    ro, rlse = _ref(q, k, v, causal, scale)
    torch.testing.assert_close(o.float(), ro, atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(lse, rlse, atol=2e-2, rtol=1e-3)


def test_fmha_packed_qkv_view():
    from simumax_b200.ops.attention import fmha_fwd
    torch.manual_seed(1)
    B, S, H, Hk, D = 1, 512, 8, 2, 128
    qkv = torch.randn(B, S, (H + 2 * Hk) * D, device="cuda").bfloat16()
    q = qkv[..., : H * D].view(B, S, H, D)
    k = qkv[..., H * D: (H + Hk) * D].view(B, S, Hk, D)
    v = qkv[..., (H + Hk) * D:].view(B, S, Hk, D)
    o, _ = fmha_fwd(q, k, v, causal=True)
    ro, _ = _ref(q, k, v, True, 1.0 / math.sqrt(D))
    torch.testing.assert_close(o.float(), ro, atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("cfg", [(1, 256, 2, 2, 128), (2, 512, 8, 2, 128), (1, 200, 4, 1, 128), (1, 1024, 4, 4, 128)])
def test_fmha_bwd(cfg, causal):
    from simumax_b200.ops.attention import fmha_bwd, fmha_fwd
    torch.manual_seed(3)
    B, S, H, Hk, D = cfg
    qkv = torch.randn(B, S, (H + 2 * Hk) * D, device="cuda").bfloat16()
    q = qkv[..., : H * D].view(B, S, H, D)
    k = qkv[..., H * D: (H + Hk) * D].view(B, S, Hk, D)
    v = qkv[..., (H + Hk) * D:].view(B, S, Hk, D)
    d_o = torch.randn(B, S, H, D, device="cuda").bfloat16()
    scale = 1.0 / math.sqrt(D)
    o, lse = fmha_fwd(q, k, v, causal=causal, scale=scale)
    d_qkv = torch.full_like(qkv, float("nan"))
    dq = d_qkv[..., : H * D].view(B, S, H, D)
    dk = d_qkv[..., H * D: (H + Hk) * D].view(B, S, Hk, D)
    dv = d_qkv[..., (H + Hk) * D:].view(B, S, Hk, D)
    fmha_bwd(d_o, q, k, v, o, lse, dq, dk, dv, causal=causal, scale=scale)
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    ro, _ = _ref(qf, kf, vf, causal, scale)
    ro.backward(d_o.float())
    for name, got, ref in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        err = (got.float() - ref).norm() / ref.norm()
        assert torch.isfinite(got.float()).all(), name
        assert err < 2e-2, (name, err.item())
        torch.testing.assert_close(got.float(), ref, atol=5e-2 * ref.abs().max().item(), rtol=0)


@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("cfg", [(1, 256, 2, 2), (2, 512, 4, 4), (1, 200, 4, 2), (1, 1024, 8, 8)])
def test_fmha_bwd_mla(cfg, causal):
    """MLA head dims (qk 192 / v 128): the backward variant without S/dP look-ahead (dQ^T aliases their TMEM columns)."""
    from simumax_b200.ops.attention import fmha_bwd, fmha_fwd
    torch.manual_seed(5)
    B, S, H, Hk = cfg
    q = torch.randn(B, S, H, 192, device="cuda").bfloat16()
    k = torch.randn(B, S, Hk, 192, device="cuda").bfloat16()
    v = torch.randn(B, S, Hk, 128, device="cuda").bfloat16()
    d_o = torch.randn(B, S, H, 128, device="cuda").bfloat16()
    scale = 1.0 / math.sqrt(192)
    o, lse = fmha_fwd(q, k, v, causal=causal, scale=scale)
    dq, dk, dv = fmha_bwd(d_o, q, k, v, o, lse, causal=causal, scale=scale)
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    ro, _ = _ref(qf, kf, vf, causal, scale)
    ro.backward(d_o.float())
    for name, got, ref in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        assert torch.isfinite(got.float()).all(), name
        err = (got.float() - ref).norm() / ref.norm()
        assert err < 2e-2, (name, err.item())


@pytest.mark.parametrize("cluster", ["1", "2"])
def test_fmha_fwd2_cluster_and_cross_lengths(cluster, monkeypatch):
    """Second-generation forward with and without the 2-CTA K / V multicast cluster, and with S_q != S_kv (bottom-right causal)."""
    from simumax_b200.ops.attention import fmha_fwd
    monkeypatch.setenv("SMX_FMHA_CLUSTER", cluster)
    torch.manual_seed(7)
    for (B, Sq, Skv, H, Hk) in [(1, 768, 768, 8, 2), (1, 256, 1024, 4, 2), (2, 384, 640, 4, 4 if cluster == "1" else 2)]:
        q = torch.randn(B, Sq, H, 128, device="cuda").bfloat16()
        k = torch.randn(B, Skv, Hk, 128, device="cuda").bfloat16()
        v = torch.randn(B, Skv, Hk, 128, device="cuda").bfloat16()
        for causal in (True, False):
            o, lse = fmha_fwd(q, k, v, causal=causal)
            ro, rlse = _ref(q, k, v, causal, 1.0 / math.sqrt(128))
            torch.testing.assert_close(o.float(), ro, atol=2e-2, rtol=2e-2)
            torch.testing.assert_close(lse, rlse, atol=2e-2, rtol=1e-3)
