"""A few launches of every kernel family that had no ncu capture yet (grouped GEMM, MXFP8 GEMM, fused CE, permute / unpermute, RMSNorm
fwd / bwd, fused add + RMSNorm, SwiGLU, AdamW, embedding, MLA helpers) -- the target of `ncu --set full -k regex:...` (1 GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from simumax_b200.ops import elementwise as E
from simumax_b200.ops.gemm import gemm, gemm_mxfp8, mxfp8_quantize

which = sys.argv[1] if len(sys.argv) > 1 else "all"
R = lambda *s: torch.randn(*s, device="cuda").bfloat16()
T, H, V = 8192, 4096, 128256
if which in ("all", "grouped"):
    x, w = R(20, 1232, 5120), R(20, 3072, 5120)
    dy = R(20, 1232, 3072)
    gw = torch.zeros(20, 3072, 5120, device="cuda")
    for _ in range(3):
        gemm(x, w, "TN"); gemm(dy, w, "NN"); gemm(dy, x, "NT", out=gw, accumulate=True)
if which == "wgrad":           # fp32 main-grad accumulation: SMX_GEMM_RED=0 (read-modify-write epilogue) vs 1 (red.global.add.v4.f32)
    dy, x = R(8192, 28672), R(8192, 4096)
    gw = torch.zeros(28672, 4096, device="cuda")
    for _ in range(3):
        gemm(dy, x, "NT", out=gw, accumulate=True)
if which in ("all", "mxfp8"):
    a, b = R(8192, 8192), R(8192, 8192)
    (qa, sa), (qb, sb) = mxfp8_quantize(a), mxfp8_quantize(b)
    for _ in range(3):
        gemm_mxfp8(qa, sa, qb, sb)
if which in ("all", "elementwise"):
    x, d = R(T, H), R(T, H)
    gamma = torch.ones(H, device="cuda", dtype=torch.bfloat16)
    y, y2 = torch.empty_like(x), torch.empty_like(x)
    rstd = torch.empty(T, device="cuda")
    dg = torch.zeros(H, device="cuda")
    u = R(T, 2 * 14336)
    g = torch.empty(T, 14336, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        E.rmsnorm_fwd(x, gamma, out=y, rstd=rstd)
        E.add_rmsnorm_fwd(x, d, gamma, 1e-5, out_sum=y, out_normed=y2, rstd=rstd)
        E.rmsnorm_bwd(x, gamma, d, rstd, dg, dx=y2)
        E.swiglu_fwd(u, out=g)
        E.swiglu_bwd(u, g, out=u.clone())
        E.add(x, d, out=y)
if which in ("all", "ce"):
    lg = R(4096, V)
    tgt = torch.randint(0, V, (4096,), device="cuda")
    for _ in range(3):
        E.cross_entropy_(lg, tgt, grad_scale=1.0 / 4096)
if which in ("all", "permute"):
    x = R(4096, 5120)
    row_map = torch.randint(0, 4096, (4096 * 6,), device="cuda", dtype=torch.int32)
    inv = torch.randperm(4096 * 6, device="cuda", dtype=torch.int32).view(4096, 6).contiguous()
    probs = torch.rand(4096, 6, device="cuda")
    for _ in range(3):
        p = E.permute_rows(x, row_map)
        E.unpermute_combine(p, inv, probs, 4096, 6)
if which in ("all", "optimizer"):
    n = 256 << 20
    master, m, v, grad = (torch.zeros(n, device="cuda") for _ in range(4))
    param = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
    sq = torch.ones(1, device="cuda")
    for i in range(3):
        E.sqnorm(grad, out=sq)
        E.adamw_(master, m, v, grad, param, 1e-4, 0.9, 0.95, 1e-8, 0.1, i + 1, sqnorm_t=sq, max_norm=1.0, zero_grad=True)
if which in ("all", "embedding"):
    table = R(V, H)
    tok = torch.randint(0, V, (T,), device="cuda")
    out = torch.empty(T, H, device="cuda", dtype=torch.bfloat16)
    gtab = torch.zeros(V, H, device="cuda")
    for _ in range(3):
        E.embedding_fwd(tok, table, out)
        E.embedding_bwd(tok, out, gtab)
torch.cuda.synchronize()
print("ok")
