"""Where does the lm_head gradient error of tests/test_trainer_gpu.py come from?  Splits it into: cross-entropy kernel (dlogits from
the SAME bf16 logits in fp32), the NT wgrad GEMM (same bf16 operands in fp32), and the bf16 forward itself."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from simumax_b200.ops import elementwise as E
from simumax_b200.ops.gemm import gemm
from simumax_b200.runtime.trainer import LlamaDims, LlamaTrainer

torch.manual_seed(0)
dims = LlamaDims(hidden=256, heads=2, kv_heads=1, head_dim=128, ffn=512, layers=2, vocab=512)
tr = LlamaTrainer(dims, micro_batch=2, seq_len=128, max_grad_norm=0.0)
for k_ in ("embed", "lm_head"):
    tr.w[k_].mul_(4.0)
for k_ in ("wqkv", "wo", "wfc1", "wfc2"):
    for t_ in tr.w[k_]:
        t_.mul_(4.0)
g = torch.Generator(device="cuda").manual_seed(7)
tokens = torch.randint(0, dims.vocab, (tr.T,), device="cuda", generator=g)
targets = torch.randint(0, dims.vocab, (tr.T,), device="cuda", generator=g)
rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())
# forward up to the logits with our kernels
tr.forward(tokens, targets, loss_scale=1.0 / tr.T)
dl = tr.logits.clone()                      # bf16 dlogits
xf = tr.act["xf"].clone()
# (1) CE kernel alone: recompute logits in bf16 the same way, then fp32 softmax of those bf16 logits
lg = torch.empty_like(tr.logits)
gemm(xf, tr.w["lm_head"], "TN", out=lg)
p = torch.softmax(lg.float(), -1)
p[torch.arange(tr.T), targets] -= 1.0
ref_dl = p / tr.T
print("dlogits (kernel, bf16) vs fp32 softmax of the same bf16 logits:", rel(dl, ref_dl))
print("   of which bf16 rounding of the fp32 result alone:", rel(ref_dl.bfloat16(), ref_dl))
# (2) wgrad GEMM alone
gw = torch.zeros(dims.vocab, dims.hidden, device="cuda", dtype=torch.float32)
gemm(dl, xf, "NT", out=gw, accumulate=True)
print("NT wgrad GEMM vs fp32 matmul of the same bf16 operands:", rel(gw, dl.float().t() @ xf.float()))
# (3) logits: bf16 GEMM vs fp32 product of the same operands
print("logits bf16 GEMM vs fp32 product:", rel(lg, xf.float() @ tr.w["lm_head"].float().t()))
# (4) everything in fp32 from our xf: isolates the forward-activation error from the head
lg32 = xf.float() @ tr.w["lm_head"].float().t()
p32 = torch.softmax(lg32, -1); p32[torch.arange(tr.T), targets] -= 1.0
print("lm_head grad: ours vs fp32 head on OUR xf:", rel(gw, (p32 / tr.T).t() @ xf.float()))
print("   logits magnitude: std", float(lg32.std()), "max", float(lg32.abs().max()), " softmax max prob mean", float(torch.softmax(lg32, -1).max(-1).values.mean()))
