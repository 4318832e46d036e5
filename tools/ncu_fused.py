"""Single-process target for `ncu --set full`: the CTA-pair GEMM with the TMA-store reduce-scatter epilogue (world = 1, so every
tile goes to this rank's own slot -- the epilogue path is identical, only the destination is local) followed by the fused
slot-sum + residual + RMSNorm consumer, next to the plain GEMM of the same shape.
    ncu --set full --clock-control none --import-source on -k regex:'gemm_2cta|rs_consume' -c 6 -o gpurun_out/prof_fused python tools/ncu_fused.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from simumax_b200.ops.gemm import gemm
from simumax_b200.parallel.fused import gemm_reduce_scatter_launch
from simumax_b200.parallel.symm import SymmetricWorkspace

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29790")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (4096, 8192, 7168)
x = torch.randn(M, K, device="cuda").bfloat16() * 0.5
w = torch.randn(N, K, device="cuda").bfloat16() * 0.5
ws = SymmetricWorkspace(1024, 2 * M * N * 2 + 1024)
resid = torch.randn(M, N, device="cuda").bfloat16()
gamma = torch.ones(N, device="cuda").bfloat16()
out, normed = torch.empty_like(resid), torch.empty_like(resid)
rstd = torch.empty(M, device="cuda")
full = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    gemm_reduce_scatter_launch(x, w, ws, "TN").consume(residual=resid, out=out, gamma=gamma, normed=normed, rstd=rstd)
    gemm(x, w, "TN", out=full)
torch.cuda.synchronize()
ref = full.float() + resid.float()
print("ok rel_err", ((out.float() - ref).abs().max() / ref.abs().max()).item())
dist.destroy_process_group()
