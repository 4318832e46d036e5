"""Expert-parallel MoE MLP block (forward + backward) on the sm_100a kernels: the measured counterpart of the analytical
model's Router / Permutation / GroupLinear / UnPermutation chain (reference cost model simumax/core/transformer/
moe_module.py:20-1566; measurement contract docs/b200/b200_real_repro_guide.md: pad-to-capacity, capacity factor 1).

Data path per micro-batch (T local tokens, E experts over ``ep`` ranks, El = E / ep local experts, capacity C per (rank, expert)):

    router GEMM (tcgen05) -> softmax / top-k (tiny, torch) -> capacity assignment (sort, drop / pad)
    fused dispatch          ONE kernel: gather token rows by the permutation map and store them straight into the owning
                            rank's expert-major receive buffer over NVLink (``moe_dispatch_kernel``); no permuted copy exists
    grouped GEMM FC1 -> SwiGLU -> grouped GEMM FC2   (tcgen05, one launch for all local experts)
    fused combine           ONE kernel: expert outputs go back to the token owners' permuted-order buffer (``moe_combine_kernel``)
    unpermute + weighted sum (local kernel)

Backward mirrors it: the dispatch kernel carries ``prob * dy`` (row scale fused into the gather), grouped dgrad / wgrad with
fp32 main-grad accumulation, combine back, scatter-add (the unpermute kernel with unit weights), router backward through
the small softmax/top-k graph.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from ..ops import elementwise as E
from ..ops.gemm import gemm
from ..parallel.symm import SymmetricWorkspace


class EPMoELayer:
    def __init__(self, hidden: int, ffn: int, n_experts: int, topk: int, tokens: int, ep_group: Optional[dist.ProcessGroup] = None,
                 capacity_factor: float = 1.0, force_balanced: bool = False, seed: int = 1234, device=None,
                 weights=None, grads=None, shared: Optional[dict] = None):
        """``weights`` = (router [E, H], w1 [El, 2F, H], w2 [El, H, F]) bf16 and ``grads`` = the matching fp32 main-grad tensors when
        the parameters live in a trainer's flat buffers (runtime/deepseek_trainer.py); otherwise the layer owns random ones.
        ``shared``: a dict the MoE layers of one model pass around -- the symmetric dispatch / combine workspace, the expert-output
        buffer and the backward scratch are transient (consumed before the next layer runs), so one set serves every layer."""
        assert dist.is_initialized()
        self.group = ep_group or dist.group.WORLD
        self.ep = ep = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        assert n_experts % ep == 0
        self.H, self.F, self.E, self.k, self.T = hidden, ffn, n_experts, topk, tokens
        self.El = n_experts // ep
        # capacity per (source rank, expert): ceil(T * k / E * factor), kept a multiple of 8 rows
        c = int(-(-tokens * topk * capacity_factor // n_experts))
        self.C = (c + 7) // 8 * 8
        self.force_balanced = force_balanced
        self.device = dev = device or torch.device("cuda", torch.cuda.current_device())
        H, F, El, C = hidden, ffn, self.El, self.C
        self.rows = ep * El * C                      # rows of the permuted order (local) == rows received (expert side)
        buf = self.rows * H * 2
        self.off_recv, self.off_back = 0, (buf + 1023) // 1024 * 1024
        shared = shared if shared is not None else {}
        key = (id(self.group), self.rows, H, F)
        pool = shared.setdefault(key, {})
        if "ws" not in pool:
            pool["ws"] = SymmetricWorkspace(self.off_back, buf + 1024, group=self.group, device=dev)
        self.ws = pool["ws"]
        f32 = dict(dtype=torch.float32, device=dev)
        if weights is not None:
            self.w_router, self.w1, self.w2 = weights
            self.g_router, self.g_w1, self.g_w2 = grads
            assert self.w1.shape == (El, 2 * F, H) and self.w2.shape == (El, H, F) and self.w_router.shape == (n_experts, H)
        else:
            gen = torch.Generator(device=dev).manual_seed(seed)                     # router replicated: same seed everywhere
            self.w_router = (torch.randn(n_experts, H, device=dev, generator=gen) * 0.02).bfloat16()
            gen_e = torch.Generator(device=dev).manual_seed(seed + 1 + self.rank)
            self.w1 = (torch.randn(El, 2 * F, H, device=dev, generator=gen_e) * 0.02).bfloat16()      # [gate ; up] per expert
            self.w2 = (torch.randn(El, H, F, device=dev, generator=gen_e) * 0.02).bfloat16()
            self.g_router = torch.zeros(n_experts, H, **f32)
            self.g_w1 = torch.zeros(El, 2 * F, H, **f32)
            self.g_w2 = torch.zeros(El, H, F, **f32)
        bf = dict(dtype=torch.bfloat16, device=dev)
        M = ep * C                                   # rows per local expert
        self.M = M
        self.xin = torch.empty(El, M, H, **bf)       # dispatched tokens, saved for wgrad
        self.u = torch.empty(El, M, 2 * F, **bf)
        self.g = torch.empty(El, M, F, **bf)
        self.yperm = torch.empty(self.rows, H, **bf)     # expert outputs back in permuted order (needed for d_probs)
        for name, shp in (("yexp", (El, M, H)), ("d_u", (El, M, 2 * F)), ("d_g", (El, M, F)), ("d_in", (El, M, H))):
            if name not in pool:
                pool[name] = torch.empty(*shp, **bf)
            setattr(self, name, pool[name])
        self.logits = torch.empty(tokens, n_experts, **f32)

    # ------------------------------------------------------------------------------------------
    def _route(self, x: torch.Tensor):
        """Router GEMM + softmax / top-k; returns (probs [T, k] fp32 with grad graph, idx [T, k])."""
        gemm(x, self.w_router, "TN", out=self.logits)
        self._logits_leaf = self.logits.detach().requires_grad_(True)
        scores = torch.softmax(self._logits_leaf, dim=-1)
        if self.force_balanced:
            t = torch.arange(self.T, device=x.device).unsqueeze(1)
            idx = (t * self.k + torch.arange(self.k, device=x.device).unsqueeze(0)) % self.E
            probs = scores.gather(1, idx)
        else:
            probs, idx = scores.topk(self.k, dim=-1)
        return probs, idx

    def _assign(self, probs: torch.Tensor, idx: torch.Tensor):
        """Capacity assignment: slot (t, j) -> row expert * C + position, dropped when the expert is full; padding rows -1."""
        T, k, E_, C = self.T, self.k, self.E, self.C
        flat_e = idx.reshape(-1)
        order = torch.argsort(flat_e, stable=True)
        sorted_e = flat_e[order]
        counts = torch.bincount(flat_e, minlength=E_)
        starts = torch.cumsum(counts, 0) - counts
        pos = torch.arange(T * k, device=idx.device) - starts[sorted_e]
        keep = pos < C
        rows = (sorted_e * C + pos)[keep]
        slots = order[keep]
        row_map = torch.full((E_ * C,), -1, dtype=torch.int32, device=idx.device)
        row_map[rows] = (slots // k).to(torch.int32)
        inv_map = torch.full((T * k,), -1, dtype=torch.int32, device=idx.device)
        inv_map[slots] = rows.to(torch.int32)
        prob_rows = torch.zeros(E_ * C, dtype=torch.float32, device=idx.device)
        prob_rows[rows] = probs.detach().reshape(-1)[slots]
        return row_map, inv_map, prob_rows

    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        ws, El, C, H = self.ws, self.El, self.C, self.H
        self.x = x
        self.probs, self.idx = self._route(x)
        self.row_map, self.inv_map, self.prob_rows = self._assign(self.probs, self.idx)
        ws.moe_dispatch(x, self.row_map, self.off_recv, C, El)
        recv = ws.view(self.off_recv, torch.bfloat16, self.rows * H).view(El, self.M, H)
        E.copy_(self.xin, recv)                                   # keep for the wgrad; frees the receive buffer
        gemm(self.xin, self.w1, "TN", out=self.u)
        E.swiglu_fwd(self.u.view(-1, 2 * self.F), out=self.g.view(-1, self.F))
        gemm(self.g, self.w2, "TN", out=self.yexp)
        ws.moe_combine(self.yexp.view(-1, H), self.off_back, C, El)
        E.copy_(self.yperm, ws.view(self.off_back, torch.bfloat16, self.rows * H).view(self.rows, H))
        self.probs_det = self.probs.detach().contiguous()
        return E.unpermute_combine(self.yperm, self.inv_map, self.probs_det, self.T, self.k, out=out)

    def backward(self, dy: torch.Tensor, dx: Optional[torch.Tensor] = None) -> torch.Tensor:
        ws, El, C, H = self.ws, self.El, self.C, self.H
        # combine weights: d_probs[t, j] = <dy[t], expert output of slot (t, j)>
        d_probs = E.moe_prob_grad(dy, self.yperm, self.inv_map, self.k)
        # expert side receives prob * dy (row scale fused into the dispatch gather)
        ws.moe_dispatch(dy, self.row_map, self.off_recv, C, El, row_scale=self.prob_rows)
        d_out = ws.view(self.off_recv, torch.bfloat16, self.rows * H).view(El, self.M, H)
        gemm(d_out, self.g, "NT", out=self.g_w2, accumulate=True)
        gemm(d_out, self.w2, "NN", out=self.d_g)
        E.swiglu_bwd(self.u.view(-1, 2 * self.F), self.d_g.view(-1, self.F), out=self.d_u.view(-1, 2 * self.F))
        gemm(self.d_u, self.xin, "NT", out=self.g_w1, accumulate=True)
        gemm(self.d_u, self.w1, "NN", out=self.d_in)
        ws.moe_combine(self.d_in.view(-1, H), self.off_back, C, El)
        d_perm = ws.view(self.off_back, torch.bfloat16, self.rows * H).view(self.rows, H)
        dx = E.unpermute_combine(d_perm, self.inv_map, None, self.T, self.k, out=dx)      # scatter-add over the k slots
        # router: softmax / top-k backward on the [T, E] graph, then the two router GEMMs
        (d_logits,) = torch.autograd.grad(self.probs, self._logits_leaf, d_probs)
        dl = d_logits.bfloat16()
        gemm(dl, self.x, "NT", out=self.g_router, accumulate=True)
        dxr = gemm(dl, self.w_router, "NN")
        return E.add(dx, dxr, out=dx)

    def flops(self) -> float:
        """Useful FLOPs of one forward + backward on this rank's experts (capacity-padded rows count)."""
        return 3 * 2.0 * self.El * self.M * self.H * 3 * self.F
