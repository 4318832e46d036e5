"""Tensor-parallel + sequence-parallel Llama training step on the fused compute + collective kernels.

The measured counterpart of the analytical model's TP / SP branch (reference cost model: LinearCol / LinearRow / Embedding /
ParallelCE ``_comp_leaf_intra_net_info``, simumax/core/transformer/dense_module.py:100-128, 351-416, 646-695, 2173-2215).
Megatron semantics: column-parallel QKV / FC1 / LM head, row-parallel O / FC2, vocab-parallel embedding and cross-entropy,
the residual stream and both norms sharded over the sequence ([T / tp, H] rows per rank).

Where Megatron issues a collective next to a GEMM, this runtime launches ONE kernel:

=====================================  =================================================================================
column-linear forward  AG -> GEMM      ``allgather_gemm_inkernel`` (TN): gather warps inside the CTA-pair GEMM push the
                                       local rows over NVLink while the MMAs start on them
row-linear forward     GEMM -> RS      ``gemm_reduce_scatter_launch`` (TN): TMA-store epilogue into the owners' slots;
                                       its consumer also does the residual add and the next RMSNorm
column-linear dgrad    GEMM -> RS      same kernel, NN layout; consumer sums the slots, RMSNorm backward follows
row-linear dgrad       AG -> GEMM      same kernel, NN layout, on the gathered output gradient (wgrad reuses the buffer)
=====================================  =================================================================================

The all-gather destinations of the forward pass are the per-layer saved activations themselves (they live in the symmetric
workspace), so nothing is re-gathered in the backward pass -- 2 x L x T x H bf16 is small change on a 180 GB part.

Public API mirrors ``LlamaTrainer``: ``TPLlamaTrainer(dims, micro_batch, seq_len, tp_group)``, ``train_step(host_batches)``.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from ..ops import elementwise as E
from ..ops.attention import fmha_bwd, fmha_fwd
from ..ops.gemm import gemm
from ..parallel.fused import ag_views, allgather_gemm_inkernel, gemm_reduce_scatter_launch
from ..parallel.symm import SymmetricWorkspace
from .trainer import LlamaDims, _Flat


def _up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class TPLlamaTrainer:
    def __init__(self, dims: LlamaDims, micro_batch: int, seq_len: int, tp_group: Optional[dist.ProcessGroup] = None,
                 lr: float = 1e-4, weight_decay: float = 0.1, betas=(0.9, 0.95), max_grad_norm: float = 1.0, seed: int = 1234,
                 device=None, dp_group: Optional[dist.ProcessGroup] = None, recompute: str = "none",
                 recompute_layers: Optional[int] = None, recompute_unused_outputs: bool = True, dp_comm: str = "smx"):
        """``dp_group``: data-parallel replicas of this TP group (same TP rank in every replica).  With it the optimizer is
        ZeRO-1 over the replicas the way Megatron's distributed optimizer runs in the reference's measurement contract (overlap
        features off): one reduce-scatter of the fp32 main grads, AdamW on the local shard, one all-gather of the bf16 params --
        as this repo's in-kernel NVLink collectives over a second symmetric workspace spanning the replicas (``dp_comm="smx"``) or
        through NCCL (``"nccl"``, the baseline).

        ``recompute``: as in ``LlamaTrainer`` (``"selective"`` = QKV linear + RoPE + core attention and FC1 + SwiGLU + FC2 re-run in
        the backward pass from the kept gathered norm outputs; ``"full"`` = the whole layer re-run from its input, including its
        all-gathers and reduce-scatters)."""
        assert dist.is_initialized(), "TPLlamaTrainer needs an initialised process group (one rank per GPU)"
        assert recompute in ("none", "selective", "full"), recompute
        self.d = d = dims
        self.recompute = recompute
        self.rl = 0 if recompute == "none" else min(dims.layers, dims.layers if recompute_layers is None else recompute_layers)
        self.recompute_unused = recompute_unused_outputs
        self._unsaved = {"none": set(), "selective": {"qkv", "lse", "u", "g"},
                         "full": {"qkv", "lse", "o", "x2", "rstd2", "u", "g"}}[recompute]
        self.dp_group = dp_group
        self.dp = dist.get_world_size(dp_group) if dp_group is not None else 1
        self.dp_rank = dist.get_rank(dp_group) if dp_group is not None else 0
        self.group = tp_group or dist.group.WORLD
        self.tp = W = dist.get_world_size(self.group)
        self.rank = r = dist.get_rank(self.group)
        self.B, self.S = micro_batch, seq_len
        self.T = T = micro_batch * seq_len
        assert W >= 2 and T % W == 0 and (T // W) % 128 == 0, "tokens per rank must be a multiple of 128"
        assert d.heads % W == 0 and d.kv_heads % W == 0 and d.ffn % W == 0 and d.vocab % W == 0
        self.Ts = Ts = T // W
        self.device = dev = device or torch.device("cuda", torch.cuda.current_device())
        self.lr, self.wd, self.betas, self.max_norm = lr, weight_decay, betas, max_grad_norm
        self.step_count = 0
        H, F, L = d.hidden, d.ffn, d.layers
        self.nq, self.nk, self.hd = d.heads // W, d.kv_heads // W, d.head_dim
        self.qkvW = (self.nq + 2 * self.nk) * self.hd          # local QKV width
        self.oW = self.nq * self.hd                            # local attention-output width (K shard of the O projection)
        self.FW = F // W
        self.VW = d.vocab // W
        self.v0 = r * self.VW
        # ---- parameters: [tp-sharded matrices | replicated norm weights]; fp32 main grads share the layout
        n_shard = self.VW * H + L * (self.qkvW * H + H * self.oW + 2 * self.FW * H + H * self.FW) + self.VW * H
        n_rep = (2 * L + 1) * H
        self.n_shard, self.n_rep = _up(n_shard, 1024), _up(n_rep, 1024 * self.dp)
        self.P = self.n_shard + self.n_rep
        if self.P % (1024 * self.dp):
            self.n_shard += 1024 * self.dp - self.P % (1024 * self.dp)       # equal, 4 KB aligned ZeRO-1 shards
            self.P = self.n_shard + self.n_rep
        # ---- symmetric workspace layout (bytes)
        act = T * H * 2
        off = 0
        def take(nbytes):
            nonlocal off
            o = off
            off = _up(off + nbytes, 1024)
            return o
        full_rc = self.rl if recompute == "full" else 0        # these layers re-gather into two alternating scratch buffers
        sc1, sc2 = ([take(act), take(act)], [take(act), take(act)]) if full_rc else (None, None)
        self.off_h1 = [sc1[i & 1] if i < full_rc else take(act) for i in range(L)]     # gathered ln1 outputs = saved activations
        self.off_h2 = [sc2[i & 1] if i < full_rc else take(act) for i in range(L)]
        self.off_xf = take(act)
        self.off_dy = [take(act), take(act)]                   # gathered output gradients (backward), alternating
        self.off_slots = take(2 * act)                         # two [tp][Ts, H] slot buffers of the fused reduce-scatter
        self.off_emb_in, self.off_emb_out = take(act), take(act)
        self.off_ln_in, self.off_ln_out = take(self.n_rep * 4), take(self.n_rep * 4)
        self.off_st_in, self.off_st_out = take(T * 16), take(W * T * 16)
        self.off_sc = take(1024)
        self.ws = SymmetricWorkspace(off, 1024, group=self.group, device=dev)
        ws = self.ws
        self.zshard = self.P // self.dp                                       # ZeRO-1: this replica's slice of the flat buffers
        self.zlo = self.dp_rank * self.zshard
        self.dp_ws, p_buf, g_buf = None, None, None
        self.dp_comm = "none" if self.dp == 1 else dp_comm
        if self.dp > 1 and dp_comm == "smx":
            # bf16 params | fp32 main grads | 1 KiB scalars, peer-mapped over the data-parallel replicas (the reduce-scatter lands in
            # this replica's own slice of the main-grad buffer: no separate shard buffer, as in Megatron's distributed optimizer)
            self._zp, self._zg = 0, _up(self.P * 2, 1024)
            self._zsc = self._zg + _up(self.P * 4, 1024)
            self.dp_ws = SymmetricWorkspace(self._zsc + 1024, 256, group=dp_group, device=dev)
            p_buf = self.dp_ws.view(self._zp, torch.bfloat16, self.P)
            g_buf = self.dp_ws.view(self._zg, torch.float32, self.P)
        self.params = _Flat(self.P, torch.bfloat16, dev, p_buf)
        self.grads = _Flat(self.P, torch.float32, dev, g_buf)
        self.w, self.g = self._carve(self.params), self._carve(self.grads)
        gen = torch.Generator(device=dev).manual_seed(seed + 7919 * r)       # same weights in every data-parallel replica
        self.params.buf.normal_(0.0, 0.02, generator=gen)
        self._fill_norms()
        self.grads.buf.zero_()
        self.master = self.params.buf[self.zlo:self.zlo + self.zshard].float()
        self.m = torch.zeros(self.zshard, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.zshard, dtype=torch.float32, device=dev)
        if self.dp_ws is not None:
            self.grad_shard = self.grads.buf[self.zlo:self.zlo + self.zshard]
            self.dp_sq_in = self.dp_ws.view(self._zsc, torch.float32, 4)
            self.dp_sq_out = self.dp_ws.view(self._zsc + 256, torch.float32, 4)
        else:
            self.grad_shard = self.grads.buf[self.zlo:self.zlo + self.zshard] if self.dp > 1 else None
        self.sqnorm = ws.view(self.off_sc, torch.float32, 4)
        self.sqnorm_out = ws.view(self.off_sc + 256, torch.float32, 4)
        # ---- activations
        bf = dict(dtype=torch.bfloat16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        self.h1 = [ag_views(ws, o, T, H) for o in self.off_h1]          # (full, own rows)
        self.h2 = [ag_views(ws, o, T, H) for o in self.off_h2]
        self.xf = ag_views(ws, self.off_xf, T, H)
        self.dy = [ag_views(ws, o, T, H) for o in self.off_dy]
        shapes = {"x2": ((Ts, H), bf), "rstd2": ((Ts,), f32), "qkv": ((T, self.qkvW), bf), "o": ((T, self.oW), bf),
                  "lse": ((self.B, self.nq, self.S), f32), "u": ((T, 2 * self.FW), bf), "g": ((T, self.FW), bf)}
        self._scratch = {n: torch.empty(*shp, **kw) for n, (shp, kw) in shapes.items() if n in self._unsaved}
        self.act = {"x": torch.empty(L + 1, Ts, H, **bf), "rstd1": torch.empty(L, Ts, **f32), "rstdf": torch.empty(Ts, **f32)}
        for n, (shp, kw) in shapes.items():                   # recomputed layers share one scratch entry per unsaved tensor
            self.act[n] = torch.empty((L - self.rl) if n in self._unsaved else L, *shp, **kw)
        self.o_rc = torch.empty(T, self.oW, **bf) if recompute == "selective" else None
        self.logits = torch.empty(T, self.VW, **bf)
        self.dx = torch.empty(Ts, H, **bf)                      # residual-stream gradient (sequence shard)
        self.tmp_s = torch.empty(Ts, H, **bf)
        self.d_u = torch.empty(T, 2 * self.FW, **bf)
        self.d_g = torch.empty(T, self.FW, **bf)
        self.d_qkv = torch.empty(T, self.qkvW, **bf)
        self.d_o = torch.empty(T, self.oW, **bf)
        self.cos, self.sin = E.rope_tables(self.S, d.head_dim, d.rope_theta, device=dev)
        self.tok_dev = torch.empty(T, dtype=torch.int64, device=dev)
        self.tgt_dev = torch.empty(T, dtype=torch.int64, device=dev)
        self.loss_acc = torch.zeros(1, **f32)
        self._dy_calls = 0
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)

    # ------------------------------------------------------------------------------------------
    def _carve(self, flat: _Flat):
        d, H, L = self.d, self.d.hidden, self.d.layers
        out = {"embed": flat.take(self.VW, H), "wqkv": [], "wo": [], "wfc1": [], "wfc2": []}
        for _ in range(L):
            out["wqkv"].append(flat.take(self.qkvW, H))
            out["wo"].append(flat.take(H, self.oW))
            out["wfc1"].append(flat.take(2 * self.FW, H))       # [gate shard ; up shard]
            out["wfc2"].append(flat.take(H, self.FW))
        out["lm_head"] = flat.take(self.VW, H)
        flat.off = self.n_shard
        out["norms"] = flat.buf[self.n_shard:self.n_shard + (2 * L + 1) * H]
        out["ln1"] = [flat.take(H) for _ in range(L)]
        out["ln2"] = [flat.take(H) for _ in range(L)]
        out["final_norm"] = flat.take(H)
        return out

    def _fill_norms(self):
        self.w["norms"].fill_(1.0)

    def load_full_weights(self, full: dict) -> None:
        """Slice this rank's shards out of full (unsharded) ``LlamaTrainer.w``-style weights."""
        d, W, r, hd = self.d, self.tp, self.rank, self.hd
        nq, nk, FW = self.nq, self.nk, self.FW
        self.w["embed"].copy_(full["embed"][self.v0:self.v0 + self.VW])
        self.w["lm_head"].copy_(full["lm_head"][self.v0:self.v0 + self.VW])
        for i in range(d.layers):
            wq = full["wqkv"][i]
            q0, k0, v0 = 0, d.heads * hd, (d.heads + d.kv_heads) * hd
            self.w["wqkv"][i].copy_(torch.cat([wq[q0 + r * nq * hd: q0 + (r + 1) * nq * hd],
                                               wq[k0 + r * nk * hd: k0 + (r + 1) * nk * hd],
                                               wq[v0 + r * nk * hd: v0 + (r + 1) * nk * hd]]))
            self.w["wo"][i].copy_(full["wo"][i][:, r * nq * hd:(r + 1) * nq * hd])
            f1 = full["wfc1"][i]
            self.w["wfc1"][i].copy_(torch.cat([f1[r * FW:(r + 1) * FW], f1[d.ffn + r * FW: d.ffn + (r + 1) * FW]]))
            self.w["wfc2"][i].copy_(full["wfc2"][i][:, r * FW:(r + 1) * FW])
            self.w["ln1"][i].copy_(full["ln1"][i])
            self.w["ln2"][i].copy_(full["ln2"][i])
        self.w["final_norm"].copy_(full["final_norm"])
        self.master.copy_(self.params.buf[self.zlo:self.zlo + self.zshard])

    def memory_bytes(self) -> dict:
        act = sum(t.numel() * t.element_size() for t in self.act.values()) + sum(t.numel() * t.element_size() for t in self._scratch.values())
        return {"params_bf16": self.P * 2, "main_grads_fp32": self.P * 4, "optimizer_fp32": self.zshard * 12, "activations": act,
                "symmetric_workspace": self.ws.in_bytes + self.ws.out_bytes, "logits": self.logits.numel() * 2,
                "dp_symmetric_workspace": (self.dp_ws.in_bytes + self.dp_ws.out_bytes) if self.dp_ws is not None else 0,
                "torch_max_allocated": torch.cuda.max_memory_allocated(self.device)}

    def symmetric_workspaces(self):
        return [w for w in (self.ws, self.dp_ws) if w is not None]

    def _A(self, name: str, i: int) -> torch.Tensor:
        if name in self._unsaved:
            return self._scratch[name] if i < self.rl else self.act[name][i - self.rl]
        return self.act[name][i]

    # ------------------------------------------------------------------------------------------
    def _views(self, qkv):
        B, S, nq, nk, hd = self.B, self.S, self.nq, self.nk, self.hd
        t = qkv.view(B, S, self.qkvW)
        q = t[..., : nq * hd].view(B, S, nq, hd)
        k = t[..., nq * hd: (nq + nk) * hd].view(B, S, nk, hd)
        v = t[..., (nq + nk) * hd:].view(B, S, nk, hd)
        return q, k, v

    def _ag_gemm(self, off, w, layout, out):
        return allgather_gemm_inkernel(self.ws, off, self.T, self.d.hidden, w, layout, out=out)

    def _gemm_rs(self, a, w, layout):
        return gemm_reduce_scatter_launch(a, w, self.ws, layout, slot_off=self.off_slots)

    def _attn_segment(self, i: int, o_out: torch.Tensor, gather: bool) -> None:
        """QKV linear (+ its all-gather when ``gather``) + RoPE + core attention of layer ``i``."""
        w, A = self.w, self._A
        qkv = A("qkv", i)
        if gather:
            self._ag_gemm(self.off_h1[i], w["wqkv"][i], "TN", qkv)
        else:                                   # the gathered norm output of this layer is still in its buffer
            gemm(self.h1[i][0], w["wqkv"][i], "TN", out=qkv)
        E.rope_inplace(qkv, self.cos, self.sin, self.S, self.nq + self.nk, self.hd)
        q, k, v = self._views(qkv)
        fmha_fwd(q, k, v, causal=True, out=o_out.view(self.B, self.S, self.nq, self.hd), lse=A("lse", i))

    def _mlp_segment(self, i: int, gather: bool) -> None:
        w, A = self.w, self._A
        if gather:
            self._ag_gemm(self.off_h2[i], w["wfc1"][i], "TN", A("u", i))
        else:
            gemm(self.h2[i][0], w["wfc1"][i], "TN", out=A("u", i))
        E.swiglu_fwd(A("u", i), out=A("g", i))

    def _layer_forward(self, i: int, recomputing: bool = False) -> None:
        d, a, w, A = self.d, self.act, self.w, self._A
        L = d.layers
        if recomputing:                         # full-block recompute: from the layer input alone
            E.rmsnorm_fwd(a["x"][i], w["ln1"][i], d.eps, out=self.h1[i][1], rstd=a["rstd1"][i])
        self._attn_segment(i, A("o", i), gather=True)
        self._gemm_rs(A("o", i), w["wo"][i], "TN").consume(residual=a["x"][i], out=A("x2", i), gamma=w["ln2"][i],
                                                            normed=self.h2[i][1], rstd=A("rstd2", i), eps=d.eps)
        self._mlp_segment(i, gather=True)
        if recomputing:
            if self.recompute_unused:           # Megatron's checkpoint re-runs FC2 and its reduce-scatter; the result is dropped
                self._gemm_rs(A("g", i), w["wfc2"][i], "TN").consume(out=self.tmp_s)
            return
        last = i == L - 1
        self._gemm_rs(A("g", i), w["wfc2"][i], "TN").consume(
            residual=A("x2", i), out=a["x"][i + 1], gamma=w["final_norm"] if last else w["ln1"][i + 1],
            normed=self.xf[1] if last else self.h1[i + 1][1], rstd=a["rstdf"] if last else a["rstd1"][i + 1], eps=d.eps)

    def forward(self, tokens: torch.Tensor, targets: torch.Tensor, loss_scale: float) -> torch.Tensor:
        d, a, w, ws = self.d, self.act, self.w, self.ws
        L, H, T, Ts, W = d.layers, d.hidden, self.T, self.Ts, self.tp
        ws.barrier()          # every rank is done with the previous pass: the per-layer gathered buffers may be refilled
        # vocab-parallel embedding: masked local lookup (rows of other shards = 0), reduce-scatter over the sequence
        part = ws.view(self.off_emb_in, torch.bfloat16, T * H).view(T, H)
        E.embedding_fwd(tokens, w["embed"], part, v0=self.v0)
        ws.reduce_scatter(Ts * H * 2, in_off=self.off_emb_in, out_off=self.off_emb_out)
        E.copy_(a["x"][0], ws.view(self.off_emb_out, torch.bfloat16, Ts * H).view(Ts, H))
        E.rmsnorm_fwd(a["x"][0], w["ln1"][0], d.eps, out=self.h1[0][1], rstd=a["rstd1"][0])
        for i in range(L):
            self._layer_forward(i)
        self._ag_gemm(self.off_xf, w["lm_head"], "TN", self.logits)
        # vocab-parallel cross-entropy: local statistics -> all-gather -> loss + in-place gradient
        st_in = ws.view(self.off_st_in, torch.float32, T * 4)
        E.ce_vocab_parallel_stats(self.logits, targets, self.v0, st_in)
        ws.all_gather(T * 16, in_off=self.off_st_in, out_off=self.off_st_out)
        st_all = ws.view(self.off_st_out, torch.float32, W * T * 4)
        return E.ce_vocab_parallel_grad_(self.logits, targets, self.v0, st_all, W, grad_scale=loss_scale)

    def _gather_dx(self) -> int:
        """Put the residual-stream gradient at its place in the next gathered-gradient buffer; returns its byte offset."""
        p = self._dy_calls & 1
        self._dy_calls += 1
        E.copy_(self.dy[p][1], self.dx)
        return p

    def backward(self, tokens: torch.Tensor) -> None:
        d, a, w, g, ws, A = self.d, self.act, self.w, self.g, self.ws, self._A
        L, H, T, Ts = d.layers, d.hidden, self.T, self.Ts
        dl = self.logits
        gemm(dl, self.xf[0], "NT", out=g["lm_head"], accumulate=True)
        self._gemm_rs(dl, w["lm_head"], "NN").consume(out=self.tmp_s)
        E.rmsnorm_bwd(a["x"][L], w["final_norm"], self.tmp_s, a["rstdf"], g["final_norm"], dx=self.dx)
        for i in reversed(range(L)):
            rc = self.recompute if i < self.rl else "none"
            if rc == "full":
                self._layer_forward(i, recomputing=True)
            elif rc == "selective":
                self._mlp_segment(i, gather=False)
                if self.recompute_unused:
                    self._gemm_rs(A("g", i), w["wfc2"][i], "TN").consume(out=self.tmp_s)
            # ---- MLP.  FC2 is row-parallel: its input gradient needs the gathered output gradient
            p = self._gather_dx()
            self._ag_gemm(self.off_dy[p], w["wfc2"][i], "NN", self.d_g)
            gemm(self.dy[p][0], A("g", i), "NT", out=g["wfc2"][i], accumulate=True)
            E.swiglu_bwd(A("u", i), self.d_g, out=self.d_u)
            gemm(self.d_u, self.h2[i][0], "NT", out=g["wfc1"][i], accumulate=True)
            self._gemm_rs(self.d_u, w["wfc1"][i], "NN").consume(out=self.tmp_s)
            E.rmsnorm_bwd(A("x2", i), w["ln2"][i], self.tmp_s, A("rstd2", i), g["ln2"][i], dx=self.dx, accumulate_dx=True)
            # ---- attention
            if rc == "selective":
                self._attn_segment(i, self.o_rc, gather=False)
            p = self._gather_dx()
            self._ag_gemm(self.off_dy[p], w["wo"][i], "NN", self.d_o)
            gemm(self.dy[p][0], A("o", i), "NT", out=g["wo"][i], accumulate=True)
            q, k, v = self._views(A("qkv", i))
            dq, dk, dv = self._views(self.d_qkv)
            fmha_bwd(self.d_o.view(self.B, self.S, self.nq, self.hd), q, k, v, A("o", i).view(self.B, self.S, self.nq, self.hd),
                     A("lse", i), dq, dk, dv, causal=True, scale=1.0 / math.sqrt(self.hd))
            E.rope_inplace(self.d_qkv, self.cos, self.sin, self.S, self.nq + self.nk, self.hd, inverse=True)
            gemm(self.d_qkv, self.h1[i][0], "NT", out=g["wqkv"][i], accumulate=True)
            self._gemm_rs(self.d_qkv, w["wqkv"][i], "NN").consume(out=self.tmp_s)
            E.rmsnorm_bwd(a["x"][i], w["ln1"][i], self.tmp_s, a["rstd1"][i], g["ln1"][i], dx=self.dx, accumulate_dx=True)
        # ---- vocab-parallel embedding: gather the sequence shards of dx, scatter-add the rows this rank owns
        E.copy_(ws.view(self.off_emb_out, torch.bfloat16, Ts * H).view(Ts, H), self.dx)
        ws.all_gather(Ts * H * 2, in_off=self.off_emb_out, out_off=self.off_emb_in)
        full = ws.view(self.off_emb_in, torch.bfloat16, T * H).view(T, H)
        E.embedding_bwd(tokens, full, g["embed"], v0=self.v0)

    def optimizer_step(self) -> None:
        """Norm-weight gradients are summed over the TP group (each rank saw only its sequence shard), the global gradient
        norm counts sharded matrices once per rank and the replicated norms once, then a fused AdamW on the local shard."""
        ws = self.ws
        self.step_count += 1
        b1, b2 = self.betas
        n_rep = self.n_rep
        ln_in = ws.view(self.off_ln_in, torch.float32, n_rep)
        ln_in.copy_(self.grads.buf[self.n_shard:])
        ws.all_reduce(n_rep * 4, in_off=self.off_ln_in, out_off=self.off_ln_out, dtype=torch.float32)
        rep = ws.view(self.off_ln_out, torch.float32, n_rep)
        self.grads.buf[self.n_shard:].copy_(rep)
        self.sqnorm.zero_()
        if self.dp == 1:
            E.sqnorm(self.grads.buf[:self.n_shard], out=self.sqnorm[:1])
            if self.rank == 0:
                E.sqnorm(self.grads.buf[self.n_shard:], out=self.sqnorm[:1])
            ws.all_reduce(16, in_off=self.off_sc, out_off=self.off_sc + 256, dtype=torch.float32, ctas=1)
            E.adamw_(self.master, self.m, self.v, self.grads.buf, self.params.buf, self.lr, b1, b2, 1e-8, self.wd, self.step_count,
                     sqnorm_t=self.sqnorm_out[:1], max_norm=self.max_norm, zero_grad=True)
            return
        # ---- ZeRO-1 over the data-parallel replicas (exposed -- the reference's measurement contract has overlap off)
        lo, hi = self.zlo, self.zlo + self.zshard
        dws = self.dp_ws
        if dws is not None:          # pull-reduce my slice of every replica's main grads over NVLink (multimem.ld_reduce with NVLS)
            dws.reduce_scatter(self.zshard * 4, in_off=self._zg, out_off=self._zg + lo * 4, dtype=torch.float32)       # in place
        else:
            dist.reduce_scatter_tensor(self.grad_shard, self.grads.buf, op=dist.ReduceOp.SUM, group=self.dp_group)
        cut = min(max(self.n_shard - lo, 0), self.zshard)          # [0, cut) of my shard = TP-sharded matrices, rest = replicated norms
        if cut > 0:
            E.sqnorm(self.grad_shard[:cut], out=self.sqnorm[:1])
        if cut < self.zshard and self.rank == 0:
            E.sqnorm(self.grad_shard[cut:], out=self.sqnorm[:1])
        ws.all_reduce(16, in_off=self.off_sc, out_off=self.off_sc + 256, dtype=torch.float32, ctas=1)
        if dws is not None:
            self.dp_sq_in.copy_(self.sqnorm_out)
            dws.all_reduce(16, in_off=self._zsc, out_off=self._zsc + 256, dtype=torch.float32, ctas=1)
            sq = self.dp_sq_out
        else:
            dist.all_reduce(self.sqnorm_out, group=self.dp_group)
            sq = self.sqnorm_out
        # grads hold the SUM over replicas of per-replica means -> the norm of the mean gradient is 1/dp of theirs: prescale by
        # 1/dp inside the fused optimizer kernel (which also applies it to the clipping norm)
        E.adamw_(self.master, self.m, self.v, self.grad_shard, self.params.buf[lo:hi], self.lr, b1, b2, 1e-8, self.wd,
                 self.step_count, sqnorm_t=sq[:1], max_norm=self.max_norm, grad_prescale=1.0 / self.dp, zero_grad=False)
        self.grads.buf.zero_()
        if dws is not None:          # push my updated bf16 slice into every replica's parameter buffer (multimem.st with NVLS)
            dws.all_gather(self.zshard * 2, in_off=self._zp + lo * 2, out_off=self._zp)
        else:
            dist.all_gather_into_tensor(self.params.buf, self.params.buf[lo:hi].clone(), group=self.dp_group)

    def synchronize(self) -> None:
        torch.cuda.synchronize(self.device)

    def run_microbatches(self, device_batches: List[Tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
        n = len(device_batches)
        self.loss_acc.zero_()
        scale = 1.0 / (n * self.T)
        for tok, tgt in device_batches:
            rows = self.forward(tok, tgt, loss_scale=scale)
            self.loss_acc += rows.sum() * scale
            self.backward(tok)
        self.optimizer_step()
        return self.loss_acc

    def train_step(self, host_batches: List[Tuple[torch.Tensor, torch.Tensor]]) -> float:
        """Pinned-host token / target tensors in (every TP rank gets the same batch), python float loss out."""
        n = len(host_batches)
        self.loss_acc.zero_()
        scale = 1.0 / (n * self.T)
        for tok_h, tgt_h in host_batches:
            self.tok_dev.copy_(tok_h.view(-1), non_blocking=True)
            self.tgt_dev.copy_(tgt_h.view(-1), non_blocking=True)
            rows = self.forward(self.tok_dev, self.tgt_dev, loss_scale=scale)
            self.loss_acc += rows.sum() * scale
            self.backward(self.tok_dev)
        self.optimizer_step()
        return float(self.loss_acc.item())
