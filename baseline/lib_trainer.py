"""Library-kernel training runtime for the REFERENCE arm of bench.py (``--impl reference``).

The reference (MooreThreads/SimuMax) is an analytical model: its published accuracy numbers compare ``PerfLLM`` predictions with a
Megatron-LM + TransformerEngine run, i.e. with the vendor libraries it calibrates its system config on (cuBLAS GEMMs, cuDNN fused
attention, NCCL collectives -- simu_tools/efficency_test/*).  Megatron / TE are not installed in this image, so the measured side of the
reference arm is this file: the same Megatron-style training step (TP + SP, ZeRO-1 distributed optimizer, bf16 params, fp32 main grads,
AdamW, no recompute) written in stock PyTorch on those same libraries.  NOTHING from ``simumax_b200`` is imported here and no custom
CUDA is loaded: ``torch.matmul`` (cuBLAS), ``F.scaled_dot_product_attention`` (cuDNN / flash backends), ``F.rms_norm``,
``torch.optim.AdamW(fused=True)``, ``torch.distributed`` (NCCL).

API (mirrors the product's trainers so bench.py drives both arms the same way):
    tr = LibLlamaTrainer(dims, micro_batch, seq_len, tp_group=None, dp_group=None)
    loss = tr.train_step(list_of_pinned_host_batches)       # python float
    loss_t = tr.run_microbatches(list_of_device_batches)    # device tensor
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as F


@dataclass
class Dims:
    hidden: int
    heads: int
    kv_heads: int
    head_dim: int
    ffn: int
    layers: int
    vocab: int
    rope_theta: float = 500000.0
    eps: float = 1e-5

    @classmethod
    def from_json(cls, path: str, layers: Optional[int] = None, vocab_multiple: int = 128) -> "Dims":
        m = json.load(open(path))
        vocab = int(math.ceil(m["vocab_size"] / vocab_multiple) * vocab_multiple)
        return cls(m["hidden_size"], m["head_num"], m["kv_head_num"], m["head_size"], m["intermediate_size"],
                   layers or m["layer_num"], vocab)


# ---------------------------------------------------------------------------------------------- tensor-parallel mappings
class _GatherSeq(torch.autograd.Function):
    """Sequence-parallel all-gather in front of a column-parallel linear; backward = reduce-scatter (Megatron mappings)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        w = dist.get_world_size(group)
        out = torch.empty((x.shape[0] * w,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, dy):
        w = dist.get_world_size(ctx.group)
        out = torch.empty((dy.shape[0] // w,) + tuple(dy.shape[1:]), dtype=dy.dtype, device=dy.device)
        dist.reduce_scatter_tensor(out, dy.contiguous(), group=ctx.group)
        return out, None


class _ScatterSeq(torch.autograd.Function):
    """Reduce-scatter behind a row-parallel linear; backward = all-gather."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        w = dist.get_world_size(group)
        out = torch.empty((x.shape[0] // w,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.reduce_scatter_tensor(out, x.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, dy):
        w = dist.get_world_size(ctx.group)
        out = torch.empty((dy.shape[0] * w,) + tuple(dy.shape[1:]), dtype=dy.dtype, device=dy.device)
        dist.all_gather_into_tensor(out, dy.contiguous(), group=ctx.group)
        return out, None


_ADDMM_F32 = None      # does torch.mm(..., out_dtype=fp32) exist on this build?


def _wgrad_into(main_grad: torch.Tensor, dy: torch.Tensor, x: torch.Tensor) -> None:
    """main_grad[N, K] (fp32) += dy[T, N]^T @ x[T, K]  (bf16 operands; fp32 result where cuBLAS offers it)."""
    global _ADDMM_F32
    if _ADDMM_F32 is None:
        try:
            torch.mm(dy[:8].t(), x[:8], out_dtype=torch.float32)
            _ADDMM_F32 = True
        except Exception:
            _ADDMM_F32 = False
    if _ADDMM_F32:
        main_grad.add_(torch.mm(dy.t(), x, out_dtype=torch.float32))
    else:
        main_grad.add_(dy.t() @ x)


class _Linear(torch.autograd.Function):
    """y = x @ w^T with the weight gradient accumulated straight into the fp32 main-grad buffer (no bf16 .grad kept)."""

    @staticmethod
    def forward(ctx, x, w, main_grad):
        ctx.save_for_backward(x, w)
        ctx.mg = main_grad
        return x @ w.t()

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dy @ w
        _wgrad_into(ctx.mg, dy, x)
        return dx, None, None


class _Embedding(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, mask, w, main_grad):
        ctx.save_for_backward(idx, mask)
        ctx.mg = main_grad
        out = w.index_select(0, idx)
        if mask is not None:
            out = out * mask.unsqueeze(1).to(out.dtype)
        return out

    @staticmethod
    def backward(ctx, dy):
        idx, mask = ctx.saved_tensors
        d = dy.float()
        if mask is not None:
            d = d * mask.unsqueeze(1).float()
        ctx.mg.index_add_(0, idx, d)
        return None, None, None, None


class _SwiGLU(torch.autograd.Function):
    """silu(gate) * up on u = [gate | up]; saves only u (what Megatron's fused bias-swiglu keeps)."""

    @staticmethod
    def forward(ctx, u):
        ctx.save_for_backward(u)
        g, up = u.chunk(2, dim=-1)
        return F.silu(g) * up

    @staticmethod
    def backward(ctx, dy):
        (u,) = ctx.saved_tensors
        g, up = u.float().chunk(2, dim=-1)
        s = torch.sigmoid(g)
        d = dy.float()
        dg = d * up * (s * (1 + g * (1 - s)))
        du = d * (g * s)
        return torch.cat([dg, du], dim=-1).to(u.dtype)


def _rot(t):
    a, b = t.chunk(2, dim=-1)
    return torch.cat([-b, a], dim=-1)


class _Rope(torch.autograd.Function):
    """Rotary embedding on [B, S, h, hd]; stateless backward (inverse rotation), like a fused RoPE kernel."""

    @staticmethod
    def forward(ctx, t, cos, sin):
        ctx.save_for_backward(cos, sin)
        return t * cos + _rot(t) * sin

    @staticmethod
    def backward(ctx, d):
        cos, sin = ctx.saved_tensors
        return d * cos - _rot(d) * sin, None, None


class _VocabParallelCE(torch.autograd.Function):
    """Megatron's vocab-parallel cross entropy: max / sum-exp / target-logit all-reduces over the TP group, fp32 math."""

    @staticmethod
    def forward(ctx, logits, target, v0, group):
        lg = logits.float()
        mx = lg.max(dim=-1).values
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        lg = lg - mx.unsqueeze(1)
        local = target - v0
        mine = (local >= 0) & (local < lg.shape[1])
        idx = local.clamp(0, lg.shape[1] - 1)
        pred = lg.gather(1, idx.unsqueeze(1)).squeeze(1) * mine.float()
        dist.all_reduce(pred, group=group)
        ex = lg.exp_()
        se = ex.sum(dim=-1)
        dist.all_reduce(se, group=group)
        loss = se.log() - pred
        ex.div_(se.unsqueeze(1))
        ctx.save_for_backward(ex, idx, mine)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        sm, idx, mine = ctx.saved_tensors
        sm.scatter_add_(1, idx.unsqueeze(1), -mine.float().unsqueeze(1))
        sm.mul_(dloss.unsqueeze(1))
        return sm.to(torch.bfloat16), None, None, None


# ---------------------------------------------------------------------------------------------- trainer
class LibLlamaTrainer:
    def __init__(self, dims: Dims, micro_batch: int, seq_len: int, tp_group=None, dp_group=None, lr: float = 1e-4,
                 weight_decay: float = 0.1, betas=(0.9, 0.95), max_grad_norm: float = 1.0, seed: int = 1234, device=None):
        self.d = d = dims
        self.device = dev = device or torch.device("cuda", torch.cuda.current_device())
        self.tp_group, self.dp_group = tp_group, dp_group
        self.tp = dist.get_world_size(tp_group) if tp_group is not None else 1
        self.tp_rank = dist.get_rank(tp_group) if tp_group is not None else 0
        self.dp = dist.get_world_size(dp_group) if dp_group is not None else 1
        self.dp_rank = dist.get_rank(dp_group) if dp_group is not None else 0
        W = self.tp
        self.B, self.S, self.T = micro_batch, seq_len, micro_batch * seq_len
        self.Ts = self.T // W
        self.nq, self.nk, self.hd = d.heads // W, d.kv_heads // W, d.head_dim
        self.qkvW, self.oW, self.FW, self.VW = (self.nq + 2 * self.nk) * self.hd, self.nq * self.hd, d.ffn // W, d.vocab // W
        self.v0 = self.tp_rank * self.VW
        self.max_norm = max_grad_norm
        H, L = d.hidden, d.layers
        shapes = [("embed", (self.VW, H))]
        for i in range(L):
            shapes += [(f"wqkv{i}", (self.qkvW, H)), (f"wo{i}", (H, self.oW)), (f"wfc1{i}", (2 * self.FW, H)), (f"wfc2{i}", (H, self.FW))]
        shapes += [("lm_head", (self.VW, H))]
        n_mat = sum(a * b for _, (a, b) in shapes)
        n_norm = (2 * L + 1) * H
        pad = 1024 * self.dp
        self.P = (n_mat + n_norm + pad - 1) // pad * pad
        self.n_mat, self.n_norm = n_mat, n_norm
        self.params = torch.empty(self.P, dtype=torch.bfloat16, device=dev)
        self.main_grad = torch.zeros(self.P, dtype=torch.float32, device=dev)
        gen = torch.Generator(device=dev).manual_seed(seed + 7919 * self.tp_rank)
        self.params.normal_(0.0, 0.02, generator=gen)
        self.w, self.g = {}, {}
        off = 0
        for name, (a, b) in shapes:
            self.w[name] = self.params[off:off + a * b].view(a, b)
            self.g[name] = self.main_grad[off:off + a * b].view(a, b)
            off += a * b
        self.params[n_mat:n_mat + n_norm].fill_(1.0)
        self.norm_w = self.params[n_mat:n_mat + n_norm].view(2 * L + 1, H).detach().requires_grad_(True)   # autograd leaf (bf16 .grad)
        self.norm_g = self.main_grad[n_mat:n_mat + n_norm].view(2 * L + 1, H)
        # ZeRO-1: equal contiguous slices of the flat buffers over the data-parallel replicas
        self.zshard = self.P // self.dp
        self.zlo = self.dp_rank * self.zshard
        self.master = self.params[self.zlo:self.zlo + self.zshard].float().requires_grad_(True)
        self.grad_shard = torch.empty(self.zshard, dtype=torch.float32, device=dev) if self.dp > 1 else self.main_grad
        self.opt = torch.optim.AdamW([self.master], lr=lr, betas=betas, eps=1e-8, weight_decay=weight_decay, fused=True)
        inv = 1.0 / (d.rope_theta ** (torch.arange(0, d.head_dim, 2, device=dev, dtype=torch.float32) / d.head_dim))
        ang = torch.outer(torch.arange(seq_len, device=dev, dtype=torch.float32), inv)
        ang = torch.cat([ang, ang], dim=-1)
        self.cos = ang.cos().to(torch.bfloat16).view(1, seq_len, 1, d.head_dim)
        self.sin = ang.sin().to(torch.bfloat16).view(1, seq_len, 1, d.head_dim)
        self.tok_dev = torch.empty(self.T, dtype=torch.int64, device=dev)
        self.tgt_dev = torch.empty(self.T, dtype=torch.int64, device=dev)
        self.loss_acc = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_count = 0
        from torch.nn.attention import SDPBackend
        self._sdp_backends = [SDPBackend.CUDNN_ATTENTION, SDPBackend.FLASH_ATTENTION, SDPBackend.EFFICIENT_ATTENTION]

    # ------------------------------------------------------------------------------------------
    def _attention(self, qkv):
        from torch.nn.attention import sdpa_kernel
        B, S, nq, nk, hd = self.B, self.S, self.nq, self.nk, self.hd
        t = qkv.view(B, S, self.qkvW)
        q = t[..., : nq * hd].view(B, S, nq, hd)
        k = t[..., nq * hd:(nq + nk) * hd].view(B, S, nk, hd)
        v = t[..., (nq + nk) * hd:].view(B, S, nk, hd)
        q = _Rope.apply(q, self.cos, self.sin)
        k = _Rope.apply(k, self.cos, self.sin)
        if qkv.is_cuda:
            with sdpa_kernel(self._sdp_backends, set_priority=True):
                o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True,
                                                   enable_gqa=nq != nk)
        else:
            o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True,
                                               enable_gqa=nq != nk)
        return o.transpose(1, 2).reshape(B * S, nq * hd)

    def _loss(self, tokens, targets):
        d, w, g, H = self.d, self.w, self.g, self.d.hidden
        L, tp = d.layers, self.tp
        nw = self.norm_w
        if tp > 1:
            local = tokens - self.v0
            mine = (local >= 0) & (local < self.VW)
            x = _Embedding.apply(local.clamp(0, self.VW - 1), mine, w["embed"], g["embed"])
            x = _ScatterSeq.apply(x, self.tp_group)
        else:
            x = _Embedding.apply(tokens, None, w["embed"], g["embed"])
        gather = (lambda t: _GatherSeq.apply(t, self.tp_group)) if tp > 1 else (lambda t: t)
        scatter = (lambda t: _ScatterSeq.apply(t, self.tp_group)) if tp > 1 else (lambda t: t)
        for i in range(L):
            h = F.rms_norm(x, (H,), nw[2 * i], d.eps)
            qkv = _Linear.apply(gather(h), w[f"wqkv{i}"], g[f"wqkv{i}"])
            o = self._attention(qkv)
            x = x + scatter(_Linear.apply(o, w[f"wo{i}"], g[f"wo{i}"]))
            h = F.rms_norm(x, (H,), nw[2 * i + 1], d.eps)
            u = _Linear.apply(gather(h), w[f"wfc1{i}"], g[f"wfc1{i}"])
            x = x + scatter(_Linear.apply(_SwiGLU.apply(u), w[f"wfc2{i}"], g[f"wfc2{i}"]))
        h = F.rms_norm(x, (H,), nw[2 * L], d.eps)
        logits = _Linear.apply(gather(h), w["lm_head"], g["lm_head"])
        if tp > 1:
            return _VocabParallelCE.apply(logits, targets, self.v0, self.tp_group).sum()
        return F.cross_entropy(logits.float(), targets, reduction="sum")

    def _microbatch(self, tokens, targets, scale):
        loss = self._loss(tokens, targets) * scale
        loss.backward()
        self.loss_acc += loss.detach()

    def optimizer_step(self) -> None:
        self.step_count += 1
        ng = self.norm_w.grad
        if ng is not None:
            if self.tp > 1:                     # sequence-parallel norm weights saw only this rank's rows
                dist.all_reduce(ng, group=self.tp_group)
            self.norm_g.add_(ng)
            self.norm_w.grad = None
        if self.dp > 1:
            dist.reduce_scatter_tensor(self.grad_shard, self.main_grad, group=self.dp_group)
        gs = self.grad_shard
        sq = torch.zeros(1, dtype=torch.float32, device=self.device)
        lo = self.zlo
        cut = min(max(self.n_mat - lo, 0), self.zshard)
        sq += gs[:cut].square().sum() if cut > 0 else 0.0
        if cut < self.zshard and self.tp_rank == 0:
            sq += gs[cut:].square().sum()
        if self.tp > 1:
            dist.all_reduce(sq, group=self.tp_group)
        if self.dp > 1:
            dist.all_reduce(sq, group=self.dp_group)
        norm = sq.sqrt() / self.dp
        coef = torch.clamp(self.max_norm / (norm + 1e-6), max=1.0) / self.dp
        self.master.grad = gs.mul_(coef)
        self.opt.step()
        self.master.grad = None
        self.params[self.zlo:self.zlo + self.zshard].copy_(self.master.detach())
        if self.dp > 1:
            dist.all_gather_into_tensor(self.params, self.params[self.zlo:self.zlo + self.zshard].clone(), group=self.dp_group)
        self.main_grad.zero_()

    def run_microbatches(self, device_batches: List[Tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
        n = len(device_batches)
        self.loss_acc.zero_()
        scale = 1.0 / (n * self.T)
        for tok, tgt in device_batches:
            self._microbatch(tok, tgt, scale)
        self.optimizer_step()
        return self.loss_acc

    def train_step(self, host_batches) -> float:
        n = len(host_batches)
        self.loss_acc.zero_()
        scale = 1.0 / (n * self.T)
        for tok_h, tgt_h in host_batches:
            self.tok_dev.copy_(tok_h.view(-1), non_blocking=True)
            self.tgt_dev.copy_(tgt_h.view(-1), non_blocking=True)
            self._microbatch(self.tok_dev, self.tgt_dev, scale)
        self.optimizer_step()
        return float(self.loss_acc.item())

    def memory_bytes(self) -> dict:
        return {"params_bf16": self.P * 2, "main_grads_fp32": self.P * 4, "optimizer_fp32": self.zshard * 12,
                "torch_max_allocated": torch.cuda.max_memory_allocated(self.device)}
