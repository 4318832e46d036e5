"""Measured side of "predicted vs measured": a Megatron-style Llama training step built from this repo's sm_100a
kernels (tcgen05 GEMMs for every linear incl. the fp32-accumulating wgrad, tcgen05 flash attention, fused
RMSNorm / SwiGLU / RoPE / cross-entropy / AdamW kernels) with explicit forward/backward (no autograd), static
activation storage and a ZeRO-1 distributed optimizer over data-parallel ranks.

Semantics mirror what the analytical model assumes (reference docs/b200/b200_real_repro_guide.md:263-278):
bf16 params + fp32 main-grad accumulation fused into the wgrad GEMM, fp32 master/m/v, no activation recompute,
sequence-first flattened tokens ``[B*S, H]``, untied LM head, fused CE writing dlogits in place.

Public API:  ``trainer = LlamaTrainer(model_cfg, micro_batch, seq_len, dp_group=None)``
             ``loss = trainer.train_step(list_of_host_batches)``   (tokens/targets int64 on pinned host memory)
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from ..ops import elementwise as E
from ..ops.attention import fmha_fwd
from ..ops.gemm import gemm


@dataclass
class LlamaDims:
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
    def from_model_config(cls, m, vocab_multiple: int = 128):
        vocab = int(math.ceil(m.vocab_size / vocab_multiple) * vocab_multiple)
        return cls(m.hidden_size, m.head_num, m.kv_head_num, m.head_size, m.intermediate_size, m.layer_num, vocab)

    @property
    def qkv_out(self):
        return (self.heads + 2 * self.kv_heads) * self.head_dim

    def layer_numel(self):
        H, F = self.hidden, self.ffn
        return H + self.qkv_out * H + H * self.heads * self.head_dim + H + 2 * F * H + H * F

    def total_numel(self):
        return self.vocab * self.hidden + self.layers * self.layer_numel() + self.hidden + self.vocab * self.hidden


class _Flat:
    """Carves named views out of one flat buffer (Megatron-style contiguous param / grad buffers)."""

    def __init__(self, numel, dtype, device, buf=None):
        self.buf = torch.empty(numel, dtype=dtype, device=device) if buf is None else buf
        assert self.buf.numel() == numel and self.buf.dtype == dtype
        self.off = 0

    def align(self, multiple: int) -> int:
        """Skip to the next multiple (bucket boundary); returns the new offset."""
        self.off = (self.off + multiple - 1) // multiple * multiple
        return self.off

    def take(self, *shape):
        n = 1
        for s in shape:
            n *= s
        v = self.buf[self.off:self.off + n].view(*shape)
        self.off += n
        return v


class LlamaTrainer:
    def __init__(self, dims: LlamaDims, micro_batch: int, seq_len: int, dp_group: Optional[dist.ProcessGroup] = None,
                 lr: float = 1e-4, weight_decay: float = 0.1, betas=(0.9, 0.95), max_grad_norm: float = 1.0,
                 attention_bwd: str = "auto", seed: int = 1234, device=None, comm: str = "auto",
                 first_stage: bool = True, last_stage: bool = True, in_flight: int = 1, recompute: str = "none",
                 recompute_layers: Optional[int] = None, recompute_unused_outputs: bool = True, data_parallel: bool = True,
                 cp_group: Optional[dist.ProcessGroup] = None):
        """``first_stage`` / ``last_stage`` / ``in_flight``: pipeline-stage mode (runtime/pp_trainer.py) -- the stage owns
        ``dims.layers`` layers, the embedding only if first, final norm + LM head + loss only if last, and keeps ``in_flight``
        sets of activations (1F1B holds several micro-batches between their forward and backward).

        ``recompute`` (reference config.py:469-575 semantics for a dense block, first ``recompute_layers`` layers):
        ``"selective"`` = the reference's ``selective_recompute`` with ``attn_recompute`` + ``mlp_recompute``: the QKV linear, RoPE and
        core attention, and FC1 / SwiGLU / FC2 are re-run in the backward pass, so ``qkv``, ``lse``, ``u`` and ``g`` are not kept;
        ``"full"`` = ``full_block``: only the layer input is kept and the whole layer forward is re-run.
        ``recompute_unused_outputs``: re-run every op of a recomputed segment, as Megatron's checkpointing does (its last GEMM's output
        is thrown away); False skips those.  ``data_parallel=False``: no ZeRO sharding / gradient reduction even when a process group
        exists (pipeline stages pass it; their data-parallel groups are wired by the pipeline trainer).

        ``cp_group``: context parallelism, all-to-all flavour (DeepSpeed-Ulysses; reference cost model dense_module.py:1158-1232,
        1259-1338): every rank of the group holds ``seq_len / cp`` consecutive tokens of each sequence for everything but the core
        attention; around it the packed QKV goes through an all-to-all (this repo's in-kernel NVLink all-to-all) that trades the
        sequence shard for a head shard -- full sequence, ``heads / cp`` heads -- and the attention output goes back the same way.
        Weights are replicated over the group, so ``dp_group`` (the ZeRO-1 group) must span the cp ranks as well."""
        assert recompute in ("none", "selective", "full"), recompute
        self.first_stage, self.last_stage = first_stage, last_stage
        self.d = dims
        self.cp_group = cp_group
        self.cp = dist.get_world_size(cp_group) if cp_group is not None else 1
        self.cp_rank = dist.get_rank(cp_group) if cp_group is not None else 0
        assert seq_len % self.cp == 0 and dims.heads % self.cp == 0 and dims.kv_heads % self.cp == 0
        assert self.cp == 1 or micro_batch == 1, "context parallelism is implemented for micro-batch size 1"
        self.S_full = seq_len
        self.B, self.S = micro_batch, seq_len // self.cp            # S: tokens of a sequence this rank holds
        self.T = micro_batch * self.S
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.group = dp_group
        use_dp = data_parallel and dist.is_initialized()
        self.world = dist.get_world_size(dp_group) if use_dp else 1
        self.rank = dist.get_rank(dp_group) if use_dp else 0
        self.recompute = recompute
        self.rl = 0 if recompute == "none" else min(dims.layers, dims.layers if recompute_layers is None else recompute_layers)
        self.recompute_unused = recompute_unused_outputs
        self._unsaved = {"none": set(), "selective": {"qkv", "lse", "u", "g", "o_full"},
                         "full": {"h1", "rstd1", "qkv", "lse", "o", "x2", "h2", "rstd2", "u", "g", "o_full"}}[recompute]
        self.lr, self.wd, self.betas, self.max_norm = lr, weight_decay, betas, max_grad_norm
        self.step_count = 0
        d, dev = dims, self.device
        # ZeRO-1 buckets: embedding | one per layer | final norm + LM head.  Every bucket is padded to a multiple of
        # world * 1024 elements so each rank owns an equal, 16-byte aligned slice of every bucket; a bucket's grads can
        # be reduce-scattered as soon as its layer's backward is done (overlap with the rest of the backward pass).
        self._balign = self.world * 1024
        up = lambda n: (n + self._balign - 1) // self._balign * self._balign
        sizes = ([up(d.vocab * d.hidden) if first_stage else 0] + [up(d.layer_numel())] * d.layers
                 + [up(d.hidden + d.vocab * d.hidden) if last_stage else 0])
        self.buckets = []                      # (flat offset, padded size)
        off = 0
        for sz in sizes:
            self.buckets.append((off, sz))
            off += sz
        self.P, self.Ppad = off, off
        self.shard_off = [b_lo // self.world for b_lo, _ in self.buckets]      # offsets inside the local shard arrays
        # ---- parameters (bf16) and fp32 main grads, identical layouts.  With data parallelism both live in one
        # symmetric (peer-mapped) allocation so the ZeRO-1 reduce-scatter / all-gather run as our own NVLink kernels.
        self.shard = self.Ppad // self.world
        self.ws, self.comm = None, "none"
        p_buf = g_buf = None
        if self.world > 1:
            self.comm = "nccl" if comm == "nccl" else "smx"
            if self.comm == "smx":
                try:
                    from ..parallel.symm import SymmetricWorkspace
                    self._off_p, self._off_g = 0, self.Ppad * 2
                    self._off_sc = self._off_g + self.Ppad * 4
                    self.ws = SymmetricWorkspace(self._off_sc + 512, 256, group=dp_group)
                    p_buf = self.ws.view(self._off_p, torch.bfloat16, self.Ppad)
                    g_buf = self.ws.view(self._off_g, torch.float32, self.Ppad)
                except Exception as e:                      # no peer access / symmetric memory: NCCL baseline path
                    if comm == "smx":
                        raise
                    import warnings
                    warnings.warn(f"in-kernel collectives unavailable ({type(e).__name__}: {e}); using NCCL")
                    self.ws, self.comm, p_buf, g_buf = None, "nccl", None, None
        self.params = _Flat(self.Ppad, torch.bfloat16, dev, p_buf)
        self.grads = _Flat(self.Ppad, torch.float32, dev, g_buf)
        self.w, self.g = self._carve(self.params), self._carve(self.grads)
        gen = torch.Generator(device=dev).manual_seed(seed)
        self.params.buf.normal_(0.0, 0.02, generator=gen)
        for name in ("ln1", "ln2"):
            for t in self.w[name]:
                t.fill_(1.0)
        if last_stage:
            self.w["final_norm"].fill_(1.0)
        self.grads.buf.zero_()
        # ---- ZeRO-1 optimizer shard = concatenation of this rank's slice of every bucket
        self.master = torch.cat([self._my_slice(self.params.buf, b) for b in range(len(self.buckets))]).float()
        self.m = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        if self.ws is not None:
            self.sqnorm = self.ws.view(self._off_sc, torch.float32, 4)[:1]       # all-reduced in place (16-byte granule)
            self.sqnorm_out = self.ws.view(self._off_sc + 256, torch.float32, 4)[:1]
        else:
            self.sqnorm = torch.zeros(1, dtype=torch.float32, device=dev)
            self.sqnorm_out = self.sqnorm
        # ---- static activation storage
        T, H, F, L = self.T, d.hidden, d.ffn, d.layers
        bf = dict(dtype=torch.bfloat16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        shapes = {"h1": ((T, H), bf), "qkv": ((T, d.qkv_out), bf), "o": ((T, d.heads * d.head_dim), bf), "x2": ((T, H), bf),
                  "h2": ((T, H), bf), "u": ((T, 2 * F), bf), "g": ((T, F), bf), "rstd1": ((T,), f32), "rstd2": ((T,), f32),
                  "lse": ((self.B, d.heads // self.cp, self.S_full), f32)}
        if self.cp > 1:      # full sequence x this rank's head group: what the core attention saw (its backward needs it)
            shapes["o_full"] = ((self.S_full, d.heads // self.cp * d.head_dim), bf)
        # tensors a recomputed layer does not keep share ONE scratch entry (index 0); layers past ``rl`` keep their own
        n_of = lambda name: (L - self.rl + 1) if name in self._unsaved else L
        self._scratch = {name: torch.empty(*shp, **kw) for name, (shp, kw) in shapes.items() if name in self._unsaved}
        def new_act():
            a = {"x": torch.empty(L + 1, T, H, **bf), "xf": torch.empty(T, H, **bf), "rstdf": torch.empty(T, **f32)}
            for name, (shp, kw) in shapes.items():
                a[name] = torch.empty(n_of(name) - (1 if name in self._unsaved else 0), *shp, **kw)
            return a
        self.act_slots = [new_act() for _ in range(max(1, in_flight))]
        self.act = self.act_slots[0]
        self.o_rc = torch.empty(T, d.heads * d.head_dim, **bf) if recompute == "selective" else None
        self.logits = torch.empty(T, d.vocab, **bf) if last_stage else None
        # backward scratch (reused by every layer)
        self.dx = torch.empty(T, H, **bf)
        self.dx_b = torch.empty(T, H, **bf)
        self.d_big = torch.empty(T, 2 * F, **bf)
        self.d_f = torch.empty(T, F, **bf)
        self.d_qkv = torch.empty(T, d.qkv_out, **bf)
        self.d_o = torch.empty(T, d.heads * d.head_dim, **bf)
        self.tmp_h = torch.empty(T, H, **bf)
        self.cos, self.sin = E.rope_tables(self.S_full, d.head_dim, d.rope_theta, device=dev)
        self.cp_ws = None
        if self.cp > 1:
            from ..parallel.symm import SymmetricWorkspace
            nb = T * d.qkv_out * 2
            self.cp_ws = SymmetricWorkspace(nb, nb, group=cp_group, device=dev)
            self.qkv_loc = torch.empty(T, d.qkv_out, **bf)
        self.tok_dev = torch.empty(T, dtype=torch.int64, device=dev)
        self.tgt_dev = torch.empty(T, dtype=torch.int64, device=dev)
        self.loss_rows = None
        self.loss_acc = torch.zeros(1, dtype=torch.float32, device=dev)
        self.attention_bwd = self._pick_attention_bwd(attention_bwd)
        self.debug_hook = None       # callable(layer_index, trainer) invoked after each layer's backward
        self.comm_stream = torch.cuda.Stream(device=dev) if self.world > 1 else None
        self._ag_events = [None] * len(self.buckets)      # params of bucket b are valid on the main stream after this
        self._rs_launched = [False] * len(self.buckets)
        self.overlap = self.world > 1

    def _my_slice(self, flat_buf: torch.Tensor, b: int) -> torch.Tensor:
        lo, sz = self.buckets[b]
        per = sz // self.world
        return flat_buf[lo + self.rank * per: lo + (self.rank + 1) * per]

    def _shard_slice(self, arr: torch.Tensor, b: int) -> torch.Tensor:
        per = self.buckets[b][1] // self.world
        return arr[self.shard_off[b]: self.shard_off[b] + per]

    # ------------------------------------------------------------------------------------------
    def _carve(self, flat: _Flat):
        d = self.d
        H, F = d.hidden, d.ffn
        out = {"embed": flat.take(d.vocab if self.first_stage else 0, H), "ln1": [], "wqkv": [], "wo": [], "ln2": [], "wfc1": [],
               "wfc2": []}
        for _ in range(d.layers):
            flat.align(self._balign)
            out["ln1"].append(flat.take(H))
            out["wqkv"].append(flat.take(d.qkv_out, H))
            out["wo"].append(flat.take(H, d.heads * d.head_dim))
            out["ln2"].append(flat.take(H))
            out["wfc1"].append(flat.take(2 * F, H))
            out["wfc2"].append(flat.take(H, F))
        flat.align(self._balign)
        out["final_norm"] = flat.take(H if self.last_stage else 0)
        out["lm_head"] = flat.take(d.vocab if self.last_stage else 0, H)
        return out

    def _pick_attention_bwd(self, choice: str) -> str:
        if choice != "auto":
            return choice
        try:
            from ..ops.attention import fmha_bwd  # noqa: F401
            return "smx"
        except Exception:
            pass
        try:
            import flash_attn  # noqa: F401
            return "flash_attn"
        except Exception:
            return "sdpa_recompute"

    def symmetric_workspaces(self):
        return [w for w in (self.ws, self.cp_ws) if w is not None]

    def memory_bytes(self) -> dict:
        act = sum(t.numel() * t.element_size() for slot in self.act_slots for t in slot.values())
        act += sum(t.numel() * t.element_size() for t in self._scratch.values())
        return {"params_bf16": self.params.buf.numel() * 2, "main_grads_fp32": self.grads.buf.numel() * 4,
                "optimizer_fp32": self.shard * 12, "activations": act, "logits": self.logits.numel() * 2 if self.logits is not None else 0,
                "torch_max_allocated": torch.cuda.max_memory_allocated(self.device)}

    # ------------------------------------------------------------------------------------------
    def _attention_views(self, qkv):
        d = self.d
        B, S = self.B, self.S
        nq, nk, hd = d.heads, d.kv_heads, d.head_dim
        t = qkv.view(B, S, d.qkv_out)
        q = t[..., : nq * hd].view(B, S, nq, hd)
        k = t[..., nq * hd: (nq + nk) * hd].view(B, S, nk, hd)
        v = t[..., (nq + nk) * hd:].view(B, S, nk, hd)
        return q, k, v

    def use_slot(self, k: int) -> None:
        """Select the activation set the next forward / backward works on (pipeline stages keep several in flight)."""
        self.act = self.act_slots[k % len(self.act_slots)]

    def _A(self, name: str, i: int) -> torch.Tensor:
        """Layer ``i``'s tensor ``name``: its own slice if the layer keeps it, the shared scratch entry if it is recomputed."""
        if name in self._unsaved:
            return self._scratch[name] if i < self.rl else self.act[name][i - self.rl]
        return self.act[name][i]

    def _attn_segment_cp(self, i: int, o_out: torch.Tensor) -> None:
        """Context-parallel attention segment: QKV linear on the local tokens -> all-to-all (sequence shard -> head shard) -> RoPE at
        the global positions -> core attention over the full sequence with heads / cp heads -> all-to-all back."""
        d, w, A, ws, cp = self.d, self.w, self._A, self.cp_ws, self.cp
        Tl, S, hd = self.T, self.S_full, d.head_dim
        nq, nk = d.heads // cp, d.kv_heads // cp
        wq, wo = d.qkv_out // cp, nq * hd
        bf16 = torch.bfloat16
        gemm(A("h1", i), w["wqkv"][i], "TN", out=self.qkv_loc)        # columns grouped per destination rank: [q_r | k_r | v_r] x cp
        E.transpose01_(ws.input(bf16, Tl * d.qkv_out).view(cp, Tl, wq), self.qkv_loc.view(Tl, cp, wq))
        ws.all_to_all(Tl * wq * 2)
        qkv = A("qkv", i).view(S, wq)
        E.copy_(qkv, ws.output(bf16, S * wq).view(S, wq))
        E.rope_inplace(qkv, self.cos, self.sin, S, nq + nk, hd)
        t = qkv.view(1, S, wq)
        q, k, v = t[..., :nq * hd].view(1, S, nq, hd), t[..., nq * hd:(nq + nk) * hd].view(1, S, nk, hd), t[..., (nq + nk) * hd:].view(1, S, nk, hd)
        o_full = A("o_full", i)
        fmha_fwd(q, k, v, causal=True, out=o_full.view(1, S, nq, hd), lse=A("lse", i))
        E.copy_(ws.input(bf16, S * wo).view(S, wo), o_full)            # rows [s * Tl, (s + 1) * Tl) go to rank s
        ws.all_to_all(Tl * wo * 2)
        E.transpose01_(o_out.view(Tl, cp, wo), ws.output(bf16, S * wo).view(cp, Tl, wo))

    def _attention_backward_cp(self, i: int) -> None:
        d, A, ws, cp = self.d, self._A, self.cp_ws, self.cp
        Tl, S, hd = self.T, self.S_full, d.head_dim
        nq, nk = d.heads // cp, d.kv_heads // cp
        wq, wo = d.qkv_out // cp, nq * hd
        bf16 = torch.bfloat16
        E.transpose01_(ws.input(bf16, Tl * cp * wo).view(cp, Tl, wo), self.d_o.view(Tl, cp, wo))
        ws.all_to_all(Tl * wo * 2)
        do_full = ws.output(bf16, S * wo).view(1, S, nq, hd)
        t = A("qkv", i).view(1, S, wq)
        q, k, v = t[..., :nq * hd].view(1, S, nq, hd), t[..., nq * hd:(nq + nk) * hd].view(1, S, nk, hd), t[..., (nq + nk) * hd:].view(1, S, nk, hd)
        dt_flat = ws.input(bf16, S * wq).view(S, wq)                   # gradient of the head-sharded packed QKV, written in place
        dt = dt_flat.view(1, S, wq)
        dq, dk, dv = dt[..., :nq * hd].view(1, S, nq, hd), dt[..., nq * hd:(nq + nk) * hd].view(1, S, nk, hd), dt[..., (nq + nk) * hd:].view(1, S, nk, hd)
        from ..ops.attention import fmha_bwd
        fmha_bwd(do_full, q, k, v, A("o_full", i).view(1, S, nq, hd), A("lse", i), dq, dk, dv, causal=True, scale=1.0 / math.sqrt(hd))
        E.rope_inplace(dt_flat, self.cos, self.sin, S, nq + nk, hd, inverse=True)
        ws.all_to_all(Tl * wq * 2)
        E.transpose01_(self.d_qkv.view(Tl, cp, wq), ws.output(bf16, S * wq).view(cp, Tl, wq))

    def _attn_segment(self, i: int, o_out: torch.Tensor) -> None:
        """QKV linear + RoPE + core attention of layer ``i`` (the reference's ``attn_recompute`` segment)."""
        if self.cp > 1:
            return self._attn_segment_cp(i, o_out)
        d, w, A = self.d, self.w, self._A
        qkv = A("qkv", i)
        gemm(A("h1", i), w["wqkv"][i], "TN", out=qkv)
        E.rope_inplace(qkv, self.cos, self.sin, self.S, d.heads + d.kv_heads, d.head_dim)
        q, k, v = self._attention_views(qkv)
        fmha_fwd(q, k, v, causal=True, out=o_out.view(self.B, self.S, d.heads, d.head_dim), lse=A("lse", i))

    def _mlp_segment(self, i: int, with_fc2: bool = True) -> None:
        """FC1 + SwiGLU (+ FC2 into ``tmp_h``) of layer ``i`` (the reference's ``mlp_recompute`` segment)."""
        w, A = self.w, self._A
        gemm(A("h2", i), w["wfc1"][i], "TN", out=A("u", i))
        E.swiglu_fwd(A("u", i), out=A("g", i))
        if with_fc2:
            gemm(A("g", i), w["wfc2"][i], "TN", out=self.tmp_h)

    def _layer_forward(self, i: int, recomputing: bool = False) -> None:
        """Layer ``i`` from its input ``x[i]`` (and ``h1`` = its first norm, produced by whoever produced ``x[i]``)."""
        d, a, w, A = self.d, self.act, self.w, self._A
        L = d.layers
        x = a["x"][i]
        if recomputing:                                   # full-block recompute starts from the layer input alone
            E.rmsnorm_fwd(x, w["ln1"][i], d.eps, out=A("h1", i), rstd=A("rstd1", i))
        self._attn_segment(i, A("o", i))
        gemm(A("o", i), w["wo"][i], "TN", out=self.tmp_h)
        E.add_rmsnorm_fwd(x, self.tmp_h, w["ln2"][i], d.eps, out_sum=A("x2", i), out_normed=A("h2", i), rstd=A("rstd2", i))
        self._mlp_segment(i, with_fc2=(not recomputing) or self.recompute_unused)
        if recomputing:
            return
        if i + 1 < L:                                     # residual add fused with the next layer's input norm
            E.add_rmsnorm_fwd(A("x2", i), self.tmp_h, w["ln1"][i + 1], d.eps, out_sum=a["x"][i + 1], out_normed=A("h1", i + 1),
                              rstd=A("rstd1", i + 1))
        elif self.last_stage:
            E.add_rmsnorm_fwd(A("x2", i), self.tmp_h, w["final_norm"], d.eps, out_sum=a["x"][L], out_normed=a["xf"], rstd=a["rstdf"])
        else:
            E.add(A("x2", i), self.tmp_h, out=a["x"][L])

    def forward(self, tokens: torch.Tensor, targets: torch.Tensor, loss_scale: float) -> torch.Tensor:
        """First stage: ``tokens`` are token ids; later stages: the [T, H] hidden states received from the previous stage.
        Last stage: returns the per-row loss (logits then hold dlogits); earlier stages: the hidden states to send on."""
        d, a, w = self.d, self.act, self.w
        L = d.layers
        self._wait_params(0)
        if self.first_stage:
            E.embedding_fwd(tokens, w["embed"], a["x"][0])
        elif tokens.data_ptr() != a["x"][0].data_ptr():
            E.copy_(a["x"][0], tokens)
        self._wait_params(1)
        E.rmsnorm_fwd(a["x"][0], w["ln1"][0], d.eps, out=self._A("h1", 0), rstd=self._A("rstd1", 0))
        for i in range(L):
            self._wait_params(1 + i)
            if i + 1 < L:
                self._wait_params(2 + i)                  # the fused add + norm at the end of layer i reads ln1 of layer i + 1
            else:
                self._wait_params(L + 1)
            self._layer_forward(i)
        if not self.last_stage:
            return a["x"][L]
        gemm(a["xf"], w["lm_head"], "TN", out=self.logits)
        loss_rows = E.cross_entropy_(self.logits, targets, grad_scale=loss_scale)      # logits now hold dlogits
        return loss_rows

    def backward(self, tokens: torch.Tensor, last_microbatch: bool = False, d_out: Optional[torch.Tensor] = None):
        """``last_microbatch``: the gradients are final after this pass, so each bucket's reduce-scatter is launched on
        the communication stream the moment its layer is done.  Pipeline stages other than the last pass ``d_out`` (the
        gradient of the hidden states they sent on); stages other than the first return the gradient to send back."""
        d, a, w, g, A = self.d, self.act, self.w, self.g, self._A
        L = d.layers
        rs = self._launch_rs if (last_microbatch and self.overlap) else (lambda b: None)
        if self.last_stage:
            dl = self.logits
            gemm(dl, a["xf"], "NT", out=g["lm_head"], accumulate=True)
            gemm(dl, w["lm_head"], "NN", out=self.tmp_h)
            E.rmsnorm_bwd(a["x"][L], w["final_norm"], self.tmp_h, a["rstdf"], g["final_norm"], dx=self.dx)
        elif d_out.data_ptr() != self.dx.data_ptr():
            E.copy_(self.dx, d_out)
        rs(L + 1)
        for i in reversed(range(L)):
            dx3 = self.dx
            rc = self.recompute if i < self.rl else "none"
            if rc == "full":
                self._layer_forward(i, recomputing=True)
            elif rc == "selective":
                self._mlp_segment(i, with_fc2=self.recompute_unused)
            # MLP
            gemm(dx3, A("g", i), "NT", out=g["wfc2"][i], accumulate=True)
            gemm(dx3, w["wfc2"][i], "NN", out=self.d_f)
            E.swiglu_bwd(A("u", i), self.d_f, out=self.d_big)
            gemm(self.d_big, A("h2", i), "NT", out=g["wfc1"][i], accumulate=True)
            gemm(self.d_big, w["wfc1"][i], "NN", out=self.tmp_h)
            E.rmsnorm_bwd(A("x2", i), w["ln2"][i], self.tmp_h, A("rstd2", i), g["ln2"][i], dx=dx3, accumulate_dx=True)
            # attention
            if rc == "selective":
                self._attn_segment(i, self.o_rc)          # re-derives qkv (post-RoPE) and lse; the kept ``o`` is identical
            gemm(dx3, A("o", i), "NT", out=g["wo"][i], accumulate=True)
            gemm(dx3, w["wo"][i], "NN", out=self.d_o)
            if self.cp > 1:
                self._attention_backward_cp(i)                   # includes both all-to-alls and the inverse RoPE
            else:
                self._attention_backward(i)
                E.rope_inplace(self.d_qkv, self.cos, self.sin, self.S, d.heads + d.kv_heads, d.head_dim, inverse=True)
            gemm(self.d_qkv, A("h1", i), "NT", out=g["wqkv"][i], accumulate=True)
            gemm(self.d_qkv, w["wqkv"][i], "NN", out=self.tmp_h)
            E.rmsnorm_bwd(a["x"][i], w["ln1"][i], self.tmp_h, A("rstd1", i), g["ln1"][i], dx=dx3, accumulate_dx=True)
            if self.debug_hook is not None:
                self.debug_hook(i, self)
            rs(1 + i)
        if self.first_stage:
            E.embedding_bwd(tokens, self.dx, g["embed"])
        rs(0)
        return None if self.first_stage else self.dx

    def _attention_backward(self, i: int) -> None:
        d, a = self.d, self.act
        q, k, v = self._attention_views(self._A("qkv", i))
        dq, dk, dv = self._attention_views(self.d_qkv)
        o4 = self._A("o", i).view(self.B, self.S, d.heads, d.head_dim)
        do4 = self.d_o.view(self.B, self.S, d.heads, d.head_dim)
        scale = 1.0 / math.sqrt(d.head_dim)
        if self.attention_bwd == "smx":
            from ..ops.attention import fmha_bwd
            fmha_bwd(do4, q, k, v, o4, self._A("lse", i), dq, dk, dv, causal=True, scale=scale)
        elif self.attention_bwd == "flash_attn":
            from flash_attn.flash_attn_interface import _flash_attn_backward
            _flash_attn_backward(do4, q, k, v, o4, self._A("lse", i), dq, dk, dv, 0.0, scale, True, -1, -1, 0.0, None, False)
        else:
            qa, ka, va = (t.detach().permute(0, 2, 1, 3).requires_grad_(True) for t in (q, k, v))
            out = torch.nn.functional.scaled_dot_product_attention(qa, ka, va, is_causal=True, enable_gqa=True)
            out.backward(do4.permute(0, 2, 1, 3))
            dq.copy_(qa.grad.permute(0, 2, 1, 3)); dk.copy_(ka.grad.permute(0, 2, 1, 3)); dv.copy_(va.grad.permute(0, 2, 1, 3))

    # ------------------------------------------------------------------------------------------
    def _wait_params(self, b: int) -> None:
        ev = self._ag_events[b]
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)

    def _launch_rs(self, b: int) -> None:
        """Reduce-scatter bucket ``b`` (fp32 sum) into this rank's grad shard on the communication stream, then clear it."""
        if self.world == 1 or self._rs_launched[b]:
            return
        self._rs_launched[b] = True
        lo, sz = self.buckets[b]
        per = sz // self.world
        self.comm_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.comm_stream):
            if self.ws is not None:
                # pull-reduce my slice of every peer's main grads over NVLink (multimem.ld_reduce when NVLS is up)
                # (in place: the sum lands in this rank's own slice of the bucket, as Megatron's distributed optimizer does)
                self.ws.reduce_scatter(per * 4, in_off=self._off_g + lo * 4, out_off=self._off_g + (lo + self.rank * per) * 4,
                                       dtype=torch.float32, ctas=64)
            else:
                dist.reduce_scatter_tensor(self._my_slice(self.grads.buf, b), self.grads.buf[lo:lo + sz],
                                           op=dist.ReduceOp.SUM, group=self.group)
            mine = lo + self.rank * per                 # everything but the reduced slice is done for this step
            self.grads.buf[lo:mine].zero_()
            self.grads.buf[mine + per:lo + sz].zero_()

    def optimizer_step(self) -> None:
        """ZeRO-1: (bucketed, overlapped) reduce-scatter of the fp32 main grads, clip by global norm, fused AdamW on the
        local shard, bucketed all-gather of the updated bf16 params that the next forward waits on layer by layer."""
        self.step_count += 1
        b1, b2 = self.betas
        main = torch.cuda.current_stream(self.device)
        nb = len(self.buckets)
        if self.world > 1:
            for b in reversed(range(nb)):
                self._launch_rs(b)                       # whatever the backward pass did not launch already
            main.wait_stream(self.comm_stream)
            self._rs_launched = [False] * nb
        gsh = self.grads.buf
        self.sqnorm.zero_()
        if self.world > 1:                               # this rank's reduced slice of every bucket (the kernel accumulates into out)
            for b in range(nb):
                E.sqnorm(self._my_slice(gsh, b), out=self.sqnorm)
        else:
            E.sqnorm(gsh, out=self.sqnorm)
        if self.ws is not None:
            self.ws.view(self._off_sc, torch.float32, 4)[1:].zero_()
            self.ws.all_reduce(16, in_off=self._off_sc, out_off=self._off_sc + 256, dtype=torch.float32, ctas=1)
        elif self.world > 1:
            dist.all_reduce(self.sqnorm, group=self.group)
        if self.world > 1 and getattr(self, "norm_group", None) is not None:     # PP x DP: the clip norm spans all stages too
            dist.all_reduce(self.sqnorm_out, group=self.norm_group)
        # grads hold the SUM over ranks of per-rank means -> prescale by 1/world inside the fused optimizer kernel
        if self.world == 1:
            if getattr(self, "norm_group", None) is not None:       # pipeline stages: the clip norm spans all stages
                dist.all_reduce(self.sqnorm_out, group=self.norm_group)
            E.adamw_(self.master, self.m, self.v, gsh, self.params.buf, self.lr, b1, b2, 1e-8, self.wd, self.step_count,
                     sqnorm_t=self.sqnorm_out, max_norm=self.max_norm, zero_grad=True)
            return
        for b in range(nb):                              # forward order: the first layers' params are gathered first
            lo, sz = self.buckets[b]
            per = sz // self.world
            E.adamw_(self._shard_slice(self.master, b), self._shard_slice(self.m, b), self._shard_slice(self.v, b),
                     self._my_slice(gsh, b), self._my_slice(self.params.buf, b), self.lr, b1, b2, 1e-8, self.wd,
                     self.step_count, sqnorm_t=self.sqnorm_out, max_norm=self.max_norm, grad_prescale=1.0 / self.world,
                     zero_grad=True)
            self.comm_stream.wait_stream(main)
            with torch.cuda.stream(self.comm_stream):
                if self.ws is not None:
                    # push my updated bf16 slice into every peer's parameter buffer (multimem.st when NVLS is up)
                    self.ws.all_gather(per * 2, in_off=self._off_p + (lo + self.rank * per) * 2, out_off=self._off_p + lo * 2,
                                       ctas=32)
                else:
                    dist.all_gather_into_tensor(self.params.buf[lo:lo + sz], self._my_slice(self.params.buf, b),
                                                group=self.group)
                ev = torch.cuda.Event()
                ev.record(self.comm_stream)
                self._ag_events[b] = ev

    def synchronize(self) -> None:
        """Wait for outstanding parameter all-gathers (call before reading ``params`` from the host)."""
        if self.comm_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
        torch.cuda.synchronize(self.device)

    def run_microbatches(self, device_batches: List[Tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
        n = len(device_batches)
        self.loss_acc.zero_()
        scale = 1.0 / (n * self.T)
        for i, (tok, tgt) in enumerate(device_batches):
            rows = self.forward(tok, tgt, loss_scale=scale)
            self.loss_acc += rows.sum() * scale
            self.backward(tok, last_microbatch=(i == n - 1))
        self.optimizer_step()
        return self.loss_acc

    def train_step(self, host_batches: List[Tuple[torch.Tensor, torch.Tensor]]) -> float:
        """Public end-to-end step: pinned-host token/target tensors in, python float loss out (H2D + D2H included)."""
        n = len(host_batches)
        self.loss_acc.zero_()
        scale = 1.0 / (n * self.T)
        for i, (tok_h, tgt_h) in enumerate(host_batches):
            self.tok_dev.copy_(tok_h.view(-1), non_blocking=True)
            self.tgt_dev.copy_(tgt_h.view(-1), non_blocking=True)
            rows = self.forward(self.tok_dev, self.tgt_dev, loss_scale=scale)
            self.loss_acc += rows.sum() * scale
            self.backward(self.tok_dev, last_microbatch=(i == n - 1))
        self.optimizer_step()
        return float(self.loss_acc.item())
