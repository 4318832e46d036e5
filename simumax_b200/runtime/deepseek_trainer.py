"""DeepSeek-V2 / V3 style training step (MLA attention + dense / MoE MLP with a shared expert) on the sm_100a kernels -- the measured
counterpart of the analytical model's MLA + MoE branch (reference cost model: MLAAttention, simumax/core/transformer/
dense_module.py:2569-2853; ExpertMLP / Router / Permutation / GroupLinear, moe_module.py:20-1566) and of the published rows
``deepseekv2_tp1_ep8_pp1_dp8_l4_*`` (docs/b200/b200_release_v1.2_summary.md:25,33,41).

Layout (tp = 1): every rank holds the non-expert ("dense") parameters -- embedding, MLA projections, norms, the dense MLP of the first
``dense_layers`` layers, shared experts, routers, LM head -- data parallel over all ranks with a ZeRO-1 optimizer on this repo's
in-kernel NVLink reduce-scatter / all-gather, and E / ep routed experts per rank (expert parallel over the same ranks; with ep = dp
there is no expert data parallelism, so expert gradients are complete after the fused combine and their optimizer is local).

Per layer, forward (T local tokens):
    RMSNorm -> q_down GEMM -> RMSNorm -> q_up GEMM -> RoPE on the rotary 64 of every 192-wide head
            -> kv_down GEMM ([latent 512 | rotary key 64]) -> RoPE on the shared key slice, RMSNorm on the latent -> kv_up GEMM
               ([k_nope | v] per head) -> key assembly (k_nope + broadcast rotary slice)
            -> tcgen05 flash attention, qk head dim 192 / v head dim 128 -> output GEMM -> fused residual add + RMSNorm
    MLP:    dense layer: FC1 -> SwiGLU -> FC2;   MoE layer: shared expert (FC1 -> SwiGLU -> FC2) + routed experts (``EPMoELayer``:
            router GEMM, fused permute + all-to-all dispatch, grouped GEMMs, fused combine, weighted unpermute)
The backward pass is written out explicitly (no autograd), every weight gradient accumulates in fp32 inside the wgrad GEMM.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from ..ops import elementwise as E
from ..ops.attention import fmha_bwd, fmha_fwd
from ..ops.gemm import gemm
from .moe_layer import EPMoELayer
from .trainer import _Flat


@dataclass
class DeepSeekDims:
    hidden: int
    heads: int
    qk_nope: int            # per-head non-rotary q/k width (128)
    qk_rope: int            # rotary width (64), shared by all heads on the key side
    v_head: int             # 128
    q_lora: int             # 1536
    kv_lora: int            # 512
    ffn_dense: int          # intermediate size of the dense layers (12288)
    ffn_expert: int         # moe_ffn_hidden_size (1536)
    ffn_shared: int         # moe_shared_expert_intermediate_size (3072)
    experts: int
    topk: int
    layers: int
    dense_layers: int
    vocab: int
    rope_theta: float = 10000.0
    eps: float = 1e-6

    @classmethod
    def from_model_config(cls, m, layers: Optional[int] = None, vocab_multiple: int = 128) -> "DeepSeekDims":
        vocab = int(math.ceil(m.vocab_size / vocab_multiple) * vocab_multiple)
        return cls(m.hidden_size, m.head_num, m.qk_head_dim, m.qk_pos_emb_head_dim, m.v_head_dim, m.q_lora_rank, m.kv_lora_rank,
                   m.intermediate_size, m.moe_ffn_hidden_size, m.moe_shared_expert_intermediate_size or 0, m.expert_num, m.topk,
                   layers or m.layer_num, min(m.dense_layers, layers or m.layer_num), vocab)

    @property
    def dqk(self) -> int:
        return self.qk_nope + self.qk_rope


class DeepSeekTrainer:
    def __init__(self, dims: DeepSeekDims, micro_batch: int, seq_len: int, group: Optional[dist.ProcessGroup] = None,
                 lr: float = 1e-4, weight_decay: float = 0.1, betas=(0.9, 0.95), max_grad_norm: float = 1.0, seed: int = 1234,
                 device=None, capacity_factor: float = 1.0, force_balanced: bool = False, comm: str = "smx"):
        assert dist.is_initialized(), "DeepSeekTrainer needs an initialised process group (a single rank is fine)"
        self.d = d = dims
        self.group = group or dist.group.WORLD
        self.world = W = dist.get_world_size(self.group)          # dp = ep = world
        self.rank = dist.get_rank(self.group)
        self.B, self.S, self.T = micro_batch, seq_len, micro_batch * seq_len
        self.device = dev = device or torch.device("cuda", torch.cuda.current_device())
        self.lr, self.wd, self.betas, self.max_norm = lr, weight_decay, betas, max_grad_norm
        self.step_count = 0
        assert d.experts % W == 0
        self.El = d.experts // W
        T, H, L = self.T, d.hidden, d.layers
        nh, dqk, dv = d.heads, d.dqk, d.v_head
        # ---- dense (replicated) parameters: one flat bf16 buffer + fp32 main grads of the same layout, ZeRO-1 over the group
        self._names: List[Tuple[str, Tuple[int, ...]]] = [("embed", (d.vocab, H))]
        for i in range(L):
            self._names += [(f"ln1.{i}", (H,)), (f"wqd.{i}", (d.q_lora, H)), (f"qln.{i}", (d.q_lora,)), (f"wqu.{i}", (nh * dqk, d.q_lora)),
                            (f"wkvd.{i}", (d.kv_lora + d.qk_rope, H)), (f"kvln.{i}", (d.kv_lora,)),
                            (f"wkvu.{i}", (nh * (d.qk_nope + dv), d.kv_lora)), (f"wo.{i}", (H, nh * dv)), (f"ln2.{i}", (H,))]
            if i < d.dense_layers:
                self._names += [(f"wfc1.{i}", (2 * d.ffn_dense, H)), (f"wfc2.{i}", (H, d.ffn_dense))]
            else:
                self._names += [(f"router.{i}", (d.experts, H))]
                if d.ffn_shared:
                    self._names += [(f"ws1.{i}", (2 * d.ffn_shared, H)), (f"ws2.{i}", (H, d.ffn_shared))]
        self._names += [("final_norm", (H,)), ("lm_head", (d.vocab, H))]
        numel = lambda shp: int(torch.tensor(shp).prod())
        P = sum(numel(s) for _, s in self._names)
        self.P = (P + 1024 * W - 1) // (1024 * W) * (1024 * W)
        self.shard = self.P // W
        self.ws, p_buf, g_buf = None, None, None
        self.comm = "none" if W == 1 else comm
        if W > 1 and comm == "smx":
            from ..parallel.symm import SymmetricWorkspace
            up = lambda n: (n + 1023) // 1024 * 1024
            self._off_p, self._off_g = 0, up(self.P * 2)
            self._off_sc = self._off_g + up(self.P * 4)
            self.ws = SymmetricWorkspace(self._off_sc + 1024, 256, group=self.group, device=dev)
            p_buf = self.ws.view(self._off_p, torch.bfloat16, self.P)
            g_buf = self.ws.view(self._off_g, torch.float32, self.P)
        self.params = _Flat(self.P, torch.bfloat16, dev, p_buf)
        self.grads = _Flat(self.P, torch.float32, dev, g_buf)
        self.w = {n: self.params.take(*s) for n, s in self._names}
        self.g = {n: self.grads.take(*s) for n, s in self._names}
        gen = torch.Generator(device=dev).manual_seed(seed)          # identical dense weights on every rank
        self.params.buf.normal_(0.0, 0.02, generator=gen)
        for n, _ in self._names:
            if n.split(".")[0] in ("ln1", "ln2", "qln", "kvln", "final_norm"):
                self.w[n].fill_(1.0)
        self.grads.buf.zero_()
        lo = self.rank * self.shard
        self.master = self.params.buf[lo:lo + self.shard].float()
        self.m = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        bf = dict(dtype=torch.bfloat16, device=dev)
        if self.ws is not None:
            self.grad_shard = self.grads.buf[lo:lo + self.shard]       # the reduce-scatter lands in place in this rank's slice
            self.sq_in = self.ws.view(self._off_sc, torch.float32, 4)
            self.sq_out = self.ws.view(self._off_sc + 256, torch.float32, 4)
        else:
            self.grad_shard = self.grads.buf[lo:lo + self.shard] if W > 1 else self.grads.buf
            self.sq_in = torch.zeros(4, **f32)
            self.sq_out = self.sq_in
        # ---- routed experts (local): flat bf16 / fp32 buffers, local optimizer state
        n_moe = L - d.dense_layers
        per = self.El * (2 * d.ffn_expert * H + H * d.ffn_expert)
        self.e_params = _Flat(max(n_moe * per, 8), torch.bfloat16, dev)
        self.e_grads = _Flat(max(n_moe * per, 8), torch.float32, dev)
        gen_e = torch.Generator(device=dev).manual_seed(seed + 1 + self.rank)
        self.e_params.buf.normal_(0.0, 0.02, generator=gen_e)
        self.e_grads.buf.zero_()
        self.moe: dict = {}
        moe_shared: dict = {}           # one dispatch / combine workspace + backward scratch for all MoE layers
        for i in range(d.dense_layers, L):
            w1, w2 = self.e_params.take(self.El, 2 * d.ffn_expert, H), self.e_params.take(self.El, H, d.ffn_expert)
            g1, g2 = self.e_grads.take(self.El, 2 * d.ffn_expert, H), self.e_grads.take(self.El, H, d.ffn_expert)
            self.moe[i] = EPMoELayer(H, d.ffn_expert, d.experts, d.topk, T, ep_group=self.group, capacity_factor=capacity_factor,
                                     force_balanced=force_balanced, seed=seed + 100 * i, device=dev,
                                     weights=(self.w[f"router.{i}"], w1, w2), grads=(self.g[f"router.{i}"], g1, g2), shared=moe_shared)
        self.e_master = self.e_params.buf.float()
        self.e_m = torch.zeros_like(self.e_master)
        self.e_v = torch.zeros_like(self.e_master)
        # ---- activations (one micro-batch in flight), per layer
        def per_layer(*shape, **kw):
            return torch.empty(L, *shape, **kw)
        self.act = {
            "x": torch.empty(L + 1, T, H, **bf), "h1": per_layer(T, H, **bf), "rstd1": per_layer(T, **f32),
            "cq": per_layer(T, d.q_lora, **bf), "rstdq": per_layer(T, **f32), "cqn": per_layer(T, d.q_lora, **bf),
            "q": per_layer(T, nh * dqk, **bf), "ckv": per_layer(T, d.kv_lora + d.qk_rope, **bf), "clat": per_layer(T, d.kv_lora, **bf),
            "rstdkv": per_layer(T, **f32), "ckvn": per_layer(T, d.kv_lora, **bf), "kv": per_layer(T, nh * (d.qk_nope + dv), **bf),
            "k": per_layer(T, nh * dqk, **bf), "o": per_layer(T, nh * dv, **bf), "lse": per_layer(self.B, nh, self.S, **f32),
            "x2": per_layer(T, H, **bf), "h2": per_layer(T, H, **bf), "rstd2": per_layer(T, **f32),
            "xf": torch.empty(T, H, **bf), "rstdf": torch.empty(T, **f32),
        }
        Fd, Fs = d.ffn_dense, d.ffn_shared
        self.act_mlp = {}
        for i in range(L):
            F = Fd if i < d.dense_layers else Fs
            self.act_mlp[i] = {"u": torch.empty(T, 2 * F, **bf), "g": torch.empty(T, F, **bf)} if F else {}
        self.logits = torch.empty(T, d.vocab, **bf)
        # ---- backward scratch
        Fmax = max(Fd if d.dense_layers else 0, Fs, 8)
        self.dx = torch.empty(T, H, **bf)
        self.tmp_h = torch.empty(T, H, **bf)
        self.tmp_h2 = torch.empty(T, H, **bf)
        self.dh = torch.empty(T, H, **bf)
        self._d_u = torch.empty(T * 2 * Fmax, **bf)          # flat: every MLP flavour takes a contiguous [T, 2F] / [T, F] view
        self._d_g = torch.empty(T * Fmax, **bf)
        self.d_o = torch.empty(T, nh * dv, **bf)
        self.d_q = torch.empty(T, nh * dqk, **bf)
        self.d_k = torch.empty(T, nh * dqk, **bf)
        self.d_kv = torch.empty(T, nh * (d.qk_nope + dv), **bf)
        self.d_ckv = torch.empty(T, d.kv_lora + d.qk_rope, **bf)
        self.d_lat = torch.empty(T, d.kv_lora, **bf)
        self.d_lat_n = torch.empty(T, d.kv_lora, **bf)
        self.d_cq = torch.empty(T, d.q_lora, **bf)
        self.d_cqn = torch.empty(T, d.q_lora, **bf)
        self.cos, self.sin = E.rope_tables(self.S, d.qk_rope, d.rope_theta, device=dev)
        self.tok_dev = torch.empty(T, dtype=torch.int64, device=dev)
        self.tgt_dev = torch.empty(T, dtype=torch.int64, device=dev)
        self.loss_acc = torch.zeros(1, **f32)
        self.scale = 1.0 / math.sqrt(dqk)
        torch.cuda.synchronize(dev)
        if W > 1:
            dist.barrier(self.group)

    @classmethod
    def from_case(cls, case, seed: int = 1234, **kw) -> "DeepSeekTrainer":
        from ..configs import get_simu_model_config
        from ..core.config import ModelConfig
        mc = ModelConfig.init_from_config_file(get_simu_model_config(case.model))
        dims = DeepSeekDims.from_model_config(mc, layers=case.layers)
        assert case.tp == 1 and case.pp == 1 and case.ep == case.dp, "DeepSeekTrainer runs tp1 / pp1 with ep = dp = world"
        return cls(dims, case.mbs, case.seq, seed=seed, **kw)

    def symmetric_workspaces(self):
        seen, out = set(), []
        for w in [self.ws] + [m.ws for m in self.moe.values()]:
            if w is not None and id(w) not in seen:
                seen.add(id(w)); out.append(w)
        return out

    def total_numel(self) -> int:
        return self.P + self.e_params.buf.numel() * self.world

    def memory_bytes(self) -> dict:
        act = sum(t.numel() * t.element_size() for t in self.act.values())
        act += sum(t.numel() * t.element_size() for m in self.act_mlp.values() for t in m.values())
        moe_act = sum(t.numel() * t.element_size() for m in self.moe.values() for t in (m.xin, m.u, m.g, m.yperm))
        moe_act += sum(t.numel() * t.element_size() for m in list(self.moe.values())[:1] for t in (m.yexp, m.d_u, m.d_g, m.d_in))
        return {"dense_params_bf16": self.P * 2, "dense_main_grads_fp32": self.P * 4, "dense_optimizer_fp32": self.shard * 12,
                "expert_params_bf16": self.e_params.buf.numel() * 2, "expert_main_grads_fp32": self.e_grads.buf.numel() * 4,
                "expert_optimizer_fp32": self.e_master.numel() * 12, "activations": act, "moe_buffers": moe_act,
                "logits": self.logits.numel() * 2, "torch_max_allocated": torch.cuda.max_memory_allocated(self.device)}

    # ------------------------------------------------------------------------------------------ views
    def _q3(self, t):
        return t.view(self.T, self.d.heads, self.d.dqk)

    def _kv3(self, t):
        return t.view(self.T, self.d.heads, self.d.qk_nope + self.d.v_head)

    def _bshd(self, t3):
        return t3.unflatten(0, (self.B, self.S))

    # ------------------------------------------------------------------------------------------ forward
    def _mla_forward(self, i: int) -> None:
        d, a, w = self.d, self.act, self.w
        h1 = a["h1"][i]
        gemm(h1, w[f"wqd.{i}"], "TN", out=a["cq"][i])
        E.rmsnorm_fwd(a["cq"][i], w[f"qln.{i}"], d.eps, out=a["cqn"][i], rstd=a["rstdq"][i])
        gemm(a["cqn"][i], w[f"wqu.{i}"], "TN", out=a["q"][i])
        q3 = self._q3(a["q"][i])
        E.rope_strided_(q3[:, :, d.qk_nope:], self.cos, self.sin, self.S)
        gemm(h1, w[f"wkvd.{i}"], "TN", out=a["ckv"][i])
        ckv = a["ckv"][i]
        E.rope_strided_(ckv[:, d.kv_lora:].unsqueeze(1), self.cos, self.sin, self.S)
        E.copy2d_(a["clat"][i], ckv[:, :d.kv_lora])
        E.rmsnorm_fwd(a["clat"][i], w[f"kvln.{i}"], d.eps, out=a["ckvn"][i], rstd=a["rstdkv"][i])
        gemm(a["ckvn"][i], w[f"wkvu.{i}"], "TN", out=a["kv"][i])
        kv3 = self._kv3(a["kv"][i])
        k3 = self._q3(a["k"][i])
        E.mla_k_assemble(kv3, ckv[:, d.kv_lora:], k3, d.qk_nope)
        o4 = a["o"][i].view(self.B, self.S, d.heads, d.v_head)
        fmha_fwd(self._bshd(q3), self._bshd(k3), self._bshd(kv3[:, :, d.qk_nope:]), causal=True, scale=self.scale, out=o4, lse=a["lse"][i])
        gemm(a["o"][i], w[f"wo.{i}"], "TN", out=self.tmp_h)

    def _mlp_forward(self, i: int) -> torch.Tensor:
        """Returns the MLP output of layer ``i`` (in ``tmp_h``)."""
        d, a, w = self.d, self.act, self.w
        h2 = a["h2"][i]
        m = self.act_mlp[i]
        if i < d.dense_layers:
            gemm(h2, w[f"wfc1.{i}"], "TN", out=m["u"])
            E.swiglu_fwd(m["u"], out=m["g"])
            return gemm(m["g"], w[f"wfc2.{i}"], "TN", out=self.tmp_h)
        self.moe[i].forward(h2, out=self.tmp_h)
        if d.ffn_shared:
            gemm(h2, w[f"ws1.{i}"], "TN", out=m["u"])
            E.swiglu_fwd(m["u"], out=m["g"])
            gemm(m["g"], w[f"ws2.{i}"], "TN", out=self.tmp_h2)
            E.add(self.tmp_h, self.tmp_h2, out=self.tmp_h)
        return self.tmp_h

    def forward(self, tokens: torch.Tensor, targets: torch.Tensor, loss_scale: float) -> torch.Tensor:
        d, a, w = self.d, self.act, self.w
        L = d.layers
        E.embedding_fwd(tokens, w["embed"], a["x"][0])
        E.rmsnorm_fwd(a["x"][0], w["ln1.0"], d.eps, out=a["h1"][0], rstd=a["rstd1"][0])
        for i in range(L):
            self._mla_forward(i)
            E.add_rmsnorm_fwd(a["x"][i], self.tmp_h, w[f"ln2.{i}"], d.eps, out_sum=a["x2"][i], out_normed=a["h2"][i], rstd=a["rstd2"][i])
            y = self._mlp_forward(i)
            if i + 1 < L:
                E.add_rmsnorm_fwd(a["x2"][i], y, w[f"ln1.{i + 1}"], d.eps, out_sum=a["x"][i + 1], out_normed=a["h1"][i + 1], rstd=a["rstd1"][i + 1])
            else:
                E.add_rmsnorm_fwd(a["x2"][i], y, w["final_norm"], d.eps, out_sum=a["x"][L], out_normed=a["xf"], rstd=a["rstdf"])
        gemm(a["xf"], w["lm_head"], "TN", out=self.logits)
        return E.cross_entropy_(self.logits, targets, grad_scale=loss_scale)        # logits now hold dlogits

    # ------------------------------------------------------------------------------------------ backward
    def _mlp_backward(self, i: int) -> None:
        """dx (gradient of the layer output) -> ``dh`` = gradient w.r.t. the MLP input h2."""
        d, a, w, g = self.d, self.act, self.w, self.g
        h2, dx = a["h2"][i], self.dx
        m = self.act_mlp[i]
        if i < d.dense_layers:
            F = d.ffn_dense
            du, dg = self._d_u[:self.T * 2 * F].view(self.T, 2 * F), self._d_g[:self.T * F].view(self.T, F)
            gemm(dx, m["g"], "NT", out=g[f"wfc2.{i}"], accumulate=True)
            gemm(dx, w[f"wfc2.{i}"], "NN", out=dg)
            E.swiglu_bwd(m["u"], dg, out=du)
            gemm(du, h2, "NT", out=g[f"wfc1.{i}"], accumulate=True)
            gemm(du, w[f"wfc1.{i}"], "NN", out=self.dh)
            return
        self.moe[i].backward(dx, dx=self.dh)
        if d.ffn_shared:
            F = d.ffn_shared
            du, dg = self._d_u[:self.T * 2 * F].view(self.T, 2 * F), self._d_g[:self.T * F].view(self.T, F)
            gemm(dx, m["g"], "NT", out=g[f"ws2.{i}"], accumulate=True)
            gemm(dx, w[f"ws2.{i}"], "NN", out=dg)
            E.swiglu_bwd(m["u"], dg, out=du)
            gemm(du, h2, "NT", out=g[f"ws1.{i}"], accumulate=True)
            gemm(du, w[f"ws1.{i}"], "NN", out=self.tmp_h2)
            E.add(self.dh, self.tmp_h2, out=self.dh)

    def _mla_backward(self, i: int) -> None:
        """dx -> ``dh`` = gradient w.r.t. the attention input h1 (both down projections)."""
        d, a, w, g = self.d, self.act, self.w, self.g
        dx = self.dx
        gemm(dx, a["o"][i], "NT", out=g[f"wo.{i}"], accumulate=True)
        gemm(dx, w[f"wo.{i}"], "NN", out=self.d_o)
        q3, k3, kv3 = self._q3(a["q"][i]), self._q3(a["k"][i]), self._kv3(a["kv"][i])
        dq3, dk3, dkv3 = self._q3(self.d_q), self._q3(self.d_k), self._kv3(self.d_kv)
        v4, dv4 = self._bshd(kv3[:, :, d.qk_nope:]), self._bshd(dkv3[:, :, d.qk_nope:])
        fmha_bwd(self.d_o.view(self.B, self.S, d.heads, d.v_head), self._bshd(q3), self._bshd(k3), v4,
                 a["o"][i].view(self.B, self.S, d.heads, d.v_head), a["lse"][i], self._bshd(dq3), self._bshd(dk3), dv4,
                 causal=True, scale=self.scale)
        # keys: k_nope gradient back into the kv_up output gradient, rotary slice summed over the heads, both RoPEs undone
        E.mla_k_split_bwd(dk3, dkv3, self.d_ckv[:, d.kv_lora:], d.qk_nope)
        E.rope_strided_(self.d_ckv[:, d.kv_lora:].unsqueeze(1), self.cos, self.sin, self.S, inverse=True)
        E.rope_strided_(dq3[:, :, d.qk_nope:], self.cos, self.sin, self.S, inverse=True)
        # kv path
        gemm(self.d_kv, a["ckvn"][i], "NT", out=g[f"wkvu.{i}"], accumulate=True)
        gemm(self.d_kv, w[f"wkvu.{i}"], "NN", out=self.d_lat_n)
        E.rmsnorm_bwd(a["clat"][i], w[f"kvln.{i}"], self.d_lat_n, a["rstdkv"][i], g[f"kvln.{i}"], dx=self.d_lat)
        E.copy2d_(self.d_ckv[:, :d.kv_lora], self.d_lat)
        gemm(self.d_ckv, a["h1"][i], "NT", out=g[f"wkvd.{i}"], accumulate=True)
        gemm(self.d_ckv, w[f"wkvd.{i}"], "NN", out=self.dh)
        # q path
        gemm(self.d_q, a["cqn"][i], "NT", out=g[f"wqu.{i}"], accumulate=True)
        gemm(self.d_q, w[f"wqu.{i}"], "NN", out=self.d_cqn)
        E.rmsnorm_bwd(a["cq"][i], w[f"qln.{i}"], self.d_cqn, a["rstdq"][i], g[f"qln.{i}"], dx=self.d_cq)
        gemm(self.d_cq, a["h1"][i], "NT", out=g[f"wqd.{i}"], accumulate=True)
        gemm(self.d_cq, w[f"wqd.{i}"], "NN", out=self.tmp_h2)
        E.add(self.dh, self.tmp_h2, out=self.dh)

    def backward(self, tokens: torch.Tensor) -> None:
        d, a, w, g = self.d, self.act, self.w, self.g
        L = d.layers
        dl = self.logits
        gemm(dl, a["xf"], "NT", out=g["lm_head"], accumulate=True)
        gemm(dl, w["lm_head"], "NN", out=self.tmp_h)
        E.rmsnorm_bwd(a["x"][L], w["final_norm"], self.tmp_h, a["rstdf"], g["final_norm"], dx=self.dx)
        for i in reversed(range(L)):
            self._mlp_backward(i)
            E.rmsnorm_bwd(a["x2"][i], w[f"ln2.{i}"], self.dh, a["rstd2"][i], g[f"ln2.{i}"], dx=self.dx, accumulate_dx=True)
            self._mla_backward(i)
            E.rmsnorm_bwd(a["x"][i], w[f"ln1.{i}"], self.dh, a["rstd1"][i], g[f"ln1.{i}"], dx=self.dx, accumulate_dx=True)
        E.embedding_bwd(tokens, self.dx, g["embed"])

    # ------------------------------------------------------------------------------------------ optimizer
    def optimizer_step(self) -> None:
        """ZeRO-1 over all ranks for the dense parameters (in-kernel reduce-scatter / all-gather), local AdamW for this rank's experts;
        one global gradient norm (dense shard + local experts, summed over ranks)."""
        self.step_count += 1
        b1, b2 = self.betas
        W, ws = self.world, self.ws
        if W > 1:
            if ws is not None:
                ws.reduce_scatter(self.shard * 4, in_off=self._off_g, out_off=self._off_g + self.rank * self.shard * 4, dtype=torch.float32)
            else:
                dist.reduce_scatter_tensor(self.grad_shard, self.grads.buf, group=self.group)
        self.sq_in.zero_()
        E.sqnorm(self.grad_shard, out=self.sq_in[:1])
        E.sqnorm(self.e_grads.buf, out=self.sq_in[:1])
        if ws is not None:
            ws.all_reduce(16, in_off=self._off_sc, out_off=self._off_sc + 256, dtype=torch.float32, ctas=1)
        elif W > 1:
            dist.all_reduce(self.sq_in, group=self.group)
        sq = self.sq_out[:1]
        lo = self.rank * self.shard
        # gradients hold the SUM over ranks of per-rank means (experts: of the per-rank-scaled token contributions) -> 1 / world
        E.adamw_(self.master, self.m, self.v, self.grad_shard, self.params.buf[lo:lo + self.shard], self.lr, b1, b2, 1e-8, self.wd,
                 self.step_count, sqnorm_t=sq, max_norm=self.max_norm, grad_prescale=1.0 / W, zero_grad=(W == 1))
        E.adamw_(self.e_master, self.e_m, self.e_v, self.e_grads.buf, self.e_params.buf, self.lr, b1, b2, 1e-8, self.wd,
                 self.step_count, sqnorm_t=sq, max_norm=self.max_norm, grad_prescale=1.0 / W, zero_grad=True)
        if W > 1:
            self.grads.buf.zero_()
            if ws is not None:
                ws.all_gather(self.shard * 2, in_off=self._off_p + lo * 2, out_off=self._off_p)
            else:
                dist.all_gather_into_tensor(self.params.buf, self.params.buf[lo:lo + self.shard].clone(), group=self.group)

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

    def train_step(self, host_batches) -> float:
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
