"""In-process correctness gates for the multi-GPU runtime, callable from an already initialised process group.

``bench.py --gpus N`` runs them before it times anything, so the driver's scaling run is also a correctness run: in-kernel
collectives against NCCL on the same data, and a small tensor-parallel (x data-parallel) training step against the single-GPU
trainer with the same weights and batches.  ``tools/tp_check.py`` / ``tools/coll_bench.py`` are the long versions.
"""
from __future__ import annotations

from typing import Dict


def _rel(a, b) -> float:
    return float(((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item())


def check_collectives(nbytes_per_rank: int = 3 * 8 * 1024 * 2) -> Dict[str, object]:
    """All-gather / reduce-scatter / all-reduce (bf16 + fp32, P2P and NVLS flavours) and all-to-all of this repo's kernels
    vs ``torch.distributed`` (NCCL) on the same random data.  Collective over the default group."""
    import torch
    import torch.distributed as dist
    from ..parallel.symm import SymmetricWorkspace
    world, rank = dist.get_world_size(), dist.get_rank()
    shard = nbytes_per_rank // 2
    n = shard * world
    ws = SymmetricWorkspace(n * 4, n * 4)
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    x = torch.randn(n, device="cuda", generator=g).bfloat16()
    ok = {}
    for nvls in ([False, True] if ws.has_nvls else [False]):
        tag = "nvls" if nvls else "p2p"
        ws.input(torch.bfloat16, shard).copy_(x[:shard]); ws.all_gather(shard * 2, use_nvls=nvls)
        ref = torch.empty(n, device="cuda", dtype=torch.bfloat16); dist.all_gather_into_tensor(ref, x[:shard].contiguous())
        ok[f"all_gather_{tag}"] = bool(torch.equal(ws.output(torch.bfloat16, n), ref))
        xf = x.float()
        ws.input(torch.float32, n).copy_(xf); ws.reduce_scatter(shard * 4, use_nvls=nvls, dtype=torch.float32)
        ref = torch.empty(shard, device="cuda", dtype=torch.float32); dist.reduce_scatter_tensor(ref, xf.clone())
        ok[f"reduce_scatter_f32_{tag}"] = bool(torch.allclose(ws.output(torch.float32, shard), ref, atol=1e-4, rtol=1e-5))
        ws.input(torch.bfloat16, n).copy_(x); ws.reduce_scatter(shard * 2, use_nvls=nvls)
        ok[f"reduce_scatter_bf16_{tag}"] = bool(torch.allclose(ws.output(torch.bfloat16, shard).float(), ref, atol=0.1, rtol=0.02))
        ws.input(torch.float32, n).copy_(xf); ws.all_reduce(n * 4, use_nvls=nvls, dtype=torch.float32)
        ref = xf.clone(); dist.all_reduce(ref)
        ok[f"all_reduce_f32_{tag}"] = bool(torch.allclose(ws.output(torch.float32, n), ref, atol=1e-4, rtol=1e-5))
    ws.input(torch.bfloat16, n).copy_(x); ws.all_to_all(shard * 2)
    ref = torch.empty(n, device="cuda", dtype=torch.bfloat16); dist.all_to_all_single(ref, x.contiguous())
    ok["all_to_all"] = bool(torch.equal(ws.output(torch.bfloat16, n), ref))
    torch.cuda.synchronize()
    flag = torch.tensor([int(all(ok.values()))], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    del ws
    return {"checks": ok, "nvls": any(k.endswith("nvls") for k in ok), "ok": bool(int(flag))}


def check_tp_parity(tp: int = 0, recompute: str = "none", steps: int = 3) -> Dict[str, object]:
    """Tensor + sequence parallel (x data parallel when ``tp`` < world) training on the fused AG->GEMM / GEMM->RS kernels vs
    the single-GPU trainer with the same weights and the same batches: loss, last hidden state, weight gradients of every
    kind (column / row shards, replicated norms, vocab-parallel embedding) after one pass, then the loss trajectory over
    optimizer steps (ZeRO-1 included when dp > 1)."""
    import torch
    import torch.distributed as dist
    from .cases import Case, parallel_groups
    from .tp_trainer import TPLlamaTrainer
    from .trainer import LlamaDims, LlamaTrainer
    world, rank = dist.get_world_size(), dist.get_rank()
    W = tp or world
    dp = world // W
    tp_group, dp_group = parallel_groups(Case("selfcheck", "llama3-8b", tp=W, dp=dp))
    dims = LlamaDims(hidden=1024, heads=2 * W, kv_heads=W, head_dim=128, ffn=1024 * W, layers=2, vocab=4096 * W)
    B, S = 1, 512 * W
    T, Ts = B * S, B * S // W
    n_mb = 2
    gen = torch.Generator(device="cuda").manual_seed(5)
    # every replica's micro-batches; the single-GPU reference runs all dp * n_mb of them
    all_b = [(torch.randint(0, dims.vocab, (T,), device="cuda", generator=gen),
              torch.randint(0, dims.vocab, (T,), device="cuda", generator=gen)) for _ in range(n_mb * dp)]
    mine = all_b[(rank // W) * n_mb:(rank // W + 1) * n_mb]
    ref = LlamaTrainer(dims, B, S, data_parallel=False, seed=11, lr=2e-3)
    tr = TPLlamaTrainer(dims, B, S, tp_group=tp_group, dp_group=dp_group, seed=11, lr=2e-3, recompute=recompute)
    tr.load_full_weights(ref.w)
    r = rank % W
    tok, tgt = mine[0]
    scale = 1.0 / T
    l_ref = float((ref.forward(tok, tgt, scale).sum() * scale).item()); ref.backward(tok)
    l_tp = float((tr.forward(tok, tgt, scale).sum() * scale).item()); tr.backward(tok)
    torch.cuda.synchronize()
    L, hd, nq, nk, FW = dims.layers, dims.head_dim, tr.nq, tr.nk, tr.FW
    sl = slice(r * Ts, (r + 1) * Ts)
    checks = {
        "loss": abs(l_tp - l_ref) / abs(l_ref),
        "x_last": _rel(tr.act["x"][L], ref.act["x"][L][sl]),
        "g_lm_head": _rel(tr.g["lm_head"], ref.g["lm_head"][tr.v0:tr.v0 + tr.VW]),
        "g_embed": _rel(tr.g["embed"], ref.g["embed"][tr.v0:tr.v0 + tr.VW]),
        "g_wfc2_last": _rel(tr.g["wfc2"][L - 1], ref.g["wfc2"][L - 1][:, r * FW:(r + 1) * FW]),
        "g_wfc1_0": _rel(tr.g["wfc1"][0], torch.cat([ref.g["wfc1"][0][r * FW:(r + 1) * FW],
                                                     ref.g["wfc1"][0][dims.ffn + r * FW: dims.ffn + (r + 1) * FW]])),
        "g_wo_0": _rel(tr.g["wo"][0], ref.g["wo"][0][:, r * nq * hd:(r + 1) * nq * hd]),
        "g_wq_0": _rel(tr.g["wqkv"][0][: nq * hd], ref.g["wqkv"][0][r * nq * hd:(r + 1) * nq * hd]),
    }
    for name in ("ln1", "ln2"):
        gsum = tr.g[name][0].clone()
        dist.all_reduce(gsum, group=tp_group)
        checks[f"g_{name}_0"] = _rel(gsum, ref.g[name][0])
    worst = torch.tensor([max(checks.values())], device="cuda")
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    ref.grads.buf.zero_(); tr.grads.buf.zero_()
    lr_, lt_ = [], []
    for _ in range(steps):
        lr_.append(float(ref.run_microbatches(all_b).item()))
        lt = tr.run_microbatches(mine).clone()
        if dp > 1:
            dist.all_reduce(lt, group=dp_group)
            lt /= dp
        lt_.append(float(lt.item()))
    traj = max(abs(a - b) / abs(a) for a, b in zip(lr_, lt_))
    ok = float(worst) < 3e-2 and traj < 2e-2 and lt_[-1] < lt_[0]
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    del ref, tr
    return {"tp": W, "dp": dp, "recompute": recompute, "max_rel_err_all_ranks": float(worst), "checks_rank0": checks,
            "losses_ref_1gpu": lr_, "losses_parallel": lt_, "max_rel_loss_diff": traj, "ok": bool(int(flag))}


def check_recompute_parity() -> Dict[str, object]:
    """Single GPU: selective and full recompute re-run the same kernels on the same inputs, so the loss is identical and the gradients
    agree with the no-recompute run to the run-to-run noise of the fp32 atomics (RMSNorm dgamma, TMA reduce-add of dQ)."""
    import torch
    from .trainer import LlamaDims, LlamaTrainer
    dims = LlamaDims(hidden=512, heads=4, kv_heads=2, head_dim=128, ffn=1024, layers=3, vocab=2048)
    B, S = 1, 512
    gen = torch.Generator(device="cuda").manual_seed(3)
    tok = torch.randint(0, dims.vocab, (B * S,), device="cuda", generator=gen)
    tgt = torch.randint(0, dims.vocab, (B * S,), device="cuda", generator=gen)
    out, base = {}, None
    for mode, layers in (("none", None), ("selective", None), ("full", None), ("full", 2)):
        tr = LlamaTrainer(dims, B, S, data_parallel=False, seed=21, recompute=mode, recompute_layers=layers)
        rows = tr.forward(tok, tgt, 1.0 / (B * S))
        loss = float(rows.sum().item())
        tr.backward(tok)
        torch.cuda.synchronize()
        g = tr.grads.buf.clone()
        if base is None:
            base = (loss, g)
            out["act_bytes_none"] = tr.memory_bytes()["activations"]
        else:
            key = f"{mode}{'' if layers is None else layers}"
            out[f"{key}_loss_equal"] = loss == base[0]
            out[f"{key}_max_abs_grad_diff"] = float((g - base[1]).abs().max().item())
            out[f"{key}_rel_grad_diff"] = float(((g - base[1]).norm() / base[1].norm()).item())
            out[f"act_bytes_{key}"] = tr.memory_bytes()["activations"]
        del tr
    out["ok"] = all(v < 1e-5 for k, v in out.items() if k.endswith("rel_grad_diff")) and all(v for k, v in out.items() if k.endswith("loss_equal"))
    return out
