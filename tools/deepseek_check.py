"""DeepSeek-style training step (MLA attention + dense / MoE MLP with shared expert, expert parallel) on the sm_100a kernels vs an fp32
PyTorch autograd model with the same weights, the same batch and the same routing decisions.  1 GPU (ep = 1) or under torchrun on
N GPUs (ep = dp = N).  Checks the loss, gradients of every parameter kind (MLA down / up projections, their norms, output projection,
router, shared expert, routed experts, dense MLP, embedding, head) and that a few optimizer steps reduce the loss identically on all
ranks.  Writes gpurun_out/deepseek_check_<world>.json and prints DEEPSEEK_OK / DEEPSEEK_MISMATCH."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F_

from simumax_b200.runtime.deepseek_trainer import DeepSeekDims, DeepSeekTrainer


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def reference(tr: DeepSeekTrainer, tokens, targets, e_w1, e_w2):
    """fp32 autograd model; ``e_w1`` / ``e_w2``: per MoE layer the FULL expert weights [E, 2F, H] / [E, H, F] (gathered over ranks)."""
    d = tr.d
    T, H, nh = tr.T, d.hidden, d.heads
    W = {n: t.detach().float().requires_grad_(True) for n, t in tr.w.items()}
    EW1 = {i: t.detach().float().requires_grad_(True) for i, t in e_w1.items()}
    EW2 = {i: t.detach().float().requires_grad_(True) for i, t in e_w2.items()}
    cos, sin = tr.cos, tr.sin                      # [S, 32]

    def rms(x, g):
        return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + d.eps) * g

    def rope(t):                                   # t [B, S, h, 64], rotate-half pairs (i, i + 32)
        h = t.shape[-1] // 2
        a, b = t[..., :h], t[..., h:]
        c, s = cos[None, :, None, :], sin[None, :, None, :]
        return torch.cat([a * c - b * s, b * c + a * s], -1)

    x = W["embed"][tokens]
    for i in range(d.layers):
        h1 = rms(x, W[f"ln1.{i}"])
        q = (rms(h1 @ W[f"wqd.{i}"].t(), W[f"qln.{i}"]) @ W[f"wqu.{i}"].t()).view(tr.B, tr.S, nh, d.dqk)
        q = torch.cat([q[..., :d.qk_nope], rope(q[..., d.qk_nope:])], -1)
        ckv = h1 @ W[f"wkvd.{i}"].t()
        kr = rope(ckv[:, d.kv_lora:].view(tr.B, tr.S, 1, d.qk_rope))
        kv = (rms(ckv[:, :d.kv_lora], W[f"kvln.{i}"]) @ W[f"wkvu.{i}"].t()).view(tr.B, tr.S, nh, d.qk_nope + d.v_head)
        k = torch.cat([kv[..., :d.qk_nope], kr.expand(-1, -1, nh, -1)], -1)
        v = kv[..., d.qk_nope:]
        o = F_.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True, scale=tr.scale)
        x = x + o.transpose(1, 2).reshape(T, nh * d.v_head) @ W[f"wo.{i}"].t()
        h2 = rms(x, W[f"ln2.{i}"])
        if i < d.dense_layers:
            u = h2 @ W[f"wfc1.{i}"].t()
            Fd = d.ffn_dense
            y = (F_.silu(u[:, :Fd]) * u[:, Fd:]) @ W[f"wfc2.{i}"].t()
        else:
            m = tr.moe[i]
            scores = torch.softmax(h2 @ W[f"router.{i}"].t(), dim=-1)
            probs = scores.gather(1, m.idx)
            keep = m.inv_map.view(T, m.k) >= 0
            y = torch.zeros(T, H, device=x.device)
            Fe = d.ffn_expert
            for e in range(d.experts):
                sel = (m.idx == e) & keep
                if not sel.any():
                    continue
                t_idx, j_idx = sel.nonzero(as_tuple=True)
                hh = h2[t_idx] @ EW1[i][e].t()
                y = y.index_add(0, t_idx, ((F_.silu(hh[:, :Fe]) * hh[:, Fe:]) @ EW2[i][e].t()) * probs[t_idx, j_idx].unsqueeze(1))
            if d.ffn_shared:
                us = h2 @ W[f"ws1.{i}"].t()
                Fs = d.ffn_shared
                y = y + (F_.silu(us[:, :Fs]) * us[:, Fs:]) @ W[f"ws2.{i}"].t()
        x = x + y
    logits = rms(x, W["final_norm"]) @ W["lm_head"].t()
    loss = F_.cross_entropy(logits, targets)
    loss.backward()
    return loss.item(), W, EW1, EW2


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if "RANK" not in os.environ:
        os.environ.update(RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="29577")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    dims = DeepSeekDims(hidden=512, heads=4, qk_nope=128, qk_rope=64, v_head=128, q_lora=256, kv_lora=128, ffn_dense=1024, ffn_expert=256,
                        ffn_shared=512, experts=8, topk=2, layers=3, dense_layers=1, vocab=2048)
    B, S = 1, 512
    tr = DeepSeekTrainer(dims, B, S, seed=17, lr=1e-3, max_grad_norm=0.0)
    for n, t in tr.w.items():                      # larger weights -> non-trivial attention / routing
        if t.dim() == 2:
            t.mul_(2.0)
    tr.e_params.buf.mul_(2.0)
    lo = tr.rank * tr.shard
    tr.master.copy_(tr.params.buf[lo:lo + tr.shard])
    tr.e_master.copy_(tr.e_params.buf)
    gen = torch.Generator(device="cuda").manual_seed(100 + rank)
    tok = torch.randint(0, dims.vocab, (B * S,), device="cuda", generator=gen)
    tgt = torch.randint(0, dims.vocab, (B * S,), device="cuda", generator=gen)
    rows = tr.forward(tok, tgt, loss_scale=1.0 / tr.T)
    loss = rows.mean().item()
    tr.backward(tok)
    torch.cuda.synchronize()
    # full expert weights for the reference
    e_w1, e_w2 = {}, {}
    for i, m in tr.moe.items():
        a = [torch.empty_like(m.w1) for _ in range(world)]; b = [torch.empty_like(m.w2) for _ in range(world)]
        dist.all_gather(a, m.w1.contiguous()); dist.all_gather(b, m.w2.contiguous())
        e_w1[i], e_w2[i] = torch.cat(a), torch.cat(b)
    ref_loss, W, EW1, EW2 = reference(tr, tok, tgt, e_w1, e_w2)
    checks = {"loss": abs(loss - ref_loss) / abs(ref_loss)}
    for n in ("lm_head", "final_norm", "embed", "wo.0", "wqu.0", "wqd.0", "qln.0", "wkvu.0", "wkvd.0", "kvln.0", "ln1.0", "ln2.0", "wfc1.0", "wfc2.0",
              "wo.2", "wqu.2", "wkvu.2", "wkvd.2", "router.1", "ws1.1", "ws2.2", "ln1.2"):
        checks["g_" + n] = rel(tr.g[n], W[n].grad)
    for i, m in tr.moe.items():                    # expert gradients collect tokens from every rank: sum the per-rank references
        g1, g2 = EW1[i].grad.clone(), EW2[i].grad.clone()
        dist.all_reduce(g1); dist.all_reduce(g2)
        El = tr.El
        checks[f"g_expert_w1.{i}"] = rel(m.g_w1, g1[rank * El:(rank + 1) * El])
        checks[f"g_expert_w2.{i}"] = rel(m.g_w2, g2[rank * El:(rank + 1) * El])
    worst = torch.tensor([max(checks.values())], device="cuda")
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    ok = float(worst) < 5e-2
    # optimizer: the loss on a fixed batch goes down, identically on every rank
    tr.grads.buf.zero_(); tr.e_grads.buf.zero_()
    losses = [float(tr.run_microbatches([(tok, tgt)]).item()) for _ in range(6)]
    ok = ok and losses[-1] < losses[0] and all(math.isfinite(l) for l in losses)
    p_sum = tr.params.buf.float().sum()
    p_all = [torch.empty_like(p_sum) for _ in range(world)]
    dist.all_gather(p_all, p_sum)
    same = all(float(p) == float(p_all[0]) for p in p_all)
    ok = ok and same
    flag = torch.tensor([int(ok)], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out = {"world": world, "dims": dims.__dict__, "tokens": tr.T, "capacity": next(iter(tr.moe.values())).C, "checks_rank0": checks,
           "max_rel_err_all_ranks": float(worst), "losses": losses, "dense_params_identical_on_all_ranks": same, "ok": bool(int(flag))}
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/deepseek_check_{world}.json", "w"), indent=1)
        print(json.dumps(out))
        print("DEEPSEEK_OK" if out["ok"] else "DEEPSEEK_MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
