"""Context-parallel (Ulysses all-to-all) training step vs the single-GPU trainer on the full sequence with the same weights, under torchrun
on N >= 2 GPUs (cp = N): loss, final hidden state, weight gradients (summed over the cp ranks), loss trajectory over optimizer steps,
without and with selective recompute.  Writes gpurun_out/cp_check_<N>.json and prints CP_OK / CP_MISMATCH."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from simumax_b200.runtime.trainer import LlamaDims, LlamaTrainer


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def cp_qkv_rows(dims, cp):
    """Row permutation that turns the [q | k | v] packed QKV weight into the context-parallel layout [q_r | k_r | v_r] for r in ranks."""
    hd, nq, nk = dims.head_dim, dims.heads, dims.kv_heads
    q0, k0, v0 = 0, nq * hd, (nq + nk) * hd
    idx = []
    for r in range(cp):
        idx += list(range(q0 + r * nq // cp * hd, q0 + (r + 1) * nq // cp * hd))
        idx += list(range(k0 + r * nk // cp * hd, k0 + (r + 1) * nk // cp * hd))
        idx += list(range(v0 + r * nk // cp * hd, v0 + (r + 1) * nk // cp * hd))
    return torch.tensor(idx, device="cuda")


def run(mode, dims, S, world, rank, batches_full):
    Tl = S // world
    ref = LlamaTrainer(dims, 1, S, data_parallel=False, seed=11, lr=2e-3)
    tr = LlamaTrainer(dims, 1, S, dp_group=dist.group.WORLD, cp_group=dist.group.WORLD, seed=11, lr=2e-3, recompute=mode)
    perm = cp_qkv_rows(dims, world)
    for name in ("embed", "final_norm", "lm_head"):
        tr.w[name].copy_(ref.w[name])
    for i in range(dims.layers):
        for name in ("ln1", "wo", "ln2", "wfc1", "wfc2"):
            tr.w[name][i].copy_(ref.w[name][i])
        tr.w["wqkv"][i].copy_(ref.w["wqkv"][i][perm])
    tr.master.copy_(torch.cat([tr._my_slice(tr.params.buf, b) for b in range(len(tr.buckets))]).float())
    mine = [(t[rank * Tl:(rank + 1) * Tl].contiguous(), y[rank * Tl:(rank + 1) * Tl].contiguous()) for t, y in batches_full]
    tok, tgt = batches_full[0]
    l_ref = float((ref.forward(tok, tgt, 1.0 / S).sum() / S).item()); ref.backward(tok)
    lt = tr.forward(mine[0][0], mine[0][1], 1.0 / Tl).sum() / Tl
    tr.backward(mine[0][0])
    dist.all_reduce(lt); lt /= world
    torch.cuda.synchronize()
    L = dims.layers
    checks = {"loss": abs(float(lt) - l_ref) / abs(l_ref), "x_last": rel(tr.act["x"][L], ref.act["x"][L][rank * Tl:(rank + 1) * Tl])}
    for name, i in (("wqkv", 0), ("wo", 0), ("wfc1", L - 1), ("wfc2", 0), ("ln1", 0), ("ln2", L - 1)):
        g = tr.g[name][i].clone(); dist.all_reduce(g); g /= world
        r = ref.g[name][i][perm] if name == "wqkv" else ref.g[name][i]
        checks[f"g_{name}_{i}"] = rel(g, r)
    for name in ("lm_head", "embed"):
        g = tr.g[name].clone(); dist.all_reduce(g); g /= world
        checks[f"g_{name}"] = rel(g, ref.g[name])
    ref.grads.buf.zero_(); tr.grads.buf.zero_()
    lr_, lt_ = [], []
    for _ in range(3):
        lr_.append(float(ref.run_microbatches(batches_full).item()))
        l = tr.run_microbatches(mine).clone(); dist.all_reduce(l); lt_.append(float(l.item()) / world)
    tr.synchronize()
    traj = max(abs(a - b) / abs(a) for a, b in zip(lr_, lt_))
    worst = torch.tensor([max(checks.values())], device="cuda"); dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    ok = float(worst) < 3e-2 and traj < 2e-2 and lt_[-1] < lt_[0]
    del ref, tr
    return {"recompute": mode, "checks_rank0": checks, "max_rel_err_all_ranks": float(worst), "losses_ref": lr_, "losses_cp": lt_, "max_rel_loss_diff": traj, "ok": ok}


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    dims = LlamaDims(hidden=1024, heads=2 * world, kv_heads=world, head_dim=128, ffn=2048, layers=2, vocab=4096)
    S = 512 * world
    gen = torch.Generator(device="cuda").manual_seed(5)
    batches = [(torch.randint(0, dims.vocab, (S,), device="cuda", generator=gen), torch.randint(0, dims.vocab, (S,), device="cuda", generator=gen))
               for _ in range(2)]
    out = {"world": world, "seq": S, "runs": [run(m, dims, S, world, rank, batches) for m in ("none", "selective")]}
    flag = torch.tensor([int(all(r["ok"] for r in out["runs"]))], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["ok"] = bool(int(flag))
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/cp_check_{world}.json", "w"), indent=1)
        print(json.dumps(out))
        print("CP_OK" if out["ok"] else "CP_MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
