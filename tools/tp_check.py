"""Tensor-parallel + sequence-parallel training step (fused AG->GEMM / GEMM->RS kernels) vs the single-GPU trainer with the
same full weights and the same batch, under torchrun on >= 2 GPUs.  Checks the loss, the final hidden state, a set of weight
gradients (sharded matrices, replicated norms, vocab-parallel embedding) and the loss trajectory over optimizer steps, then
times the step.  Prints one JSON line on rank 0 and writes gpurun_out/tp_check_<world>.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from simumax_b200.ops import _lib
from simumax_b200.runtime.tp_trainer import TPLlamaTrainer
from simumax_b200.runtime.trainer import LlamaDims, LlamaTrainer


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    W = world
    dims = LlamaDims(hidden=1024, heads=2 * W, kv_heads=W, head_dim=128, ffn=1024 * W, layers=2, vocab=4096 * W)
    B, S = 1, 512 * W
    T, Ts = B * S, B * S // W
    gen = torch.Generator(device="cuda").manual_seed(5)
    batches = [(torch.randint(0, dims.vocab, (T,), device="cuda", generator=gen),
                torch.randint(0, dims.vocab, (T,), device="cuda", generator=gen)) for _ in range(2)]
    ref = LlamaTrainer(dims, B, S, data_parallel=False, seed=11, lr=2e-3)
    tp = TPLlamaTrainer(dims, B, S, seed=11, lr=2e-3)
    tp.load_full_weights(ref.w)
    out = {"world": world, "dims": dims.__dict__, "tokens": T}
    # ---- one forward / backward, no optimizer: activations and gradients
    tok, tgt = batches[0]
    scale = 1.0 / T
    l_ref = float((ref.forward(tok, tgt, scale).sum() * scale).item())
    ref.backward(tok)
    l_tp = float((tp.forward(tok, tgt, scale).sum() * scale).item())
    tp.backward(tok)
    torch.cuda.synchronize()
    L, hd, nq, nk, FW = dims.layers, dims.head_dim, tp.nq, tp.nk, tp.FW
    r = rank
    sl = slice(r * Ts, (r + 1) * Ts)
    checks = {
        "loss": abs(l_tp - l_ref) / abs(l_ref),
        "x_last": rel(tp.act["x"][L], ref.act["x"][L][sl]),
        "g_lm_head": rel(tp.g["lm_head"], ref.g["lm_head"][tp.v0:tp.v0 + tp.VW]),
        "g_embed": rel(tp.g["embed"], ref.g["embed"][tp.v0:tp.v0 + tp.VW]),
        "g_wfc2_last": rel(tp.g["wfc2"][L - 1], ref.g["wfc2"][L - 1][:, r * FW:(r + 1) * FW]),
        "g_wfc1_0": rel(tp.g["wfc1"][0], torch.cat([ref.g["wfc1"][0][r * FW:(r + 1) * FW],
                                                    ref.g["wfc1"][0][dims.ffn + r * FW: dims.ffn + (r + 1) * FW]])),
        "g_wo_0": rel(tp.g["wo"][0], ref.g["wo"][0][:, r * nq * hd:(r + 1) * nq * hd]),
        "g_wq_0": rel(tp.g["wqkv"][0][: nq * hd], ref.g["wqkv"][0][r * nq * hd:(r + 1) * nq * hd]),
        "g_wv_0": rel(tp.g["wqkv"][0][(nq + nk) * hd:], ref.g["wqkv"][0][(dims.heads + dims.kv_heads) * hd + r * nk * hd:
                                                                        (dims.heads + dims.kv_heads) * hd + (r + 1) * nk * hd]),
    }
    for name in ("ln1", "ln2"):
        gsum = tp.g[name][0].clone()
        dist.all_reduce(gsum)
        checks[f"g_{name}_0"] = rel(gsum, ref.g[name][0])
    t = torch.tensor([max(checks.values())], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["checks_rank0"] = checks
    out["max_rel_err_all_ranks"] = float(t)
    ok = float(t) < 3e-2
    # ---- training trajectory: both trainers take the same steps
    ref.grads.buf.zero_(); tp.grads.buf.zero_()
    lr_, lt_ = [], []
    for _ in range(4):
        lr_.append(float(ref.run_microbatches(batches).item()))
        lt_.append(float(tp.run_microbatches(batches).item()))
    out["losses_ref"], out["losses_tp"] = lr_, lt_
    traj = max(abs(a - b) / abs(a) for a, b in zip(lr_, lt_))
    out["max_rel_loss_diff"] = traj
    ok = ok and traj < 2e-2 and lt_[-1] < lt_[0]
    # ---- timing (device events, max over ranks)
    def timed(fn, iters=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        v = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v)
    out["ms_per_step_ref_1gpu"] = timed(lambda: ref.run_microbatches(batches))
    _lib.reset_launches()
    out["ms_per_step_tp"] = timed(lambda: tp.run_microbatches(batches))
    out["launches_per_step_tp"] = _lib.launches() // 7
    out["ok"] = bool(ok)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/tp_check_{world}.json", "w"), indent=1)
        print(json.dumps(out))
        print("TP_OK" if ok else "TP_MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
