"""Data-parallel ZeRO-1 step: in-kernel NVLink collectives vs the NCCL path, run under torchrun on >= 2 GPUs.
Checks (1) both communication back-ends produce the same parameters, (2) all ranks hold identical parameters,
(3) the loss goes down.  Prints one JSON line on rank 0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from simumax_b200.runtime.trainer import LlamaDims, LlamaTrainer


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    dims = LlamaDims(hidden=512, heads=4, kv_heads=2, head_dim=128, ffn=1024, layers=2, vocab=1024)
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    batches = [(torch.randint(0, dims.vocab, (2 * 256,), device="cuda", generator=g),
                torch.randint(0, dims.vocab, (2 * 256,), device="cuda", generator=g)) for _ in range(2)]
    out = {}
    params = {}
    for comm in ("smx", "nccl"):
        tr = LlamaTrainer(dims, micro_batch=2, seq_len=256, dp_group=dist.group.WORLD, comm=comm, seed=7, lr=3e-3)
        losses = [float(tr.run_microbatches(batches).item()) for _ in range(4)]
        torch.cuda.synchronize()
        params[comm] = tr.params.buf[: tr.P].float().clone()
        gathered = [torch.empty_like(params[comm]) for _ in range(world)]
        dist.all_gather(gathered, params[comm])
        out[comm] = {"comm_used": tr.comm, "losses": losses,
                     "ranks_identical": all(bool(torch.equal(gathered[0], t)) for t in gathered[1:])}
        del tr
    diff = (params["smx"] - params["nccl"]).abs().max().item()
    out["max_param_diff_smx_vs_nccl"] = diff
    out["ok"] = bool(out["smx"]["ranks_identical"] and out["nccl"]["ranks_identical"] and diff < 2e-2
                     and out["smx"]["losses"][-1] < out["smx"]["losses"][0] and out["smx"]["comm_used"] == "smx")
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
