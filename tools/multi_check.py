"""Multi-GPU correctness sweep in ONE process group (torchrun, N >= 2): in-kernel collectives vs NCCL, TP (+SP) trainer vs single GPU
without / with selective / with full recompute, and TP x DP (N >= 4) with the in-kernel ZeRO-1.  Writes gpurun_out/multi_check_<N>.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from simumax_b200.runtime import selfcheck


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    out = {"world": world, "collectives": selfcheck.check_collectives()}
    for mode in ("none", "selective", "full"):
        out[f"tp{world}_recompute_{mode}"] = selfcheck.check_tp_parity(recompute=mode)
    if world >= 4:
        out["tp2_x_dp"] = selfcheck.check_tp_parity(tp=2)
        out["tp2_x_dp_selective"] = selfcheck.check_tp_parity(tp=2, recompute="selective")
    ok = all(v["ok"] for k, v in out.items() if isinstance(v, dict))
    out["ok"] = ok
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/multi_check_{world}.json", "w"), indent=1)
        print(json.dumps({k: (v["ok"] if isinstance(v, dict) else v) for k, v in out.items()}))
        for k, v in out.items():
            if isinstance(v, dict) and "max_rel_err_all_ranks" in v:
                print(k, "max rel err", v["max_rel_err_all_ranks"], "traj", v["max_rel_loss_diff"], v["losses_ref_1gpu"], v["losses_parallel"])
        print("MULTI_OK" if ok else "MULTI_MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
