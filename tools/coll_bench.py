"""Multi-GPU collective check + sweep: in-kernel P2P / NVLS collectives vs NCCL (torch.distributed).

Launch:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/coll_bench.py
Timing: CUDA events on the launching stream after a barrier, max over ranks; bus bandwidth uses the ring
convention bytes*(n-1)/n*scale (all_reduce scale 2) so the numbers are comparable with nccl-tests / the reference fit.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from simumax_b200.parallel.symm import SymmetricWorkspace


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    max_total = int(os.environ.get("COLL_MAX_BYTES", 256 << 20))
    ws = SymmetricWorkspace(max_total, max_total)
    res = {"world": world, "nvls": ws.has_nvls, "rows": []}
    # ---------------- correctness on a small odd-ish size
    n = 8 * 1024 * world * 3
    torch.manual_seed(rank)
    x = torch.randn(n, device="cuda").bfloat16()
    shard = n // world
    ok = {}
    for nvls in ([False, True] if ws.has_nvls else [False]):
        tag = "nvls" if nvls else "p2p"
        ws.input(torch.bfloat16, shard).copy_(x[:shard]); ws.all_gather(shard * 2, use_nvls=nvls); torch.cuda.synchronize()
        ref = torch.empty(n, device="cuda", dtype=torch.bfloat16); dist.all_gather_into_tensor(ref, x[:shard].contiguous())
        ok[f"all_gather_{tag}"] = bool(torch.equal(ws.output(torch.bfloat16, n), ref))
        ws.input(torch.bfloat16, n).copy_(x); ws.reduce_scatter(shard * 2, use_nvls=nvls); torch.cuda.synchronize()
        ref = torch.empty(shard, device="cuda", dtype=torch.float32); dist.reduce_scatter_tensor(ref, x.float())
        ok[f"reduce_scatter_{tag}"] = bool(torch.allclose(ws.output(torch.bfloat16, shard).float(), ref, atol=0.1, rtol=0.02))
        ws.input(torch.bfloat16, n).copy_(x); ws.all_reduce(n * 2, use_nvls=nvls); torch.cuda.synchronize()
        ref = x.float().clone(); dist.all_reduce(ref)
        ok[f"all_reduce_{tag}"] = bool(torch.allclose(ws.output(torch.bfloat16, n).float(), ref, atol=0.1, rtol=0.02))
        # fp32 reductions (ZeRO-1 main-grad path)
        xf = x.float()
        ws.input(torch.float32, n).copy_(xf); ws.reduce_scatter(shard * 4, use_nvls=nvls, dtype=torch.float32); torch.cuda.synchronize()
        ref = torch.empty(shard, device="cuda", dtype=torch.float32); dist.reduce_scatter_tensor(ref, xf.clone())
        ok[f"reduce_scatter_f32_{tag}"] = bool(torch.allclose(ws.output(torch.float32, shard), ref, atol=1e-4, rtol=1e-5))
        ws.input(torch.float32, n).copy_(xf); ws.all_reduce(n * 4, use_nvls=nvls, dtype=torch.float32); torch.cuda.synchronize()
        ref = xf.clone(); dist.all_reduce(ref)
        ok[f"all_reduce_f32_{tag}"] = bool(torch.allclose(ws.output(torch.float32, n), ref, atol=1e-4, rtol=1e-5))
    ws.input(torch.bfloat16, n).copy_(x); ws.all_to_all(shard * 2); torch.cuda.synchronize()
    ref = torch.empty(n, device="cuda", dtype=torch.bfloat16); dist.all_to_all_single(ref, x.contiguous())
    ok["all_to_all_p2p"] = bool(torch.equal(ws.output(torch.bfloat16, n), ref))
    flag = torch.tensor([int(all(ok.values()))], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["correct"] = ok
    res["all_ranks_correct"] = bool(int(flag))
    if rank == 0:
        print("correctness", ok, "all ranks:", bool(int(flag)), "nvls:", ws.has_nvls, flush=True)
    # ---------------- bandwidth sweep
    sizes = [1 << 20, 4 << 20, 16 << 20, 64 << 20, 256 << 20]
    sizes = [s for s in sizes if s <= max_total]
    nccl_in = torch.empty(max_total // 2, device="cuda", dtype=torch.bfloat16).normal_()
    nccl_out = torch.empty(max_total // 2, device="cuda", dtype=torch.bfloat16)
    for total in sizes:
        iters = 20 if total <= (16 << 20) else 8
        shard_b = total // world // 16 * 16
        el = total // 2
        sh_el = shard_b // 2
        f = (world - 1) / world
        row = {"bytes": total}
        variants = [("p2p", False)] + ([("nvls", True)] if ws.has_nvls else [])
        for tag, nv in variants:
            for ctas in (32, 64, 128):
                t = timed(lambda: ws.all_gather(shard_b, use_nvls=nv, ctas=ctas), iters); row[f"ag_{tag}_c{ctas}_GBps"] = shard_b * world * f / t / 1e6
                t = timed(lambda: ws.reduce_scatter(shard_b, use_nvls=nv, ctas=ctas), iters); row[f"rs_{tag}_c{ctas}_GBps"] = shard_b * world * f / t / 1e6
                t = timed(lambda: ws.all_reduce(total, use_nvls=nv, ctas=ctas), iters); row[f"ar_{tag}_c{ctas}_GBps"] = total * 2 * f / t / 1e6
        for ctas in (32, 64, 128):
            t = timed(lambda: ws.all_to_all(shard_b, ctas=ctas), iters); row[f"a2a_p2p_c{ctas}_GBps"] = shard_b * world * f / t / 1e6
        t = timed(lambda: dist.all_gather_into_tensor(nccl_out[:sh_el * world], nccl_in[:sh_el]), iters); row["ag_nccl_GBps"] = shard_b * world * f / t / 1e6
        t = timed(lambda: dist.reduce_scatter_tensor(nccl_out[:sh_el], nccl_in[:sh_el * world]), iters); row["rs_nccl_GBps"] = shard_b * world * f / t / 1e6
        t = timed(lambda: dist.all_reduce(nccl_in[:el]), iters); row["ar_nccl_GBps"] = total * 2 * f / t / 1e6
        t = timed(lambda: dist.all_to_all_single(nccl_out[:sh_el * world], nccl_in[:sh_el * world]), iters); row["a2a_nccl_GBps"] = shard_b * world * f / t / 1e6
        res["rows"].append(row)
        if rank == 0:
            best = lambda p: max(v for k, v in row.items() if k.startswith(p) and "nccl" not in k)
            print(f"{total >> 20:4d} MiB  AG {best('ag_'):7.1f} (nccl {row['ag_nccl_GBps']:7.1f})  RS {best('rs_'):7.1f} (nccl {row['rs_nccl_GBps']:7.1f})  "
                  f"AR {best('ar_'):7.1f} (nccl {row['ar_nccl_GBps']:7.1f})  A2A {best('a2a_'):7.1f} (nccl {row['a2a_nccl_GBps']:7.1f})", flush=True)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(res, open(f"gpurun_out/coll_bench_{world}.json", "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
