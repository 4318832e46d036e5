"""Collective calibration: sweep message sizes at 2/4/8 GPUs for (a) this repo's in-kernel NVLink collectives
(P2P ld/st and NVLS multimem, csrc/collectives.cu) and (b) NCCL as the baseline curve, then fit the model's
``time_us = wire_bytes / (gbps * eff) + latency_us * (n - 1) * scale + fixed_latency_us`` per op and per group size.

Launch under torchrun (one process per GPU):

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m simumax_b200.calib.collectives \
        --out gpurun_out/coll_sweep_8.json

Every time is device-side (CUDA events), max over ranks.  The fit is CPU-only and unit-tested without a GPU.
(reference: simu_tools/efficency_test/nccl_fit.py:17-61 fits nccl-tests text dumps; apply_ws_comm_model.py:31-58 writes
the per-comm-num tables.)
"""
from __future__ import annotations

import argparse
import json
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

GIB = 1024 ** 3
RING = {"all_reduce": (2, -1), "all_gather": (1, -1), "reduce_scatter": (1, -1), "all2all": (1, -1), "p2p": (1, 0)}


def wire_bytes(op: str, size: float, n: int) -> float:
    """Bytes the ring model charges for a collective on a ``size``-byte tensor over ``n`` ranks."""
    scale, offset = RING[op]
    w = size * scale
    return w + (w / n) * offset


def fit_bw_latency(op: str, n: int, sizes: Sequence[float], times_us: Sequence[float], nominal_gbps: float,
                   relative: bool = True) -> Dict[str, float]:
    """Least-squares line through (wire bytes, time); returns the JSON fields the system config stores.

    ``relative=True`` weights each point by 1/time so small and large messages count equally (the reference fits the
    first 11 nccl-tests rows unweighted, which lets the largest sizes dominate)."""
    x = np.array([wire_bytes(op, s, n) for s in sizes], dtype=np.float64)
    y = np.array(times_us, dtype=np.float64)
    w = 1.0 / y if relative else np.ones_like(y)
    A = np.stack([x * w, w], axis=1)
    (a, b), *_ = np.linalg.lstsq(A, y * w, rcond=None)
    a = max(a, 1e-15)
    b = max(b, 0.0)
    bw_gbs = 1.0 / a / GIB * 1e6          # GiB/s, the unit the model multiplies gbps by
    scale, _ = RING[op]
    hops = max(n - 1, 1) * scale if op != "p2p" else 1
    pred = a * x + b
    return {"efficient_factor": float(bw_gbs / nominal_gbps), "latency_us": float(b / hops), "fixed_latency_us": 0.0,
            "fit_bw_gibs": float(bw_gbs), "fit_intercept_us": float(b),
            "max_rel_err": float(np.max(np.abs(pred - y) / y)), "points": len(x)}


def apply_to_system(system: dict, fits_by_world: Dict[int, Dict[str, Dict[str, float]]],
                    nets: Sequence[str] = ("high_intra_node", "low_intra_node"),
                    fixed: Optional[Dict[str, dict]] = None) -> dict:
    """Write fitted collective parameters (largest world size = default, others in the ``*_by_comm_num`` tables)."""
    if not fits_by_world:
        return system
    top = max(fits_by_world)
    for net in nets:
        ops = system["networks"][net]["op"]
        for op, f in fits_by_world[top].items():
            cfg = ops.setdefault(op, {"scale": RING[op][0], "offset": RING[op][1]})
            cfg["efficient_factor"] = round(f["efficient_factor"], 4)
            cfg["latency_us"] = round(f["latency_us"], 4)
            cfg["fixed_latency_us"] = round(f.get("fixed_latency_us", 0.0), 4)
            if op == "p2p":
                continue
            cfg["efficient_factor_by_comm_num"] = {str(ws): round(fits_by_world[ws][op]["efficient_factor"], 4)
                                                   for ws in sorted(fits_by_world) if op in fits_by_world[ws]}
            cfg["latency_us_by_comm_num"] = {str(ws): round(fits_by_world[ws][op]["latency_us"], 4)
                                             for ws in sorted(fits_by_world) if op in fits_by_world[ws]}
            cfg["fixed_latency_us_by_comm_num"] = {}
            if fixed and op in fixed:
                cfg["fixed_latency_us"] = round(fixed[op]["fixed_latency_us"], 4)
                cfg["fixed_latency_us_by_comm_num"] = {k: round(v, 4) for k, v in fixed[op]["fixed_latency_us_by_comm_num"].items()}
    return system


# ------------------------------------------------------------------------------------------------------------------
def _timed(fn, iters, dist, torch, warm=3):
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
    t = torch.tensor([s.elapsed_time(e) / iters * 1e3], device="cuda")      # us
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def sweep(sizes: Optional[List[int]] = None, max_bytes: int = 256 << 20, ctas_options=(32, 64, 128)) -> dict:
    """Run inside an initialised NCCL process group. Returns {"world", "nvls", "impl": {op: {"sizes", "times_us"}}}."""
    import torch
    import torch.distributed as dist
    from ..parallel.symm import SymmetricWorkspace
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = sizes or [1 << k for k in range(16, 29) if (1 << k) <= max_bytes]      # 64 KiB .. 256 MiB
    ws = SymmetricWorkspace(max_bytes, max_bytes)
    src = torch.empty(max_bytes // 2, device="cuda", dtype=torch.bfloat16).normal_()
    dst = torch.empty(max_bytes // 2, device="cuda", dtype=torch.bfloat16)
    out = {"world": world, "nvls": bool(ws.has_nvls), "smx": {}, "nccl": {}, "smx_variant": {}}
    ops = ("all_gather", "reduce_scatter", "all_reduce", "all2all", "p2p")
    for op in ops:
        out["smx"][op] = {"sizes": [], "times_us": []}
        out["nccl"][op] = {"sizes": [], "times_us": []}
        out["smx_variant"][op] = []
    for total in sizes:
        iters = 30 if total <= (4 << 20) else (15 if total <= (64 << 20) else 6)
        shard = total // world // 16 * 16
        el, sh_el = total // 2, shard // 2
        variants = [("p2p", False)] + ([("nvls", True)] if ws.has_nvls else [])

        def best(run):
            cands = {}
            for tag, nv in variants:
                for c in ctas_options:
                    cands[f"{tag}_c{c}"] = _timed(lambda: run(nv, c), iters, dist, torch)
            k = min(cands, key=cands.get)
            return cands[k], k

        for op, run, nbytes in (
            ("all_gather", lambda nv, c: ws.all_gather(shard, use_nvls=nv, ctas=c), shard * world),
            ("reduce_scatter", lambda nv, c: ws.reduce_scatter(shard, use_nvls=nv, ctas=c), shard * world),
            ("all_reduce", lambda nv, c: ws.all_reduce(total, use_nvls=nv, ctas=c), total),
            ("all2all", lambda nv, c: ws.all_to_all(shard, ctas=c), shard * world),
        ):
            t, k = best(run)
            out["smx"][op]["sizes"].append(nbytes); out["smx"][op]["times_us"].append(t); out["smx_variant"][op].append(k)
        # neighbour put (pipeline send/recv analogue): every rank writes `total` bytes into its right neighbour
        t = _timed(lambda: ws.put((rank + 1) % world, total), iters, dist, torch)
        out["smx"]["p2p"]["sizes"].append(total); out["smx"]["p2p"]["times_us"].append(t); out["smx_variant"]["p2p"].append("put")
        # NCCL baseline curve
        for op, run, nbytes in (
            ("all_gather", lambda: dist.all_gather_into_tensor(dst[:sh_el * world], src[:sh_el]), shard * world),
            ("reduce_scatter", lambda: dist.reduce_scatter_tensor(dst[:sh_el], src[:sh_el * world]), shard * world),
            ("all_reduce", lambda: dist.all_reduce(src[:el]), total),
            ("all2all", lambda: dist.all_to_all_single(dst[:sh_el * world], src[:sh_el * world]), shard * world),
        ):
            out["nccl"][op]["sizes"].append(nbytes); out["nccl"][op]["times_us"].append(_timed(run, iters, dist, torch))

        def sendrecv():
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, src[:el], (rank + 1) % world),
                                           dist.P2POp(dist.irecv, dst[:el], (rank - 1) % world)])
            for r in reqs:
                r.wait()
        out["nccl"]["p2p"]["sizes"].append(total); out["nccl"]["p2p"]["times_us"].append(_timed(sendrecv, iters, dist, torch))
        src.normal_()           # all_reduce in place keeps summing: re-randomise to stay finite
        if rank == 0:
            row = "  ".join(f"{op[:3]} {wire_bytes(op, out['smx'][op]['sizes'][-1], world) / out['smx'][op]['times_us'][-1] / 1e3:6.1f}"
                            f"/{wire_bytes(op, out['nccl'][op]['sizes'][-1], world) / out['nccl'][op]['times_us'][-1] / 1e3:6.1f}" for op in ops)
            print(f"[coll n={world}] {total >> 10:8d} KiB  busGB/s smx/nccl: {row}", flush=True)
    return out


def measure_burst(total_bytes: int = 32 << 20, ops_per_window: int = 20, keep: int = 10, ctas: int = 32) -> dict:
    """Fixed-latency probe: collectives interleaved with a 6144^3 GEMM (this repo's tcgen05 kernel) the way a training
    step interleaves them; each collective is bracketed by its own CUDA events so launch gaps and the wake-up after a
    compute kernel are included.  Returns per-op mean microseconds over the middle ``keep`` of ``ops_per_window`` ops, max
    over ranks.  (reference: measure_comm_burst_window_worker.py:64-71,112-135,201-218 does this with torch.mm + NCCL.)"""
    import torch
    import torch.distributed as dist
    from ..ops.gemm import gemm
    from ..parallel.symm import SymmetricWorkspace
    world, rank = dist.get_world_size(), dist.get_rank()
    ws = SymmetricWorkspace(total_bytes, total_bytes)
    a = torch.randn(6144, 6144, device="cuda").bfloat16()
    b = torch.randn(6144, 6144, device="cuda").bfloat16()
    c = torch.empty(6144, 6144, device="cuda", dtype=torch.bfloat16)
    src = torch.empty(total_bytes // 2, device="cuda", dtype=torch.bfloat16).normal_()
    dst = torch.empty(total_bytes // 2, device="cuda", dtype=torch.bfloat16)
    shard = total_bytes // world // 16 * 16
    sh_el = shard // 2
    nv = bool(ws.has_nvls)
    runs = {
        "smx": {"all_reduce": lambda: ws.all_reduce(total_bytes, use_nvls=nv, ctas=ctas),
                "all_gather": lambda: ws.all_gather(shard, use_nvls=nv, ctas=ctas),
                "reduce_scatter": lambda: ws.reduce_scatter(shard, use_nvls=nv, ctas=ctas),
                "all2all": lambda: ws.all_to_all(shard, ctas=ctas)},
        "nccl": {"all_reduce": lambda: dist.all_reduce(src),
                 "all_gather": lambda: dist.all_gather_into_tensor(dst[:sh_el * world], src[:sh_el]),
                 "reduce_scatter": lambda: dist.reduce_scatter_tensor(dst[:sh_el], src[:sh_el * world]),
                 "all2all": lambda: dist.all_to_all_single(dst[:sh_el * world], src[:sh_el * world])},
    }
    sizes = {"all_reduce": total_bytes, "all_gather": shard * world, "reduce_scatter": shard * world, "all2all": shard * world}
    out = {"world": world, "bytes": sizes, "smx": {}, "nccl": {}}
    for impl, table in runs.items():
        for op, run in table.items():
            for _ in range(3):
                gemm(a, b, "TN", out=c); run()
            torch.cuda.synchronize(); dist.barrier()
            evs = []
            for _ in range(ops_per_window):
                gemm(a, b, "TN", out=c)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run(); e1.record()
                evs.append((e0, e1))
            torch.cuda.synchronize()
            lo = (ops_per_window - keep) // 2
            us = sum(e0.elapsed_time(e1) for e0, e1 in evs[lo:lo + keep]) / keep * 1e3
            t = torch.tensor([us], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[impl][op] = float(t)
            src.normal_()
    return out


def measure_p2p_overlap(total_bytes: int = 64 << 20, gemms: int = 8) -> dict:
    """Pipeline send/recv the way a 1F1B schedule issues it: NCCL ``isend/irecv`` to the ring neighbours on a side
    stream while the main stream runs GEMMs.  Reports the p2p time alone, the GEMM time alone, and both overlapped
    (so the model's ``p2p`` efficiency and the compute slowdown under communication are both visible)."""
    import torch
    import torch.distributed as dist
    from ..ops.gemm import gemm
    world, rank = dist.get_world_size(), dist.get_rank()
    a = torch.randn(6144, 6144, device="cuda").bfloat16()
    b = torch.randn(6144, 6144, device="cuda").bfloat16()
    c = torch.empty(6144, 6144, device="cuda", dtype=torch.bfloat16)
    src = torch.empty(total_bytes // 2, device="cuda", dtype=torch.bfloat16).normal_()
    dst = torch.empty_like(src)
    side = torch.cuda.Stream()

    def compute():
        for _ in range(gemms):
            gemm(a, b, "TN", out=c)

    def p2p():
        reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, src, (rank + 1) % world),
                                       dist.P2POp(dist.irecv, dst, (rank - 1) % world)])
        for r in reqs:
            r.wait()

    def both():
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            p2p()
        compute()
        torch.cuda.current_stream().wait_stream(side)

    out = {"world": world, "bytes": total_bytes}
    for name, fn in (("gemm_only_us", compute), ("p2p_only_us", p2p), ("overlapped_us", both)):
        out[name] = _timed(fn, 5, dist, torch)
    out["p2p_GBps_alone"] = total_bytes / out["p2p_only_us"] / 1e3
    out["exposed_us"] = max(0.0, out["overlapped_us"] - out["gemm_only_us"])
    return out


def fixed_latency_from_burst(burst_by_world: Dict[int, dict], fits_by_world: Dict[int, Dict[str, Dict[str, float]]],
                             which: str = "smx", spread_us: float = 25.0) -> Dict[str, dict]:
    """fixed_us = burst-measured time - what the fitted line predicts; one scalar per op when the group sizes agree
    within ``spread_us``, otherwise a per-comm-num table (reference run_one_click_benchmark.py:371-414)."""
    out: Dict[str, dict] = {}
    ops = set()
    for b in burst_by_world.values():
        ops |= set(b[which])
    for op in ops:
        per = {}
        for n, b in burst_by_world.items():
            f = fits_by_world.get(n, {}).get(op)
            if f is None or op not in b[which]:
                continue
            x = wire_bytes(op, b["bytes"][op], n)
            model = x / (f["fit_bw_gibs"] * GIB) * 1e6 + f["fit_intercept_us"]
            per[str(n)] = max(0.0, b[which][op] - model)
        if not per:
            continue
        vals = list(per.values())
        if max(vals) - min(vals) <= spread_us:
            out[op] = {"fixed_latency_us": sum(vals) / len(vals), "fixed_latency_us_by_comm_num": {}}
        else:
            out[op] = {"fixed_latency_us": 0.0, "fixed_latency_us_by_comm_num": per}
    return out


def fit_all(sweep_result: dict, nominal_gbps: float = 900.0, which: str = "smx") -> Dict[str, Dict[str, float]]:
    n = sweep_result["world"]
    return {op: fit_bw_latency(op, n, d["sizes"], d["times_us"], nominal_gbps)
            for op, d in sweep_result[which].items() if d["sizes"]}


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--max-bytes", type=int, default=256 << 20)
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = sweep(max_bytes=a.max_bytes)
    res["burst"] = measure_burst()
    res["p2p_overlap"] = measure_p2p_overlap()
    if dist.get_rank() == 0:
        res["fit_smx"] = fit_all(res, which="smx")
        res["fit_nccl"] = fit_all(res, which="nccl")
        path = a.out or f"gpurun_out/coll_sweep_{res['world']}.json"
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        json.dump(res, open(path, "w"), indent=1)
        for op in res["fit_smx"]:
            s, c = res["fit_smx"][op], res["fit_nccl"][op]
            print(f"[fit n={res['world']}] {op:15s} smx eff {s['efficient_factor']:.3f} lat {s['latency_us']:.2f}us (err {s['max_rel_err']:.2f})"
                  f"   nccl eff {c['efficient_factor']:.3f} lat {c['latency_us']:.2f}us (err {c['max_rel_err']:.2f})")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
