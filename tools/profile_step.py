"""Kernel-time shares of one training step of the primary bench case at the current world size (torch.profiler / CUPTI on rank 0;
shares only -- never a benchmark number).  1 GPU: python tools/profile_step.py; N GPUs: torchrun ... tools/profile_step.py [--case NAME]
Writes gpurun_out/step_profile_<N>.json."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from simumax_b200.runtime import cases as C

ap = argparse.ArgumentParser()
ap.add_argument("--mbc", type=int, default=2)
ap.add_argument("--layers", type=int, default=0)
ap.add_argument("--zoo-case", default=None, help="a case name of tools/b200/perf_vs_real.py::zoo() instead of the primary bench case")
a = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1 or a.zoo_case:             # DeepSeek / symmetric-memory trainers want a process group even on one GPU
    if "RANK" not in os.environ:
        os.environ.update(RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=os.environ.get("MASTER_PORT", "29572"))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if dist.is_initialized() else 0
if a.zoo_case:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "b200"))
    import perf_vs_real
    case = perf_vs_real.zoo()[a.zoo_case]
else:
    case = C.primary_case(world)
    case.mbc = a.mbc
if a.layers:
    case.layers = a.layers
mc, dims = C.model_dims(case)
tr = C.build_trainer(case)
g = torch.Generator(device="cuda").manual_seed(5)
T = case.mbs * case.seq // case.cp
batches = [(torch.randint(0, mc.vocab_size, (T,), device="cuda", generator=g), torch.randint(0, mc.vocab_size, (T,), device="cuda", generator=g))
           for _ in range(case.mbc)]
for _ in range(2):
    tr.run_microbatches(batches)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); tr.run_microbatches(batches); e1.record(); torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1)
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.run_microbatches(batches)
    torch.cuda.synchronize()
if rank == 0:
    rows = []
    for ev in prof.key_averages():
        t = getattr(ev, "device_time_total", None) or getattr(ev, "cuda_time_total", 0)
        if t:
            rows.append({"name": ev.key[:100], "count": ev.count, "total_us": t})
    rows.sort(key=lambda r: -r["total_us"])
    tot = sum(r["total_us"] for r in rows)
    for r in rows[:28]:
        print(f"{r['total_us'] / 1e3:9.3f} ms {100 * r['total_us'] / tot:5.1f}%  n={r['count']:5d}  {r['name'][:90]}")
    print(f"sum of kernel time {tot / 1e3:.2f} ms; step (unprofiled, device-timed) {step_ms:.2f} ms; busy fraction {tot / 1e3 / step_ms:.3f}")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"world": world, "case": case.name, "mbc": case.mbc, "layers": case.layers or dims.layers, "step_ms_unprofiled": step_ms, "kernel_ms_sum": tot / 1e3,
               "kernels": rows[:60]}, open(f"gpurun_out/step_profile_{world}{'_' + a.zoo_case if a.zoo_case else ''}.json", "w"), indent=1)
if dist.is_initialized():
    dist.destroy_process_group()
