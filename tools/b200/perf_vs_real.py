#!/usr/bin/env python
"""Predicted-vs-measured table: run real training steps with this repo's runtime (LlamaTrainer) and compare iteration
time and peak memory with what ``PerfLLM`` predicts for the same (system, strategy, model) triple.

    python tools/b200/perf_vs_real.py --system b200_sm100a_calibrated --out profiles/perf_vs_real.json        # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/b200/perf_vs_real.py ...  # DP=8

Cases are data-parallel only (the measured runtime implements DP + ZeRO-1); every row is device-timed (CUDA events,
max over ranks) after warm-up, and memory is the CUDA high-water mark of the process.
(reference: tools/b200/run_megatron_perf_real_pipeline.py / _batch.py drive Megatron-LM and regex its logs; here the
"real" side is in-repo.)
"""
from __future__ import annotations

import argparse
import contextlib
import gc
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

#        name                      model         layers  seq    mbs  mbc
CASES = [
    ("llama3_8b_full_s4096",       "llama3-8b",  None,   4096,  1,   4),
    ("llama3_8b_l16_s4096_mbs2",   "llama3-8b",  16,     4096,  2,   2),
    ("llama3_8b_l16_s8192",        "llama3-8b",  16,     8192,  1,   2),
    ("llama3_70b_l4_s4096",        "llama3-70b", 4,      4096,  1,   4),
    ("llama3_70b_l8_s4096",        "llama3-70b", 8,      4096,  1,   2),
]


TP_CASES = [      # measured with TPLlamaTrainer: tp = world, sequence parallel, dp = 1 (every rank sees the same batch)
    ("llama3_8b_full_s4096_tp",    "llama3-8b",  None,   4096,  1,   4),
    ("llama3_70b_l12_s4096_tp",    "llama3-70b", 12,     4096,  1,   4),
    ("llama3_70b_l12_s8192_tp",    "llama3-70b", 12,     8192,  1,   2),
    # the remaining tp8_dp1 rows of the reference's published B200 table (same model / layers / seq / micro-batches)
    ("llama3_70b_l12_s4096_mbc8_tp",  "llama3-70b", 12,   4096,  1,   8),
    ("llama3_70b_l12_s4096_mbc32_tp", "llama3-70b", 12,   4096,  1,   32),
    ("llama3_405b_l4_s4096_tp",       "llama3-405b_padding_128", 4, 4096, 1, 4),
    ("llama3_405b_l4_s4096_mbc8_tp",  "llama3-405b_padding_128", 4, 4096, 1, 8),
]

# (model, layers, seq, mbs, mbc, tp, dp) -> (Megatron-LM + TE measured ms, peak GiB, reference-predicted ms) as published by the
# reference authors for 8 x B200 (docs/b200/b200_release_v1.2_summary.md).  Quoted, not re-measured here.
PUBLISHED = {
    ("llama3-70b", 12, 4096, 1, 4, 8, 1): (280.9, 28.29, 255.71), ("llama3-70b", 12, 4096, 1, 8, 8, 1): (556.7, 28.29, 493.39),
    ("llama3-70b", 12, 4096, 1, 32, 8, 1): (2073.7, 28.29, 1919.4), ("llama3-70b", 12, 4096, 1, 4, 2, 4): (725.5, 60.68, 698.81),
    ("llama3-70b", 12, 4096, 1, 4, 4, 2): (409.8, 39.07, 395.23), ("llama3-70b", 12, 4096, 1, 8, 2, 4): (1387.0, 60.68, 1309.8),
    ("llama3-70b", 12, 4096, 1, 8, 4, 2): (765.6, 39.07, 747.66),
    ("llama3-405b_padding_128", 4, 4096, 1, 4, 8, 1): (258.8, 38.57, 255.83), ("llama3-405b_padding_128", 4, 4096, 1, 8, 8, 1): (488.1, 38.57, 486.81),
    ("llama3-405b_padding_128", 4, 4096, 1, 4, 2, 4): (907.3, 81.59, 860.73), ("llama3-405b_padding_128", 4, 4096, 1, 4, 4, 2): (476.0, 52.89, 452.92),
}


def predict(model, layers, seq, mbs, mbc, world, system, tp=1, runtime_semantics=False):
    from simumax_b200 import PerfLLM
    from simumax_b200.configs import get_simu_model_config, get_simu_strategy_config, get_simu_system_config
    from simumax_b200.core.config import StrategyConfig
    from simumax_b200.core.utils import parse_bytes
    st = StrategyConfig.init_from_config_file(get_simu_strategy_config("tp1_pp1_dp8_mbs1"))
    st.world_size, st.tp_size, st.pp_size = world, tp, 1
    st.enable_sequence_parallel = tp > 1
    st.seq_len, st.micro_batch_size, st.micro_batch_num = seq, mbs, mbc
    st.dp_overlap_model = True      # the runtime overlaps ZeRO-1 collectives with backward / next forward
    # runtime semantics: TP collectives fused into the GEMMs (measured exposure), single-pass AdamW; off = what Megatron + NCCL do
    st.tp_overlap_model = st.fused_optimizer_model = bool(runtime_semantics)
    p = PerfLLM()
    with contextlib.redirect_stdout(io.StringIO()):
        p.configure(strategy_config=st, model_config=get_simu_model_config(model), system_config=get_simu_system_config(system))
        if layers:
            p.model_config.layer_num = layers
        p.model_config.padded_vocab_size, p.model_config.make_vocab_size_divisible_by = True, 128
        p.run_estimate()
        cost, mem = p.analysis_cost().data, p.analysis_mem().data
    tok_s = cost["throughput_per_accelerator"] * world
    return {"iter_ms": mbs * mbc * (world // tp) * seq / tok_s * 1e3, "tokens_per_s": tok_s, "mfu": cost["mfu"],
            "peak_gib": parse_bytes(mem["peak_mem"]) / 2 ** 30 if isinstance(mem.get("peak_mem"), str) else None,
            "miss": {k: len(v) for k, v in p.system.miss_efficiency.items()}}


def measure(model, layers, seq, mbs, mbc, steps, warmup, tp=1):
    import torch
    import torch.distributed as dist
    from simumax_b200.configs import get_simu_model_config
    from simumax_b200.core.config import ModelConfig
    from simumax_b200.runtime.trainer import LlamaDims, LlamaTrainer
    mc = ModelConfig.init_from_config_file(get_simu_model_config(model))
    dims = LlamaDims.from_model_config(mc)
    if layers:
        dims.layers = layers
    world = dist.get_world_size() if dist.is_initialized() else 1
    torch.cuda.reset_peak_memory_stats()
    if tp > 1:
        from simumax_b200.runtime.tp_trainer import TPLlamaTrainer
        tp_group = dp_group = None
        if world > tp:
            rank, dp = dist.get_rank(), world // tp
            tp_groups = [dist.new_group(list(range(i * tp, (i + 1) * tp))) for i in range(dp)]
            dp_groups = [dist.new_group(list(range(j, world, tp))) for j in range(tp)]
            tp_group, dp_group = tp_groups[rank // tp], dp_groups[rank % tp]
        tr = TPLlamaTrainer(dims, mbs, seq, tp_group=tp_group, dp_group=dp_group)
    else:
        tr = LlamaTrainer(dims, mbs, seq, dp_group=dist.group.WORLD if world > 1 else None)
    g = torch.Generator(device="cuda").manual_seed(5)
    batches = [(torch.randint(0, mc.vocab_size, (mbs * seq,), device="cuda", generator=g),
                torch.randint(0, mc.vocab_size, (mbs * seq,), device="cuda", generator=g)) for _ in range(mbc)]
    for _ in range(warmup):
        tr.run_microbatches(batches)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tr.run_microbatches(batches)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    peak = torch.tensor([float(torch.cuda.max_memory_allocated())], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(peak, op=dist.ReduceOp.MAX)
    out = {"iter_ms": float(ms), "tokens_per_s": mbs * mbc * (world // tp) * seq / float(ms) * 1e3, "peak_gib": float(peak) / 2 ** 30}
    del tr, batches
    gc.collect()
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--system", default="b200_sm100a_calibrated")
    ap.add_argument("--out", default="gpurun_out/perf_vs_real.json")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--cases", nargs="*", default=None)
    ap.add_argument("--tp", action="store_true", help="tensor-parallel cases (sequence parallel, fused kernels)")
    ap.add_argument("--tp-size", type=int, default=0, help="with --tp: tensor-parallel size (default = world size); the remaining "
                    "factor runs data parallel with ZeRO-1, e.g. 8 GPUs --tp-size 2 = the reference's tp2_dp4 rows")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if world > 1 else 0
    rows = []
    tp = (a.tp_size or world) if a.tp else 1
    assert world % tp == 0
    for name, model, layers, seq, mbs, mbc in (TP_CASES if a.tp else CASES):
        if a.cases and name not in a.cases:
            continue
        real = measure(model, layers, seq, mbs, mbc, a.steps, a.warmup, tp)
        if rank == 0:
            pred = predict(model, layers, seq, mbs, mbc, world, a.system, tp)
            pred_rt = predict(model, layers, seq, mbs, mbc, world, a.system, tp, runtime_semantics=True)
            row = {"case": name + (str(tp) if a.tp else ""), "world": world, "tp": tp, "model": model, "layers": layers, "seq": seq, "mbs": mbs, "mbc": mbc,
                   "pred_ms": pred["iter_ms"], "real_ms": real["iter_ms"],
                   "time_err_pct": 100 * (pred["iter_ms"] - real["iter_ms"]) / real["iter_ms"],
                   "pred_gib": pred["peak_gib"], "real_gib": real["peak_gib"],
                   "mem_err_pct": 100 * (pred["peak_gib"] - real["peak_gib"]) / real["peak_gib"] if pred["peak_gib"] else None,
                   "pred_runtime_semantics_ms": pred_rt["iter_ms"],
                   "time_err_runtime_semantics_pct": 100 * (pred_rt["iter_ms"] - real["iter_ms"]) / real["iter_ms"],
                   "pred_mfu": pred["mfu"],
                   "published_megatron": (lambda v: None if v is None else {"real_ms": v[0], "real_gib": v[1], "reference_pred_ms": v[2],
                                                                           "speedup_vs_megatron": v[0] / real["iter_ms"]})(
                       PUBLISHED.get((model, layers, seq, mbs, mbc, tp, world // tp))), "real_tokens_per_s": real["tokens_per_s"], "efficiency_misses": pred["miss"]}
            rows.append(row)
            print(f"{name:28s} n={world} time pred {row['pred_ms']:9.2f} ms real {row['real_ms']:9.2f} ms ({row['time_err_pct']:+6.2f} %, runtime semantics {row['time_err_runtime_semantics_pct']:+6.2f} %)   "
                  f"mem pred {row['pred_gib']:.2f} GiB real {row['real_gib']:.2f} GiB ({row['mem_err_pct']:+6.2f} %)"
                  + (f"   published Megatron {row['published_megatron']['real_ms']:.1f} ms -> {row['published_megatron']['speedup_vs_megatron']:.2f}x"
                     if row["published_megatron"] else ""), flush=True)
    if rank == 0:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump({"system": a.system, "world": world, "rows": rows}, open(a.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
