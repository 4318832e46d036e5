#!/usr/bin/env python
"""Predicted-vs-measured pipeline: run real training steps with this repo's runtime and compare iteration time and peak memory with
what ``PerfLLM`` predicts for the same (system, strategy, model) triple -- the in-repo counterpart of the reference's Megatron
pipeline (tools/b200/run_megatron_perf_real_pipeline.py: phases + state file + ``--reuse``; run_megatron_perf_real_batch.py:587-632,
710-730: per-case summary ``.md`` / ``.json``).

    python tools/b200/perf_vs_real.py                                   # 1 GPU: every case of the zoo that fits one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/b200/perf_vs_real.py --cases llama3_70b_l12_tp2_pp1_dp4_mbc4
    python tools/b200/perf_vs_real.py --phase predict --reuse           # CPU only: re-predict measured cases with another --system

Phases: ``measure`` (GPU; device-timed steps, max over ranks; peak memory = allocator high-water mark + symmetric allocations,
max over ranks) and ``predict`` (CPU).  Every finished phase of every case is recorded in the state file (``--state``); ``--reuse``
skips phases the state already holds, so an interrupted sweep resumes and a re-calibrated system config only costs the predict phase.
Outputs: ``<out>.json`` (rows) and ``<out>.md`` (table in the layout of docs/b200/b200_release_v1.2_summary.md).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from simumax_b200.runtime.cases import Case, predict, published  # noqa: E402


def zoo():
    """The validation rows this repo can execute, keyed by name; ``world`` GPUs each."""
    c = []
    # one GPU
    c += [Case("llama3_8b_tp1_mbc8", "llama3-8b", None, 4096, 1, 8),
          Case("llama3_8b_l16_mbs2_mbc2", "llama3-8b", 16, 4096, 2, 2),
          Case("llama3_8b_l16_s8192_mbc2", "llama3-8b", 16, 8192, 1, 2),
          Case("llama3_70b_l4_mbc4", "llama3-70b", 4, 4096, 1, 4),
          Case("llama3_70b_l8_mbc2", "llama3-70b", 8, 4096, 1, 2),
          Case("llama3_8b_l16_selective_mbc4", "llama3-8b", 16, 4096, 1, 4, recompute="selective",
               strategy_overrides={"enable_recompute": True, "recompute_granularity": "selective_recompute", "recompute_layer_num": 16,
                                   "attn_recompute": True, "mlp_recompute": True}),
          Case("llama3_8b_l16_full_mbc4", "llama3-8b", 16, 4096, 1, 4, recompute="full",
               strategy_overrides={"enable_recompute": True, "recompute_granularity": "full_block", "recompute_layer_num": 16})]
    # tensor parallel (+ data parallel): rows / slices of the published table
    for tp in (2, 4, 8):
        c.append(Case(f"llama3_8b_tp{tp}_pp1_dp1_mbc8", "llama3-8b", None, 4096, 1, 8, tp=tp, strategy_preset="tp8_pp1_dp1_mbs1"))
        c.append(Case(f"llama3_70b_l12_tp{tp}_pp1_dp1_mbc4", "llama3-70b", 12, 4096, 1, 4, tp=tp, strategy_preset="tp8_pp1_dp1_mbs1"))
    c += [Case("llama3_70b_l12_tp2_pp1_dp4_mbc4", "llama3-70b", 12, 4096, 1, 4, tp=2, dp=4, strategy_preset="tp2_pp1_dp4_mbs1"),
          Case("llama3_70b_l12_tp4_pp1_dp2_mbc4", "llama3-70b", 12, 4096, 1, 4, tp=4, dp=2, strategy_preset="tp4_pp1_dp2_mbs1"),
          Case("llama3_70b_l12_tp2_pp1_dp2_mbc4", "llama3-70b", 12, 4096, 1, 4, tp=2, dp=2, strategy_preset="tp2_pp1_dp4_mbs1"),
          Case("llama3_70b_l12_tp8_pp1_dp1_mbc8", "llama3-70b", 12, 4096, 1, 8, tp=8, strategy_preset="tp8_pp1_dp1_mbs1"),
          Case("llama3_405b_l4_tp8_pp1_dp1_mbc4", "llama3-405b_padding_128", 4, 4096, 1, 4, tp=8, strategy_preset="tp8_pp1_dp1_mbs1"),
          Case("llama3_70b_l12_tp2_pp1_dp4_mbc8_selective", "llama3-70b", 12, 4096, 1, 8, tp=2, dp=4, recompute="selective",
               strategy_preset="tp2_pp1_dp4_mbs1_selective_recompute"),
          Case("llama3_70b_l12_tp2_pp1_dp1_mbc4_selective", "llama3-70b", 12, 4096, 1, 4, tp=2, dp=1, recompute="selective",
               strategy_preset="tp2_pp1_dp4_mbs1_selective_recompute"),
          Case("llama3_70b_l12_tp2_pp1_dp1_mbc4_full", "llama3-70b", 12, 4096, 1, 4, tp=2, dp=1, recompute="full",
               strategy_preset="tp2_pp1_dp4_mbs1_full_recompute")]
    # pipeline parallel
    c += [Case("llama3_8b_tp1_pp2_dp1_mbc8", "llama3-8b", None, 4096, 1, 8, pp=2, strategy_preset="tp1_pp2_dp4_mbs1"),
          Case("llama3_8b_tp1_pp2_dp4_mbc8", "llama3-8b", None, 4096, 1, 8, pp=2, dp=4, strategy_preset="tp1_pp2_dp4_mbs1"),
          Case("llama3_70b_l12_tp1_pp2_dp4_mbc4", "llama3-70b", 12, 4096, 1, 4, pp=2, dp=4, strategy_preset="tp1_pp2_dp4_mbs1"),
          Case("llama3_8b_tp1_pp2_dp2_mbc8", "llama3-8b", None, 4096, 1, 8, pp=2, dp=2, strategy_preset="tp1_pp2_dp4_mbs1"),
          Case("llama3_8b_tp1_pp4_dp1_mbc8", "llama3-8b", None, 4096, 1, 8, pp=4, dp=1, strategy_preset="tp1_pp2_dp4_mbs1")]
    # context parallel (Ulysses all-to-all), long sequences: rows / slices of the published CP table
    c += [Case("llama3_70b_l12_tp1_cp8_seq32768_mbc4", "llama3-70b", 12, 32768, 1, 4, cp=8),
          Case("llama3_70b_l12_tp1_cp8_seq131072_mbc4", "llama3-70b", 12, 131072, 1, 4, cp=8),
          Case("llama3_8b_tp1_cp8_seq32768_mbc4", "llama3-8b", None, 32768, 1, 4, cp=8),
          Case("llama3_8b_l16_tp1_cp4_seq32768_mbc2", "llama3-8b", 16, 32768, 1, 2, cp=4),
          Case("llama3_8b_l16_tp1_cp2_seq16384_mbc2", "llama3-8b", 16, 16384, 1, 2, cp=2)]
    # expert parallel (DeepSeek-V2: MLA + MoE)
    c += [Case("deepseekv2_l4_ep8_pp1_dp8_mbc4", "deepseekv2", 4, 4096, 1, 4, ep=8, dp=8, strategy_preset="ep8_pp1_dp8_mbs1"),
          Case("deepseekv2_l4_ep2_pp1_dp2_mbc4", "deepseekv2", 4, 4096, 1, 4, ep=2, dp=2, strategy_preset="ep8_pp1_dp8_mbs1"),
          Case("deepseekv2_l4_ep4_pp1_dp4_mbc4", "deepseekv2", 4, 4096, 1, 4, ep=4, dp=4, strategy_preset="ep8_pp1_dp8_mbs1"),
          Case("deepseekv2_l2_ep1_pp1_dp1_mbc2", "deepseekv2", 2, 4096, 1, 2, ep=1, dp=1, strategy_preset="ep8_pp1_dp8_mbs1")]
    return {x.name: x for x in c}


MOE_OVERRIDES = {"moe_pad_expert_input_to_capacity": True, "capacity": 1}


def measure(case: Case, steps: int, warmup: int) -> dict:
    import torch
    import torch.distributed as dist
    from simumax_b200.runtime import cases as C
    mc, _ = C.model_dims(case)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    torch.cuda.reset_peak_memory_stats()
    tr = C.build_trainer(case)
    g = torch.Generator(device="cuda").manual_seed(5 + rank // (case.tp * case.pp))
    T = case.mbs * case.seq // case.cp                   # context parallel: every rank feeds its slice of the sequence
    batches = [(torch.randint(0, mc.vocab_size, (T,), device="cuda", generator=g),
                torch.randint(0, mc.vocab_size, (T,), device="cuda", generator=g)) for _ in range(case.mbc)]
    for _ in range(warmup):
        tr.run_microbatches(batches)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = tr.run_microbatches(batches)
    e1.record()
    torch.cuda.synchronize()
    mem = C.device_memory_gib(tr)
    vals = torch.tensor([e0.elapsed_time(e1) / steps, mem["peak_gib"], mem["symmetric_gib"]], device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    out = {"iter_ms": float(vals[0]), "tokens_per_s": case.tokens_per_step / float(vals[0]) * 1e3, "peak_gib": float(vals[1]),
           "symmetric_gib": float(vals[2]), "loss": float(loss.item()), "trainer": type(tr).__name__, "steps": steps, "warmup": warmup}
    C.release(tr)
    del tr, batches
    import gc
    gc.collect(); torch.cuda.empty_cache()
    return out


def row_of(case: Case, st: dict) -> dict:
    m, p = st.get("measure"), st.get("predict")
    row = {"case": case.name, "world": case.world, "parallelism": case.parallelism(), "model": case.model, "layers": case.layers,
           "seq": case.seq, "mbs": case.mbs, "mbc": case.mbc}
    if m:
        row.update(real_ms=m["iter_ms"], real_gib=m["peak_gib"], real_tokens_per_s=m["tokens_per_s"], trainer=m.get("trainer"))
    if p:
        row.update(pred_ms=p["iter_ms"], pred_gib=p["peak_gib"], system=p["system"], efficiency_misses=p.get("efficiency_misses"),
                   pred_megatron_semantics_ms=p.get("megatron_semantics_ms"))
    if m and p:
        row["time_err_pct"] = 100 * (p["iter_ms"] - m["iter_ms"]) / m["iter_ms"]
        row["mem_err_pct"] = 100 * (p["peak_gib"] - m["peak_gib"]) / m["peak_gib"] if p.get("peak_gib") else None
    pub = published(case)
    if pub:
        row["published_megatron"] = {"real_ms": pub[0], "real_gib": pub[1], "reference_pred_ms": pub[2],
                                     "reference_time_err_pct": 100 * (pub[2] - pub[0]) / pub[0]}
        if m:
            row["published_megatron"]["speedup_vs_megatron"] = pub[0] / m["iter_ms"]
    return row


def write_summary(rows, out_stem: str, system: str) -> None:
    json.dump({"system": system, "generated": time.strftime("%Y-%m-%d %H:%M:%S"), "rows": rows}, open(out_stem + ".json", "w"), indent=1)
    f = lambda v, fmt="{:.2f}": "-" if v is None else fmt.format(v)
    lines = [f"# Predicted vs measured (system config `{system}`)", "",
             "Measured = this repo's runtime on the hand-written sm_100a kernels (device-timed, max over ranks; memory = allocator high-water "
             "mark + symmetric allocations). Error = (predicted - measured) / measured.  `Megatron (published)` quotes the reference authors' "
             "8 x B200 Megatron-LM + TE measurement of the same row (BASELINE.md section 2), not re-measured here.", "",
             "| case | GPUs | layout | real ms | perf ms | time err | real GiB | perf GiB | mem err | Megatron (published) ms | speed-up |",
             "|---|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|"]
    errs, merrs = [], []
    for r in rows:
        pub = r.get("published_megatron") or {}
        lines.append(f"| {r['case']} | {r['world']} | {r['parallelism']} | {f(r.get('real_ms'))} | {f(r.get('pred_ms'))} | "
                     f"{f(r.get('time_err_pct'), '{:+.2f}%')} | {f(r.get('real_gib'))} | {f(r.get('pred_gib'))} | {f(r.get('mem_err_pct'), '{:+.2f}%')} | "
                     f"{f(pub.get('real_ms'))} | {f(pub.get('speedup_vs_megatron'), '{:.2f}x')} |")
        if r.get("time_err_pct") is not None:
            errs.append(r["time_err_pct"])
        if r.get("mem_err_pct") is not None:
            merrs.append(r["mem_err_pct"])
    if errs:
        lines += ["", f"Derived: n = {len(errs)}, time error range {min(errs):+.2f}% ... {max(errs):+.2f}%, mean |err| {sum(abs(e) for e in errs) / len(errs):.2f}%"
                  + (f"; memory error range {min(merrs):+.2f}% ... {max(merrs):+.2f}%" if merrs else "") + "."]
    open(out_stem + ".md", "w").write("\n".join(lines) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--system", default="b200_sm100a_calibrated")
    ap.add_argument("--out", default="gpurun_out/perf_vs_real", help="output stem: <out>.json, <out>.md, default state <out>.state.json")
    ap.add_argument("--state", default=None)
    ap.add_argument("--reuse", action="store_true", help="skip phases already recorded in the state file")
    ap.add_argument("--phase", default="all", choices=["all", "measure", "predict"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--cases", nargs="*", default=None, help="case names (default: every zoo case whose world size matches the launch)")
    ap.add_argument("--list", action="store_true")
    a = ap.parse_args()
    Z = zoo()
    if a.list:
        for n, c in Z.items():
            print(f"{n:48s} world {c.world}  {c.parallelism()}")
        return 0
    state_path = a.state or a.out + ".state.json"
    # --reuse: keep (and skip) what the state file holds; a predict-only run always starts from the recorded measurements
    state = json.load(open(state_path)) if ((a.reuse or a.phase == "predict") and os.path.exists(state_path)) else {}
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    names = a.cases or [n for n, c in Z.items() if c.world == world or (a.phase == "predict" and n in state)]
    unknown = [n for n in names if n not in Z]
    assert not unknown, f"unknown cases {unknown}; --list shows the zoo"
    do_measure = a.phase in ("all", "measure")
    do_predict = a.phase in ("all", "predict")
    if do_measure:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        if "RANK" not in os.environ:           # DeepSeek / symmetric-memory trainers want a process group even on one GPU
            os.environ.update(RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=os.environ.get("MASTER_PORT", "29571"))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    save = lambda: (rank == 0) and (os.makedirs(os.path.dirname(state_path) or ".", exist_ok=True), json.dump(state, open(state_path, "w"), indent=1))
    for n in names:
        case = Z[n]
        st = state.setdefault(n, {})
        if do_measure and case.world == world and not (a.reuse and "measure" in st):
            try:
                st["measure"] = measure(case, a.steps, a.warmup)
                st.pop("measure_error", None)
            except Exception as e:                               # one failing case must not stop the sweep
                st["measure_error"] = f"{type(e).__name__}: {e}"[:300]
                import torch
                torch.cuda.empty_cache()
            save()
        if do_predict and rank == 0 and not (a.reuse and st.get("predict", {}).get("system") == a.system):
            try:
                if case.model.startswith("deepseek"):
                    case.strategy_overrides = dict(case.strategy_overrides)
                p = predict(case, system=a.system, runtime_semantics=True)
                p["megatron_semantics_ms"] = predict(case, system=a.system, runtime_semantics=False)["iter_ms"]
                st["predict"] = p
            except Exception as e:
                st["predict_error"] = f"{type(e).__name__}: {e}"[:300]
            save()
        if rank == 0:
            r = row_of(case, st)
            print(f"{n:46s} real {r.get('real_ms', float('nan')):9.2f} ms  perf {r.get('pred_ms', float('nan')):9.2f} ms  "
                  f"({r.get('time_err_pct', float('nan')):+6.2f} %)   mem real {r.get('real_gib', float('nan')):7.2f} perf {r.get('pred_gib', float('nan')):7.2f} GiB"
                  + (f"  ERROR {st.get('measure_error') or st.get('predict_error')}" if (st.get('measure_error') or st.get('predict_error')) else ""), flush=True)
    if rank == 0:
        rows = [row_of(Z[n], state[n]) for n in state if n in Z and (state[n].get("measure") or state[n].get("predict"))]
        write_summary(rows, a.out, a.system)
        print(f"wrote {a.out}.json / .md ({len(rows)} rows), state {state_path}")
    if do_measure:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main() or 0)
